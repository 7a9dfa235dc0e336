// Device context, launch logic and the device-facing part of the C ABI.
//
// Replaces the orchestration inside one iteration of the reference's hot loops
// (run.rs:122-143 SE, run.rs:249-274 PE): "batch of records in -> ordered batch of rewritten
// records out".  Everything per-read happens in extract_kernel (bk_kernels.cuh); there is no CPU
// implementation of it here.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/barkit_b200.h"
#include "bk_internal.h"
#include "bk_extract.cuh"
#include "bk_tokenize.cuh"
#include "bk_compact.cuh"
#include "bk_match.cuh"

namespace {

constexpr int kSlots = 2;
constexpr size_t kMaxDynSmem = 200 * 1024;    // leave room for >= 1 CTA/SM with static smem
constexpr uint32_t kPassMinRecords = 200000;  // batches large enough to judge the host copy against the bus
constexpr size_t kTargetDynSmem = 101 * 1024 + 512;  // aim for >= 3 resident CTAs (input ring + output tile) per SM

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct Scratch {  // [tile counter | DevResult | desc stream 0 | desc stream 1]
  DevBuf buf;
  uint32_t ntiles_cap = 0;
};

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evh0 = nullptr, evh1 = nullptr;  // around the host -> device copies of a pass-through batch
  DevBuf text[2], off[2], out[2];
  Scratch scratch;
  bk::DevResult* h_res = nullptr;  // pinned
  uint32_t n = 0;
  uint64_t out_cap[2] = {0, 0};
  bool busy = false;
  uint32_t launches = 0;
  // host pass-through of the mate without a pattern (see bk_submit)
  bool passthrough = false;
  int pt = -1;                   // index of the mate kept on the host for the batch in flight
  bk_mate_in host_pt = {};       // the caller's buffers of that mate: read again in bk_wait
  DevBuf match;                  // the other mate's match table on the device
  int32_t* h_match = nullptr;    // ... and its pinned host copy
  size_t h_match_cap = 0;
  uint64_t h2d_bytes = 0;        // of the batch in flight
  // two-kernel form of a paired launch with one pattern (enqueue_split): hand-over and the compaction's look-back state
  DevBuf split_match, split_hlen, split_scratch;
  DevBuf tab_err;
  DevBuf tab_extras[2];     // ... and, for a program without a fixed shape, the spans of every match (16 bytes per read)
  DevBuf tab_match[2];      // match tables of a paired batch that runs as match_kernel + two single-end launches (enqueue_tabled)
  bk::CompactParams cpar;   // the compaction in flight, should it have to run again (compact_redo)
  bool cpar_valid = false;
  // raw-text batches (bk_submit_text): record offsets are made on the device
  bool raw = false;
  DevBuf tok[2];                      // per mate: TokResult | ticket | look-back descriptors
  bk::TokResult* h_tok = nullptr;     // pinned, one per mate
  uint64_t raw_len[2] = {0, 0};
  bool raw_no_eol[2] = {false, false};  // the mate's last record has no line end
  uint32_t raw_maxrec[2] = {0, 0};    // what the launch in flight was planned with
};

}  // namespace

struct bk_ctx {
  int device = 0;
  uint32_t flags = 0;
  bool paired = false;
  bk_devprog prog[2];
  // bounded variable repetition: a mate's pattern as several fixed-length programs (prog[m] is the first)
  std::vector<bk_devprog> var[2];
  DevBuf vdev[2];          // ... and their copy in device memory (nullptr when there is one variant)
  DevBuf vmdev[2];         // the general form (prog[m].vm_len > 0): its instructions in device memory
  uint32_t growth[2] = {0, 0};  // largest header growth over the variants
  int sm_count = 0;
  Slot slot[kSlots];
  DevBuf tok;  // bk_tokenize_device scratch: block counts + TokResult
  // Paired mode with exactly one pattern: the mate without one stays on the host (pass_mate = its index)
  // and `solo` is the single-end view of this context that processes the other mate.
  int pass_mate = -1;
  bk_ctx* solo = nullptr;
  // Paired mode with two patterns: the single-end views of the two mates (enqueue_tabled)
  bk_ctx* view[2] = {nullptr, nullptr};
  int host_threads = 1;
  uint32_t maxrec_seen[2] = {0, 0};  // longest record of earlier raw-text batches, per mate (bk_submit_text)
  bool compact_careful = false;      // a batch's pattern-less mate was not in the writer's framing: size from the text first
  // The host copy is worth it while it takes less time than the bus would need for the same mate.
  // Both are measured on every large pass-through batch; when the host side is the slower one twice in
  // a row (few or slow host cores, several GPUs sharing the host's memory bandwidth) the context
  // sends both mates to the device from then on.  BK_FLAG_HOST_MATE_PINNED pins the host copy.
  bool pass_pinned = false;
  int pass_slow = 0;
  std::string last_error;
};

namespace {

bool cuda_ok(bk_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  ctx->last_error = std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e);
  return false;
}
#define CK(call)                                  \
  do {                                            \
    if (!cuda_ok(ctx, (call), #call)) return BK_E_CUDA; \
  } while (0)

int ensure(bk_ctx* ctx, DevBuf* b, size_t bytes) {
  if (b->cap >= bytes && b->p) return BK_OK;
  if (b->p) CK(cudaFree(b->p));
  b->p = nullptr;
  b->cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CK(cudaMalloc(&b->p, want));
  b->cap = want;
  return BK_OK;
}

struct LaunchPlan {
  uint32_t tile_recs;
  uint32_t ntiles;
  uint32_t in_cap[2], out_cap[2];
  size_t smem;
  int grid;
};

// A parametric program's header growth depends on the read: every extra copy of an atom inside a capture group adds a
// sequence and a quality byte to the header.  Upper bound for a record of at most `maxrec` bytes (its sequence line
// is less than half of it).
uint64_t par_growth(const bk_devprog& pg, uint64_t maxrec) {
  uint64_t g = 0;
  for (uint32_t i = 0; i < pg.nins; ++i) {
    const uint64_t e = std::min<uint64_t>(pg.ins[i].max_extra == 0xFFFF ? maxrec / 2 : pg.ins[i].max_extra, maxrec / 2);
    g += 2 * e * (uint64_t)__builtin_popcount(pg.ins[i].capmask);
  }
  if (pg.vm_len)  // the general form: caps[c].len = the longest span the capture can have
    for (uint32_t c = 0; c < pg.ncaps; ++c) g += 2 * std::min<uint64_t>(pg.caps[c].len, maxrec / 2);
  return g;
}
// ... and summed over a batch of n records with text_len bytes in all: a capture grows by at most its bound per record
// and by no more than the reads' own bytes over the batch
uint64_t par_growth_total(const bk_devprog& pg, uint64_t n, uint64_t text_len) {
  uint64_t g = 0;
  for (uint32_t i = 0; i < pg.nins; ++i) {
    const uint64_t per = pg.ins[i].max_extra == 0xFFFF ? text_len : 2 * (uint64_t)pg.ins[i].max_extra * n;
    g += std::min(per, text_len) * (uint64_t)__builtin_popcount(pg.ins[i].capmask);
  }
  if (pg.vm_len)
    for (uint32_t c = 0; c < pg.ncaps; ++c)
      g += std::min(pg.caps[c].len == 0xFFFF ? text_len : 2 * (uint64_t)pg.caps[c].len * n, text_len);
  return g;
}

size_t smem_for(const bk_ctx* ctx, uint32_t R, const uint32_t maxrec[2], uint32_t in_cap[2],
                uint32_t out_cap[2], int stages) {
  size_t total = 0;
  int nm = ctx->paired ? 2 : 1;
  for (int m = 0; m < 2; ++m) in_cap[m] = out_cap[m] = 0;
  for (int m = 0; m < nm; ++m) {
    uint32_t growth = ctx->prog[m].present ? ctx->growth[m] + (uint32_t)par_growth(ctx->prog[m], maxrec[m]) : 0;
    // A trimmed match takes its own bytes out of the seq and out of the qual line (parse.rs:138-148), so a record of a
    // fixed-shape program grows by its header growth LESS twice the match length, if at all (C3: 67 - 70: the output
    // tile needs no more room than the input records; every byte saved here is a byte for the input ring).
    if (ctx->prog[m].present && !(ctx->flags & BK_FLAG_SKIP_TRIM) && !bk_free_shape(ctx->prog[m]) && !ctx->var[m].empty()) {
      uint32_t net = 0;
      for (const bk_devprog& v : ctx->var[m]) net = std::max(net, v.hdr_growth > 2 * v.total_len ? v.hdr_growth - 2 * v.total_len : 0u);
      growth = net;
    }
    in_cap[m] = (uint32_t)(((size_t)R * maxrec[m] + 32 + 15) & ~size_t(15));
    out_cap[m] = (uint32_t)(((size_t)R * ((size_t)maxrec[m] + growth + 2) + 32 + 15) & ~size_t(15));
    total += (size_t)stages * in_cap[m] + out_cap[m];
  }
  return total;
}

// extract_kernel<PAIRED, GENERAL>: GENERAL iff a pattern needs a search (not anchored) or -r is set
bool has_variants(const bk_ctx* ctx) { return ctx->var[0].size() > 1 || ctx->var[1].size() > 1; }

bool needs_general(const bk_ctx* ctx) {
  if (ctx->flags & BK_FLAG_RC) return true;
  if (has_variants(ctx)) return true;
  for (int m = 0; m < 2; ++m)
    if (ctx->prog[m].present && !ctx->prog[m].anchor_start && !ctx->prog[m].anchor_end) return true;
  return false;
}

template <bool PAIRED, bool GENERAL, bool VAR = false, int ST = bk::kStages, bool TABLE = false>
int plan_launch(bk_ctx* ctx, uint32_t n, const uint32_t maxrec[2], LaunchPlan* lp) {
  constexpr int kThreads = bk::cta_warps(PAIRED) * 32;
  // Two CTAs per SM (228 KB of shared memory per SM, each CTA's static part included) whenever the records
  // allow it: the kernel is bound by ring slots in flight, and the instantiations differ in static shared
  // memory (the general ones carry the search-sharing buffers), so the budget is taken from the kernel itself.
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, bk::extract_kernel<PAIRED, GENERAL, VAR, ST, TABLE>));
  constexpr int kWant = PAIRED ? 2 : 4;  // CTAs per SM the kernel is compiled for (launch bounds): 30-40 warps per SM
  const size_t target = std::min<size_t>(kTargetDynSmem, (233472 / kWant - fa.sharedSizeBytes) & ~size_t(127));
  uint32_t R = bk::kTileRecs;
  // prefer several resident tiles per SM; never go below 8 records unless a tile would not fit
  while (R > 8 && smem_for(ctx, R, maxrec, lp->in_cap, lp->out_cap, ST) > target) --R;
  while (R > 0 && smem_for(ctx, R, maxrec, lp->in_cap, lp->out_cap, ST) > kMaxDynSmem) --R;
  if (R == 0) {
    ctx->last_error = "record too long for the on-chip staging buffer";
    return BK_E_CAPACITY;
  }
  CK(cudaFuncSetAttribute(bk::extract_kernel<PAIRED, GENERAL, VAR, ST, TABLE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                          cudaSharedmemCarveoutMaxShared));
  int per_sm = 0, other = 0;  // `other`: CTAs per SM that registers and threads allow
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&other, bk::extract_kernel<PAIRED, GENERAL, VAR, ST, TABLE>, kThreads, 0));
  for (;;) {
    lp->tile_recs = R;
    lp->smem = smem_for(ctx, R, maxrec, lp->in_cap, lp->out_cap, ST);
    CK(cudaFuncSetAttribute(bk::extract_kernel<PAIRED, GENERAL, VAR, ST, TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)lp->smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bk::extract_kernel<PAIRED, GENERAL, VAR, ST, TABLE>,
                                                     kThreads, lp->smem));
    if (per_sm >= kWant || other < kWant || R <= 24) break;  // (long records: the tile stays as it is, fewer CTAs share the SM)
    --R;
  }
  lp->ntiles = (n + R - 1) / R;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > kWant) per_sm = kWant;
  long long grid = (long long)per_sm * ctx->sm_count;
  if (grid > (long long)lp->ntiles) grid = lp->ntiles;
  lp->grid = (int)std::max<long long>(grid, 1);
  if (ctx->flags & BK_FLAG_VERBOSE)
    fprintf(stderr, "[bk] tile_recs %u smem %zu B ctas/sm %d grid %d ntiles %u\n", lp->tile_recs, lp->smem, per_sm,
            lp->grid, lp->ntiles);
  return BK_OK;
}

// [512 B header | tile descriptors, 8 B each | block descriptors (one per 32 tiles), 16 B each | block accumulators, 16 B each]
size_t block_desc_offset(uint32_t ntiles) { return 512 + (((size_t)ntiles * 8 + 15) & ~size_t(15)); }
size_t block_acc_offset(uint32_t ntiles) { return block_desc_offset(ntiles) + ((size_t)ntiles / 32 + 1) * 16; }
size_t events_offset(uint32_t ntiles) { return block_acc_offset(ntiles) + ((size_t)ntiles / 32 + 1) * 16; }
#ifdef BK_PROFILE_PHASES
size_t scratch_bytes(uint32_t ntiles) { return events_offset(ntiles) + (size_t)ntiles * 128; }  // + event trace
#else
size_t scratch_bytes(uint32_t ntiles) { return events_offset(ntiles); }
#endif

// enqueue: scratch reset + extract kernel.  All pointers are device pointers.
// `co`: the launch is the first half of a split paired batch (enqueue_split): single-end, keep masks for the compaction
// that follows it, figures filed under the mate's own index.
struct CoLaunch {
  uint32_t* hlen = nullptr;              // in: where the header lengths go (nullable)
  uint32_t rslot = 0;                    // in: 0 / 1
  bool want_kmask = false;               // in
  unsigned long long* kmask = nullptr;   // out
  uint32_t tile_recs = 0;                // out
  // match tables (enqueue_tabled): this mate's own, made by match_kernel (the launch then matches nothing itself),
  // and the other mate's of the pair (keep rule)
  const int32_t* tab_own = nullptr;
  const uint4* xtab = nullptr;  // with tab_own, for a parametric program
  // a PAIRED launch in table form: ptab[m] = mate m's table (null: the mate matches for itself), pxtab[m] its extras
  bool pair_tables = false;
  const int32_t* ptab[2] = {nullptr, nullptr};
  const uint4* pxtab[2] = {nullptr, nullptr};
  const int32_t* tab_other = nullptr;
  DevBuf* scratch = nullptr;             // in: the launch's look-back state goes here instead of the slot's buffer ...
  bk::DevResult* result = nullptr;       // in: ... and its figures into this result, which it does not reset
  bool no_nout = false;                  // in: the other mate's launch counts the kept pairs
};
int enqueue_extract(bk_ctx* ctx, Slot* s, cudaStream_t stream, uint32_t n, const bk_mate_in* in1,
                    const bk_mate_in* in2, uint8_t* out1, uint64_t cap1, uint8_t* out2, uint64_t cap2,
                    int32_t* match1, int32_t* match2, uint32_t* launches, CoLaunch* co = nullptr);

int enqueue_tabled(bk_ctx* ctx, Slot* s, cudaStream_t stream, uint32_t n, const bk_mate_in* in1, const bk_mate_in* in2,
                   uint8_t* out1, uint64_t cap1, uint8_t* out2, uint64_t cap2, int32_t* match1, int32_t* match2,
                   uint32_t* launches);
int enqueue_match(bk_ctx* ctx, int m, Slot* s, int ti, cudaStream_t stream, uint32_t n, const bk_mate_in* in,
                  bk::DevResult* result, uint32_t* launches);
// a mate whose pattern needs more than one anchored test per read
bool is_hard(const bk_ctx* ctx, int m) {
  const bk_devprog& pg = ctx->prog[m];
  return pg.present && ((ctx->flags & BK_FLAG_RC) || ctx->var[m].size() > 1 || bk_free_shape(pg) || (!pg.anchor_start && !pg.anchor_end));
}
// parametric programs (unbounded repetition) are only ever matched by match_kernel
bool is_parametric(const bk_ctx* ctx) { return bk_free_shape(ctx->prog[0]) || bk_free_shape(ctx->prog[1]); }

// A paired launch with exactly one pattern, as two kernels (bk_compact.cuh): the single-end extract kernel on the
// pattern's mate, leaving its match table and header lengths on the device, then the ordered compaction of the
// other mate.  `ctx->solo` is the single-end view of the context.
int enqueue_split(bk_ctx* ctx, Slot* s, cudaStream_t stream, uint32_t n, const bk_mate_in* in1, const bk_mate_in* in2,
                  uint8_t* out1, uint64_t cap1, uint8_t* out2, uint64_t cap2, int32_t* match1, int32_t* match2,
                  uint32_t* launches) {
  const int pm = ctx->prog[0].present ? 0 : 1, pt = 1 - pm;
  const bk_mate_in* in[2] = {in1, in2};
  uint8_t* const out[2] = {out1, out2};
  const uint64_t cap[2] = {cap1, cap2};
  int32_t* const user_match[2] = {match1, match2};
  if (!in[pt] || !in[pt]->text || !in[pt]->rec_off) { ctx->last_error = "null batch pointer"; return BK_E_ARG; }
  if (in[pt]->text_len >= (1ull << 32)) { ctx->last_error = "batch text must be < 4 GiB per mate"; return BK_E_CAPACITY; }
  if (((uintptr_t)in[pt]->text & 15) != 0) { ctx->last_error = "batch text must be 16-byte aligned"; return BK_E_ARG; }
  int rc;
  if ((rc = ensure(ctx, &s->split_hlen, ((size_t)n + 1) * 4)) != BK_OK) return rc;
  // everything the compaction needs zeroed goes in front of the extract kernel: the two kernels must be adjacent in
  // the stream for the programmatic dependent launch
  const uint32_t ntiles = n ? (n + bk::kCompactRecs - 1) / bk::kCompactRecs : 0;
  uint8_t* sc = nullptr;
  if (n) {
    if ((rc = ensure(ctx, &s->split_scratch, scratch_bytes(ntiles))) != BK_OK) return rc;
    sc = (uint8_t*)s->split_scratch.p;
    CK(cudaMemsetAsync(sc, 0, scratch_bytes(ntiles), stream));
  }
  if (user_match[pt]) CK(cudaMemsetAsync(user_match[pt], 0xFF, (size_t)n * 4, stream));  // no pattern: nothing matches (-1)
  CoLaunch co;
  co.hlen = (uint32_t*)s->split_hlen.p;
  co.rslot = (uint32_t)pm;
  co.want_kmask = true;
  rc = enqueue_extract(ctx->solo, s, stream, n, in[pm], nullptr, out[pm], cap[pm], nullptr, 0, user_match[pm], nullptr,
                       launches, &co);
  if (rc != BK_OK) { ctx->last_error = ctx->solo->last_error; return rc; }
  if (n == 0) return BK_OK;
  bk::CompactParams C;
  memset(&C, 0, sizeof C);
  C.text = in[pt]->text;
  C.off = in[pt]->rec_off;
  C.kmask = co.kmask;
  C.kmask_recs = co.tile_recs;
  C.hlen = co.hlen;
  C.out = out[pt];
  C.cap = cap[pt];
  C.desc[0] = (unsigned long long*)(sc + 512);
  C.desc[1] = (unsigned long long*)(sc + block_desc_offset(ntiles));
  C.desc[2] = (unsigned long long*)(sc + block_acc_offset(ntiles));
  C.ticket = (unsigned int*)sc;
  C.result = (bk::DevResult*)((uint8_t*)s->scratch.buf.p + 64);  // the single-end launch's
  C.n = n;
  C.ntiles = ntiles;
  C.text_len = (uint32_t)in[pt]->text_len;
  C.stream = (uint32_t)pt;
  // (Launched behind the extract kernel in the ordinary way.  Running the two side by side -- programmatic dependent
  // launch, the single-end kernel at three CTAs per SM with a 6-deep ring, this kernel following it tile by tile
  // through the keep masks -- was built and measured: 3.66 ms against 3.11 ms per 10 M pairs of C3.  The few CTAs of
  // this kernel that fit next to the extract kernel's have too few bytes in flight to keep up, and what they do add
  // lengthens the extract kernel's memory round trips, which is what bounds it.  DESIGN.md section 4c.)
  if (ctx->compact_careful) bk::compact_kernel<true><<<ntiles, bk::kCompactThreads, 0, stream>>>(C);
  else bk::compact_kernel<false><<<ntiles, bk::kCompactThreads, 0, stream>>>(C);
  CK(cudaGetLastError());
  s->cpar = C;
  s->cpar_valid = true;
  if (launches) *launches += 1;
  return BK_OK;
}

int enqueue_extract(bk_ctx* ctx, Slot* s, cudaStream_t stream, uint32_t n, const bk_mate_in* in1,
                    const bk_mate_in* in2, uint8_t* out1, uint64_t cap1, uint8_t* out2, uint64_t cap2,
                    int32_t* match1, int32_t* match2, uint32_t* launches, CoLaunch* co) {
  if (ctx->paired && !in2) { ctx->last_error = "paired context needs two mates"; return BK_E_ARG; }
  if (ctx->paired && ctx->solo && (is_parametric(ctx) || !(ctx->flags & BK_FLAG_STAGE_BOTH_MATES)))
    return enqueue_split(ctx, s, stream, n, in1, in2, out1, cap1, out2, cap2, match1, match2, launches);
  if (!ctx->paired && in2) { ctx->last_error = "single-end context got two mates"; return BK_E_ARG; }
  const bk_mate_in* in[2] = {in1, in2};
  uint32_t maxrec[2] = {0, 0};
  for (int m = 0; m < (ctx->paired ? 2 : 1); ++m) {
    if (!in[m] || !in[m]->text || !in[m]->rec_off) { ctx->last_error = "null batch pointer"; return BK_E_ARG; }
    if (in[m]->text_len >= (1ull << 32)) { ctx->last_error = "batch text must be < 4 GiB per mate"; return BK_E_CAPACITY; }
    if (((uintptr_t)in[m]->text & 15) != 0) { ctx->last_error = "batch text must be 16-byte aligned"; return BK_E_ARG; }
    if (in[m]->max_rec_bytes == 0) { ctx->last_error = "max_rec_bytes must be set"; return BK_E_ARG; }
    maxrec[m] = in[m]->max_rec_bytes;
  }
  if (((uintptr_t)out1 & 15) != 0 || ((uintptr_t)out2 & 15) != 0) {  // the tile stores are 16-byte bulk copies
    ctx->last_error = "output buffers must be 16-byte aligned";
    return BK_E_ARG;
  }
  // (two anchored patterns stay in the paired kernel: 1.51 against 1.77 ms per 4 M pairs for the split form)
  if (ctx->paired && ctx->view[0] && !(co && co->pair_tables) && (is_hard(ctx, 0) || is_hard(ctx, 1)) &&
      (is_parametric(ctx) || !(ctx->flags & (BK_FLAG_STAGE_BOTH_MATES | BK_FLAG_MATCH_IN_KERNEL))))
    return enqueue_tabled(ctx, s, stream, n, in1, in2, out1, cap1, out2, cap2, match1, match2, launches);
  // A single-end launch whose pattern needs a search: match_kernel first, this launch takes its matches from the table
  // (unless the caller has done that already: enqueue_tabled)
  const int32_t* self_tab = nullptr;
  const uint4* self_xtab = nullptr;
  bk::DevResult* merr = nullptr;
  // -- when its pattern is a list of variants or a parametric program: those matchers are lane-per-read loops that
  // gain from match_kernel's occupancy (unanchored {m,n}: 1.75 x) and must not sit next to the emit code; a plain
  // unanchored pattern or a -r retry is as fast or faster inside the extract kernel (one read of the text).
  const bool self_match = !ctx->paired && n && !(co && co->tab_own) &&
                          (bk_free_shape(ctx->prog[0]) || (ctx->var[0].size() > 1 && !(ctx->flags & BK_FLAG_MATCH_IN_KERNEL)));
  if (self_match) {
    const int ti = co ? (int)co->rslot : 0;
    int rcm;
    if ((rcm = ensure(ctx, &s->tab_err, 64)) != BK_OK) return rcm;
    merr = (bk::DevResult*)s->tab_err.p;
    CK(cudaMemsetAsync(merr, 0, sizeof(bk::DevResult), stream));
    if ((rcm = enqueue_match(ctx, 0, s, ti, stream, n, in1, merr, launches)) != BK_OK) return rcm;
    self_tab = (const int32_t*)s->tab_match[ti].p;
    self_xtab = (const uint4*)s->tab_extras[ti].p;
  }
  LaunchPlan lp;
  const bool table = self_tab || (co && (co->tab_own || co->pair_tables));
  const bool general = !table && needs_general(ctx);
  const bool var = has_variants(ctx);
  // which instantiation: 0..5 as before, 6 / 7 = single-end with the matches taken from a table (plain / variants)
  const int kid = table && ctx->paired ? (var || is_parametric(ctx) ? 9 : 8) : table ? (var || bk_free_shape(ctx->prog[0]) ? 7 : 6) : var ? (ctx->paired ? 0 : 1) : ctx->paired ? (general ? 2 : 3) : (general ? 4 : 5);
  bk::KParams P;
  memset(&P, 0, sizeof P);
  int rc;
  switch (kid) {
    case 0: rc = plan_launch<true, true, true>(ctx, n, maxrec, &lp); break;
    case 1: rc = plan_launch<false, true, true>(ctx, n, maxrec, &lp); break;
    case 2: rc = plan_launch<true, true>(ctx, n, maxrec, &lp); break;
    case 3: rc = plan_launch<true, false>(ctx, n, maxrec, &lp); break;
    case 4: rc = plan_launch<false, true>(ctx, n, maxrec, &lp); break;
    case 5: rc = plan_launch<false, false>(ctx, n, maxrec, &lp); break;
    case 6: rc = plan_launch<false, false, false, bk::kStages, true>(ctx, n, maxrec, &lp); break;
    case 7: rc = plan_launch<false, false, true, bk::kStages, true>(ctx, n, maxrec, &lp); break;
    case 8: rc = plan_launch<true, false, false, bk::kStages, true>(ctx, n, maxrec, &lp); break;
    default: rc = plan_launch<true, false, true, bk::kStages, true>(ctx, n, maxrec, &lp); break;
  }
  if (rc != BK_OK) return rc;
  const size_t kmask_off = (scratch_bytes(lp.ntiles) + 15) & ~size_t(15);  // (keep masks behind everything else)
  const size_t sc_bytes = co && co->want_kmask ? kmask_off + (size_t)lp.ntiles * 8 : scratch_bytes(lp.ntiles);
  DevBuf* const scbuf = co && co->scratch ? co->scratch : &s->scratch.buf;
  rc = ensure(ctx, scbuf, sc_bytes);
  if (rc != BK_OK) return rc;
  uint8_t* sc = (uint8_t*)scbuf->p;
  CK(cudaMemsetAsync(sc, 0, sc_bytes, stream));

  P.prog[0] = ctx->prog[0];
  P.prog[1] = ctx->prog[1];
  for (int m = 0; m < (ctx->paired ? 2 : 1); ++m) {
    P.text[m] = in[m]->text;
    P.off[m] = in[m]->rec_off;
    P.text_len[m] = (uint32_t)in[m]->text_len;
    P.in_cap[m] = lp.in_cap[m];
    P.out_cap[m] = lp.out_cap[m];
    P.nvar[m] = (uint32_t)ctx->var[m].size();
    P.vprog[m] = (const bk_devprog*)ctx->vdev[m].p;
  }
  P.out[0] = out1; P.out[1] = out2;
  P.cap[0] = cap1; P.cap[1] = cap2;
  P.match[0] = match1; P.match[1] = match2;
  P.result = (bk::DevResult*)(sc + 64);
  P.tab_in[0] = self_tab;
  P.xtab[0] = self_xtab;
  if (co) {
    P.hlen[0] = co->hlen;
    P.rslot = co->rslot;
    if (co->want_kmask) P.kmask = co->kmask = (unsigned long long*)(sc + kmask_off);
    co->tile_recs = lp.tile_recs;
    if (co->tab_own) { P.tab_in[0] = co->tab_own; P.xtab[0] = co->xtab; }
    P.tab_in[1] = co->tab_other;
    if (co->pair_tables)
      for (int m = 0; m < 2; ++m) { P.tab_in[m] = co->ptab[m]; P.xtab[m] = co->pxtab[m]; }
    P.no_nout = co->no_nout ? 1u : 0u;
    if (co->result) P.result = co->result;
  }
  P.tile_counter = (unsigned int*)sc;
  P.prof = (unsigned long long*)(sc + events_offset(lp.ntiles));  // only written by a BK_PROFILE_PHASES build
  P.desc[0] = (unsigned long long*)(sc + 512);
  P.desc[1] = (unsigned long long*)(sc + block_desc_offset(lp.ntiles));
  P.desc[2] = (unsigned long long*)(sc + block_acc_offset(lp.ntiles));
  P.n = n;
  P.ntiles = lp.ntiles;
  P.tile_recs = lp.tile_recs;
  P.flags = ctx->flags;
  if (n > 0) {
    const int tp = bk::cta_warps(true) * 32, ts = bk::cta_warps(false) * 32;
    switch (kid) {
      case 0: bk::extract_kernel<true, true, true><<<lp.grid, tp, lp.smem, stream>>>(P); break;
      case 1: bk::extract_kernel<false, true, true><<<lp.grid, ts, lp.smem, stream>>>(P); break;
      case 2: bk::extract_kernel<true, true><<<lp.grid, tp, lp.smem, stream>>>(P); break;
      case 3: bk::extract_kernel<true, false><<<lp.grid, tp, lp.smem, stream>>>(P); break;
      case 4: bk::extract_kernel<false, true><<<lp.grid, ts, lp.smem, stream>>>(P); break;
      case 5: bk::extract_kernel<false, false><<<lp.grid, ts, lp.smem, stream>>>(P); break;
      case 6: bk::extract_kernel<false, false, false, bk::kStages, true><<<lp.grid, ts, lp.smem, stream>>>(P); break;
      case 7: bk::extract_kernel<false, false, true, bk::kStages, true><<<lp.grid, ts, lp.smem, stream>>>(P); break;
      case 8: bk::extract_kernel<true, false, false, bk::kStages, true><<<lp.grid, tp, lp.smem, stream>>>(P); break;
      default: bk::extract_kernel<true, false, true, bk::kStages, true><<<lp.grid, tp, lp.smem, stream>>>(P); break;
    }
    if (merr) bk::merge_error_kernel<<<1, 1, 0, stream>>>(P.result, merr);
    CK(cudaGetLastError());
    if (launches) *launches += 1;
  }
  return BK_OK;
}

// match_kernel for one mate: its match table (start | variant << 24, or -1) into `table`
// -> s->tab_match[ti] (and s->tab_extras[ti] for a parametric program)
int enqueue_match(bk_ctx* ctx, int m, Slot* s, int ti, cudaStream_t stream, uint32_t n, const bk_mate_in* in,
                  bk::DevResult* result, uint32_t* launches) {
  int rce;
  if ((rce = ensure(ctx, &s->tab_match[ti], ((size_t)n + 1) * 4)) != BK_OK) return rce;
  if (bk_free_shape(ctx->prog[m]) && (rce = ensure(ctx, &s->tab_extras[ti], ((size_t)n + 1) * 16)) != BK_OK) return rce;
  if (n == 0) return BK_OK;
  const uint32_t maxrec = in->max_rec_bytes;
  // 32 reads per warp and two CTAs per SM while the records allow it; long records: fewer reads per tile, then fewer
  // warps per CTA (one warp may take nearly all of the SM's shared memory)
  const size_t half = (233472 / 2 - 2048) & ~size_t(127);
  uint32_t R = bk::kTileRecs;
  int W = bk::kMatchWarps;
  auto cap_of = [&](uint32_t r) { return (uint32_t)(((size_t)r * maxrec + 48 + 15) & ~size_t(15)); };
  while (R > 8 && (size_t)W * cap_of(R) + 64 > half) --R;
  while (R > 1 && (size_t)W * cap_of(R) + 64 > kMaxDynSmem) --R;
  while (W > 1 && (size_t)W * cap_of(R) + 64 > kMaxDynSmem) --W;
  if ((size_t)W * cap_of(R) + 64 > kMaxDynSmem) {
    ctx->last_error = "record too long for the on-chip staging buffer";
    return BK_E_CAPACITY;
  }
  bk::MatchParams M;
  memset(&M, 0, sizeof M);
  M.prog = ctx->prog[m];
  M.text = in->text;
  M.off = in->rec_off;
  M.match = (int32_t*)s->tab_match[ti].p;
  M.spans = (uint4*)s->tab_extras[ti].p;
  M.result = result;
  M.nvar = (uint32_t)ctx->var[m].size();
  M.vprog = (const bk_devprog*)ctx->vdev[m].p;
  M.n = n;
  M.tile_recs = R;
  M.ntiles = (n + R - 1) / R;
  M.cap = cap_of(R);
  M.rc = (ctx->flags & BK_FLAG_RC) ? 1u : 0u;
  M.vmcode = (const uint2*)ctx->vmdev[m].p;
  const bool vm = ctx->prog[m].vm_len > 0;
  const auto kernel = vm ? bk::match_kernel<true> : bk::match_kernel<false>;
  const size_t smem = (size_t)W * M.cap + 64;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, W * 32, smem));
  if (per_sm < 1) per_sm = 1;
  long long grid = std::min<long long>((long long)per_sm * ctx->sm_count, ((long long)M.ntiles + W - 1) / W);
  if (ctx->flags & BK_FLAG_VERBOSE)
    fprintf(stderr, "[bk] match_kernel mate %d: tile_recs %u cap %u warps %d ctas/sm %d grid %lld\n", m + 1, R, M.cap, W, per_sm, grid);
  kernel<<<(int)std::max<long long>(grid, 1), W * 32, smem, stream>>>(M);
  CK(cudaGetLastError());
  if (launches) *launches += 1;
  return BK_OK;
}

// A paired batch with two patterns of which at least one needs a search (not anchored, -r, variable repetition), as
// separate launches on one stream: match_kernel on every such mate, then one single-end extract launch per mate.
// A mate whose pattern is a plain anchored one is matched by its own extract launch, which runs first and leaves its
// table for the other's keep rule (run.rs:149-160); a matched mate's launch takes its matches from the table.
// Against the paired general kernel this reads a searched mate's text twice (once here, once staged for the emit),
// and gains the single-end kernel's four CTAs per SM and an emit path without any search code next to it.
int enqueue_tabled(bk_ctx* ctx, Slot* s, cudaStream_t stream, uint32_t n, const bk_mate_in* in1, const bk_mate_in* in2,
                   uint8_t* out1, uint64_t cap1, uint8_t* out2, uint64_t cap2, int32_t* match1, int32_t* match2,
                   uint32_t* launches) {
  const bk_mate_in* in[2] = {in1, in2};
  uint8_t* const out[2] = {out1, out2};
  const uint64_t cap[2] = {cap1, cap2};
  int32_t* const user_match[2] = {match1, match2};
  for (int m = 0; m < 2; ++m) {
    if (!in[m] || !in[m]->text || !in[m]->rec_off) { ctx->last_error = "null batch pointer"; return BK_E_ARG; }
    if (in[m]->text_len >= (1ull << 32)) { ctx->last_error = "batch text must be < 4 GiB per mate"; return BK_E_CAPACITY; }
    if (((uintptr_t)in[m]->text & 15) != 0) { ctx->last_error = "batch text must be 16-byte aligned"; return BK_E_ARG; }
    if (in[m]->max_rec_bytes == 0) { ctx->last_error = "max_rec_bytes must be set"; return BK_E_ARG; }
  }
  int rc;
  for (int m = 0; m < 2; ++m)
    if ((rc = ensure(ctx, &s->tab_match[m], ((size_t)n + 1) * 4)) != BK_OK) return rc;
  if ((rc = ensure(ctx, &s->tab_err, 64)) != BK_OK) return rc;
  for (int m = 0; m < 2; ++m)  // (before the table pointers are taken: enqueue_match only finds them large enough)
    if (bk_free_shape(ctx->prog[m]) && (rc = ensure(ctx, &s->tab_extras[m], ((size_t)n + 1) * 16)) != BK_OK) return rc;
  const bool hard[2] = {is_hard(ctx, 0), is_hard(ctx, 1)};
  const int first = hard[0] && !hard[1] ? 1 : 0;  // the mate whose extract launch runs first
  int32_t* const tab[2] = {(int32_t*)s->tab_match[0].p, (int32_t*)s->tab_match[1].p};
  // 1. the matchers.  They run before the launch that resets the batch's result, so a tile that does not fit their
  // buffers is flagged in a result of their own and merged at the end.
  bk::DevResult* const merr = (bk::DevResult*)s->tab_err.p;
  CK(cudaMemsetAsync(merr, 0, sizeof(bk::DevResult), stream));
  for (int m = 0; m < 2; ++m)
    if (hard[m] && (rc = enqueue_match(ctx, m, s, m, stream, n, in[m], merr, launches)) != BK_OK) return rc;
  // 2a. (default) ONE paired launch in table form: both mates staged, a searched mate's front warp reads its table
  if (!(ctx->flags & BK_FLAG_SPLIT_TWO_PATTERNS)) {
    CoLaunch co;
    co.pair_tables = true;
    for (int m = 0; m < 2; ++m)
      if (hard[m]) {
        co.ptab[m] = tab[m];
        co.pxtab[m] = (const uint4*)s->tab_extras[m].p;
      }
    // the kernel writes a mate's matches (variant bits stripped) where the caller asked for them
    rc = enqueue_extract(ctx, s, stream, n, in1, in2, out1, cap1, out2, cap2, match1, match2, launches, &co);
    if (rc == BK_OK) {
      if (n) {
        bk::merge_error_kernel<<<1, 1, 0, stream>>>((bk::DevResult*)((uint8_t*)s->scratch.buf.p + 64), merr);
        CK(cudaGetLastError());
      }
      return BK_OK;
    }
    // records too long to stage both mates at once: the single-end launches below need half the shared memory
    if (rc != BK_E_CAPACITY) return rc;
  }
  // 2b. one single-end extract launch per mate
  for (int k = 0; k < 2; ++k) {
    const int m = k == 0 ? first : 1 - first;
    CoLaunch co;
    co.rslot = (uint32_t)m;
    co.tab_other = tab[1 - m];
    co.tab_own = hard[m] ? tab[m] : nullptr;
    co.xtab = (const uint4*)s->tab_extras[m].p;
    if (k == 1) {
      co.scratch = &s->split_scratch;
      co.result = (bk::DevResult*)((uint8_t*)s->scratch.buf.p + 64);  // the first launch's, which has reset it
      co.no_nout = true;
    }
    // a mate that matches for itself leaves its table for the other's keep rule
    int32_t* const mout = hard[m] ? user_match[m] : tab[m];
    rc = enqueue_extract(ctx->view[m], s, stream, n, in[m], nullptr, out[m], cap[m], nullptr, 0, mout, nullptr, launches, &co);
    if (rc != BK_OK) { ctx->last_error = ctx->view[m]->last_error; return rc; }
    if (!hard[m] && user_match[m] && n) CK(cudaMemcpyAsync(user_match[m], tab[m], (size_t)n * 4, cudaMemcpyDeviceToDevice, stream));
  }
  if (n) {
    bk::merge_error_kernel<<<1, 1, 0, stream>>>((bk::DevResult*)((uint8_t*)s->scratch.buf.p + 64), merr);
    CK(cudaGetLastError());
  }
  return BK_OK;
}

// After the stream has been synchronised: the optimistic compaction met a record in foreign framing (bit 2) -- run
// that compaction again with sizes taken from the text, and stay careful for this context's later batches.
int compact_redo(bk_ctx* ctx, Slot& s) {
  if (!(s.h_res->error & 4u) || !s.cpar_valid) return BK_OK;
  ctx->compact_careful = true;
  const bk::CompactParams& C = s.cpar;
  CK(cudaMemsetAsync(s.split_scratch.p, 0, scratch_bytes(C.ntiles), s.stream));
  CK(cudaMemsetAsync(&C.result->error, 0, sizeof(unsigned int), s.stream));
  bk::compact_kernel<true><<<C.ntiles, bk::kCompactThreads, 0, s.stream>>>(C);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(s.h_res, C.result, sizeof(bk::DevResult), cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return BK_OK;
}

int check_dev_result(bk_ctx* ctx, const bk::DevResult& r) {
  if (r.error & 2u) { ctx->last_error = "a record is longer than max_rec_bytes"; return BK_E_ARG; }
  if (r.error & 1u) { ctx->last_error = "output buffer too small"; return BK_E_CAPACITY; }
  if (r.error & 8u) {  // match_vm gave up on a read
    ctx->last_error = "unsupported pattern construct (a read needs more backtracking states for this pattern than the device matcher keeps)";
    return BK_E_UNSUPPORTED;
  }
  return BK_OK;
}

void fill_result(bk_result* res, const bk::DevResult& r, uint32_t n, float ms, uint32_t launches) {
  if (!res) return;
  res->out1_len = r.out_len[0];
  res->out2_len = r.out_len[1];
  res->n_in = n;
  res->n_out = r.n_out;
  res->n_match1 = r.n_match[0];
  res->n_match2 = r.n_match[1];
  res->kernel_ms = ms;
  res->launches = launches;
  res->h2d_bytes = res->d2h_bytes = 0;
}

}  // namespace

extern "C" {

int bk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int bk_warmup(int device) {
  if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) {
    cudaGetLastError();
    return BK_E_CUDA;
  }
  return BK_OK;
}

bk_ctx* bk_ctx_create(int device, const bk_program* p1, const bk_program* p2, uint32_t flags, int* status,
                      char* err, size_t errcap) {
  auto fail = [&](int st, const std::string& m) -> bk_ctx* {
    if (status) *status = st;
    if (err && errcap) {
      size_t k = std::min(m.size(), errcap - 1);
      memcpy(err, m.data(), k);
      err[k] = 0;
    }
    return nullptr;
  };
  if (!p1 && !p2) return fail(BK_E_ARG, "at least one pattern is required");
  if (!(flags & BK_FLAG_PAIRED) && (p2 || !p1))
    return fail(BK_E_ARG, "single-end mode takes pattern1 only (run.rs:43)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(BK_E_CUDA, std::string("no CUDA device available (") +
                               (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                               "); this library has no CPU path");
  if (device < 0 || device >= ndev) return fail(BK_E_ARG, "device index out of range");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(BK_E_CUDA, cudaGetErrorString(e));
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(BK_E_CUDA, cudaGetErrorString(e));
  if (prop.major < 10) return fail(BK_E_CUDA, "device is not sm_100 (Blackwell B200) class");
  bk_ctx* ctx = new (std::nothrow) bk_ctx();
  if (!ctx) return fail(BK_E_CAPACITY, "out of memory");
  ctx->device = device;
  ctx->flags = flags;
  ctx->paired = (flags & BK_FLAG_PAIRED) != 0;
  ctx->sm_count = prop.multiProcessorCount;
  memset(ctx->prog, 0, sizeof ctx->prog);
  const bk_program* pp[2] = {p1, p2};
  for (int m = 0; m < 2; ++m) {
    if (!pp[m]) continue;
    ctx->prog[m] = pp[m]->prog.dev;
    ctx->var[m] = pp[m]->prog.variants;
    if (ctx->var[m].empty()) ctx->var[m].push_back(pp[m]->prog.dev);
    for (const bk_devprog& v : ctx->var[m]) ctx->growth[m] = std::max(ctx->growth[m], v.hdr_growth);
    if (ctx->var[m].size() > 1) {
      const size_t bytes = ctx->var[m].size() * sizeof(bk_devprog);
      if (cudaMalloc(&ctx->vdev[m].p, bytes) != cudaSuccess ||
          cudaMemcpy(ctx->vdev[m].p, ctx->var[m].data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        std::string msg = std::string("CUDA error uploading the pattern's variants: ") + cudaGetErrorString(cudaGetLastError());
        bk_ctx_destroy(ctx);
        return fail(BK_E_CUDA, msg);
      }
      ctx->vdev[m].cap = bytes;
    }
    if (ctx->prog[m].vm_len) {
      const size_t bytes = pp[m]->prog.vm.size() * sizeof(bk_vmop);
      if (pp[m]->prog.vm.size() != ctx->prog[m].vm_len || cudaMalloc(&ctx->vmdev[m].p, bytes) != cudaSuccess ||
          cudaMemcpy(ctx->vmdev[m].p, pp[m]->prog.vm.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        std::string msg = std::string("CUDA error uploading the pattern's program: ") + cudaGetErrorString(cudaGetLastError());
        bk_ctx_destroy(ctx);
        return fail(BK_E_CUDA, msg);
      }
      ctx->vmdev[m].cap = bytes;
    }
  }
  for (int i = 0; i < kSlots; ++i) {
    Slot& s = ctx->slot[i];
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&s.ev0) != cudaSuccess || cudaEventCreate(&s.ev1) != cudaSuccess ||
        cudaEventCreate(&s.evh0) != cudaSuccess || cudaEventCreate(&s.evh1) != cudaSuccess ||
        cudaHostAlloc((void**)&s.h_res, sizeof(bk::DevResult), cudaHostAllocDefault) != cudaSuccess) {
      std::string m = std::string("CUDA resource creation failed: ") + cudaGetErrorString(cudaGetLastError());
      bk_ctx_destroy(ctx);
      return fail(BK_E_CUDA, m);
    }
  }
  if (ctx->paired && (p1 == nullptr) != (p2 == nullptr)) {
    // One pattern: `solo` is the single-end view of this context for the pattern's mate.  It serves the host
    // pass-through (bk_submit: the other mate never leaves the host) and the two-kernel device form (enqueue_split).
    bk_ctx* solo = new (std::nothrow) bk_ctx();
    if (!solo) { bk_ctx_destroy(ctx); return fail(BK_E_CAPACITY, "out of memory"); }
    const int pm = p1 ? 0 : 1;
    solo->device = device;
    solo->flags = flags & ~(uint32_t)(BK_FLAG_PAIRED | BK_FLAG_DEVICE_BOTH_MATES | BK_FLAG_HOST_MATE_PINNED | BK_FLAG_STAGE_BOTH_MATES);
    solo->paired = false;
    solo->sm_count = ctx->sm_count;
    memset(solo->prog, 0, sizeof solo->prog);
    solo->prog[0] = ctx->prog[pm];
    solo->var[0] = ctx->var[pm];
    solo->vdev[0] = ctx->vdev[pm];  // (shared, owned by ctx)
    solo->vmdev[0] = ctx->vmdev[pm];
    solo->growth[0] = ctx->growth[pm];
    ctx->solo = solo;
    if (!(flags & BK_FLAG_DEVICE_BOTH_MATES)) ctx->pass_mate = 1 - pm;
    ctx->pass_pinned = (flags & BK_FLAG_HOST_MATE_PINNED) != 0;
  }
  if (ctx->paired && p1 && p2) {
    for (int m = 0; m < 2; ++m) {  // single-end views of the two mates (enqueue_tabled)
      bk_ctx* v = new (std::nothrow) bk_ctx();
      if (!v) { bk_ctx_destroy(ctx); return fail(BK_E_CAPACITY, "out of memory"); }
      v->device = device;
      v->flags = flags & ~(uint32_t)(BK_FLAG_PAIRED | BK_FLAG_DEVICE_BOTH_MATES | BK_FLAG_HOST_MATE_PINNED | BK_FLAG_STAGE_BOTH_MATES);
      v->paired = false;
      v->sm_count = ctx->sm_count;
      memset(v->prog, 0, sizeof v->prog);
      v->prog[0] = ctx->prog[m];
      v->var[0] = ctx->var[m];
      v->vdev[0] = ctx->vdev[m];  // (shared, owned by ctx)
      v->vmdev[0] = ctx->vmdev[m];
      v->growth[0] = ctx->growth[m];
      ctx->view[m] = v;
    }
  }
  ctx->host_threads = std::max(1, std::min((int)std::thread::hardware_concurrency(), 32));
  if (status) *status = BK_OK;
  if (err && errcap) err[0] = 0;
  return ctx;
}

void bk_ctx_destroy(bk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < kSlots; ++i) {
    Slot& s = ctx->slot[i];
    if (s.stream) cudaStreamSynchronize(s.stream);
    for (int m = 0; m < 2; ++m) {
      if (s.text[m].p) cudaFree(s.text[m].p);
      if (s.off[m].p) cudaFree(s.off[m].p);
      if (s.out[m].p) cudaFree(s.out[m].p);
    }
    if (s.scratch.buf.p) cudaFree(s.scratch.buf.p);
    if (s.match.p) cudaFree(s.match.p);
    if (s.h_match) cudaFreeHost(s.h_match);
    if (s.h_tok) cudaFreeHost(s.h_tok);
    if (s.split_match.p) cudaFree(s.split_match.p);
    if (s.split_hlen.p) cudaFree(s.split_hlen.p);
    if (s.split_scratch.p) cudaFree(s.split_scratch.p);
    if (s.tab_err.p) cudaFree(s.tab_err.p);
    for (int m = 0; m < 2; ++m) {
      if (s.tab_match[m].p) cudaFree(s.tab_match[m].p);
      if (s.tab_extras[m].p) cudaFree(s.tab_extras[m].p);
    }
    for (int m = 0; m < 2; ++m)
      if (s.tok[m].p) cudaFree(s.tok[m].p);
    if (s.h_res) cudaFreeHost(s.h_res);
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    if (s.evh0) cudaEventDestroy(s.evh0);
    if (s.evh1) cudaEventDestroy(s.evh1);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (ctx->tok.p) cudaFree(ctx->tok.p);
  for (int m = 0; m < 2; ++m) {
    if (ctx->vdev[m].p) cudaFree(ctx->vdev[m].p);
    if (ctx->vmdev[m].p) cudaFree(ctx->vmdev[m].p);
  }
  delete ctx->solo;
  delete ctx->view[0];
  delete ctx->view[1];
  delete ctx;
}

const char* bk_last_error(const bk_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

// Pinned host memory.  Anonymous memory on transparent huge pages, touched and then registered with CUDA, comes up
// 2-3 times faster than cudaHostAlloc (measured on the B200 hosts: 0.04 s against 0.10 s per 256 MiB, and five times
// faster to release), because the driver pins 2 MiB pages instead of 4 KiB ones; a file -> file run allocates several
// hundred MiB of batch buffers before its first batch moves.  cudaHostAlloc is the fallback.
namespace {
std::mutex g_host_mu;
std::map<void*, size_t> g_host_maps;  // registered mappings: base -> length
}  // namespace

void* bk_host_alloc(size_t bytes) {
  if (!bytes) bytes = 1;
  constexpr size_t kHuge = 2u << 20;
  if (bytes >= 4 * kHuge) {
    const size_t len = (bytes + kHuge - 1) & ~(kHuge - 1);
    void* raw = mmap(nullptr, len + kHuge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (raw != MAP_FAILED) {
      // keep the 2 MiB-aligned part, give the rest back
      uint8_t* base = (uint8_t*)(((uintptr_t)raw + kHuge - 1) & ~(uintptr_t)(kHuge - 1));
      if (base > (uint8_t*)raw) munmap(raw, (size_t)(base - (uint8_t*)raw));
      uint8_t* end = (uint8_t*)raw + len + kHuge;
      if (end > base + len) munmap(base + len, (size_t)(end - (base + len)));
      madvise(base, len, MADV_HUGEPAGE);
      for (size_t o = 0; o < len; o += 4096) base[o] = 0;  // fault the pages in (huge ones where the kernel gives them)
      if (cudaHostRegister(base, len, cudaHostRegisterDefault) == cudaSuccess) {
        std::lock_guard<std::mutex> g(g_host_mu);
        g_host_maps[base] = len;
        return base;
      }
      cudaGetLastError();
      munmap(base, len);
    }
  }
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void bk_host_free(void* p) {
  if (!p) return;
  size_t len = 0;
  {
    std::lock_guard<std::mutex> g(g_host_mu);
    auto it = g_host_maps.find(p);
    if (it != g_host_maps.end()) {
      len = it->second;
      g_host_maps.erase(it);
    }
  }
  if (len) {
    cudaHostUnregister(p);
    munmap(p, len);
  } else {
    cudaFreeHost(p);
  }
}

int bk_ctx_nslots(const bk_ctx* ctx) { return ctx ? kSlots : 0; }

int bk_ctx_set_host_threads(bk_ctx* ctx, int n) {
  if (!ctx || n < 1) return BK_E_ARG;
  ctx->host_threads = std::min(n, 64);
  return BK_OK;
}

uint64_t bk_out_bound(const bk_ctx* ctx, int mate, uint32_t n, uint64_t text_len) {
  if (!ctx || mate < 1 || mate > 2) return 0;
  const bk_devprog& pg = ctx->prog[mate - 1];
  // out record <= in record + 1 (a final record without EOL gains one) + header growth
  // (a parametric program: its captures may also grow by the read's own bytes, par_growth summed over the records)
  return text_len + (uint64_t)n * ((pg.present ? ctx->growth[mate - 1] : 0) + 1) + 16 + (pg.present ? par_growth_total(pg, n, text_len) : 0);
}

int bk_submit(bk_ctx* ctx, int slot, uint32_t n, const bk_mate_in* in1, const bk_mate_in* in2) {
  if (!ctx || slot < 0 || slot >= kSlots || !in1) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[slot];
  if (s.busy) { ctx->last_error = "slot is busy: call bk_wait first"; return BK_E_ARG; }
  const bk_mate_in* in[2] = {in1, in2};
  const int nm = ctx->paired ? 2 : 1;
  if (ctx->paired != (in2 != nullptr)) { ctx->last_error = "mate count does not match the context mode"; return BK_E_ARG; }
  bk_mate_in dev[2];
  for (int m = 0; m < nm; ++m) {
    if (!in[m]->text || !in[m]->rec_off) { ctx->last_error = "null batch pointer"; return BK_E_ARG; }
    if (in[m]->text_len >= (1ull << 32)) { ctx->last_error = "batch text must be < 4 GiB per mate"; return BK_E_CAPACITY; }
  }
  s.raw = false;
  s.passthrough = ctx->pass_mate >= 0;
  s.pt = ctx->pass_mate;
  if (s.passthrough) {
    // Exactly one mate has a pattern.  The other one is only ever copied or dropped (run.rs:149-160), and the
    // caller's buffer already holds it: send the pattern's mate alone, as a single-end batch (an unmatched
    // read of it drops the pair either way, run.rs:71-78), fetch its match table with the output, and let
    // bk_wait assemble the pattern-less mate's output from the host copy (bk_fastq.cc, gather_kept).
    const int pm = 1 - ctx->pass_mate;
    bk_ctx* solo = ctx->solo;
    int rc;
    if ((rc = ensure(ctx, &s.text[pm], in[pm]->text_len + 64)) != BK_OK) return rc;
    if ((rc = ensure(ctx, &s.off[pm], ((size_t)n + 1) * 4)) != BK_OK) return rc;
    const uint64_t ob = bk_out_bound(ctx, pm + 1, n, in[pm]->text_len);
    if ((rc = ensure(ctx, &s.out[pm], ob)) != BK_OK) return rc;
    if ((rc = ensure(ctx, &s.match, ((size_t)n + 1) * 4)) != BK_OK) return rc;
    if (s.h_match_cap < (size_t)n + 1) {
      if (s.h_match) CK(cudaFreeHost(s.h_match));
      s.h_match = nullptr;
      s.h_match_cap = 0;
      const size_t want = (size_t)n + 1 + n / 8;
      CK(cudaHostAlloc((void**)&s.h_match, want * 4, cudaHostAllocDefault));
      s.h_match_cap = want;
    }
    s.out_cap[pm] = ob;
    s.h2d_bytes = in[pm]->text_len + ((uint64_t)n + 1) * 4;
    s.host_pt = *in[ctx->pass_mate];
    CK(cudaEventRecord(s.evh0, s.stream));
    CK(cudaMemcpyAsync(s.text[pm].p, in[pm]->text, in[pm]->text_len, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.off[pm].p, in[pm]->rec_off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s.stream));
    CK(cudaEventRecord(s.evh1, s.stream));
    bk_mate_in d = *in[pm];
    d.text = (const uint8_t*)s.text[pm].p;
    d.rec_off = (const uint32_t*)s.off[pm].p;
    s.launches = 0;
    CK(cudaEventRecord(s.ev0, s.stream));
    rc = enqueue_extract(solo, &s, s.stream, n, &d, nullptr, (uint8_t*)s.out[pm].p, ob, nullptr, 0,
                         (int32_t*)s.match.p, nullptr, &s.launches);
    if (rc != BK_OK) { ctx->last_error = solo->last_error; return rc; }
    CK(cudaEventRecord(s.ev1, s.stream));
    CK(cudaMemcpyAsync(s.h_res, (uint8_t*)s.scratch.buf.p + 64, sizeof(bk::DevResult), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(s.h_match, s.match.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s.stream));
    s.n = n;
    s.busy = true;
    return BK_OK;
  }
  s.h2d_bytes = 0;
  for (int m = 0; m < nm; ++m) {
    int rc;
    s.h2d_bytes += in[m]->text_len + ((uint64_t)n + 1) * 4;
    if ((rc = ensure(ctx, &s.text[m], in[m]->text_len + 64)) != BK_OK) return rc;
    if ((rc = ensure(ctx, &s.off[m], ((size_t)n + 1) * 4)) != BK_OK) return rc;
    uint64_t ob = bk_out_bound(ctx, m + 1, n, in[m]->text_len);
    if ((rc = ensure(ctx, &s.out[m], ob)) != BK_OK) return rc;
    s.out_cap[m] = ob;
    CK(cudaMemcpyAsync(s.text[m].p, in[m]->text, in[m]->text_len, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.off[m].p, in[m]->rec_off, ((size_t)n + 1) * 4, cudaMemcpyHostToDevice, s.stream));
    dev[m] = *in[m];
    dev[m].text = (const uint8_t*)s.text[m].p;
    dev[m].rec_off = (const uint32_t*)s.off[m].p;
  }
  s.launches = 0;
  CK(cudaEventRecord(s.ev0, s.stream));
  int rc = enqueue_extract(ctx, &s, s.stream, n, &dev[0], ctx->paired ? &dev[1] : nullptr,
                           (uint8_t*)s.out[0].p, s.out_cap[0], (uint8_t*)s.out[1].p, s.out_cap[1], nullptr,
                           nullptr, &s.launches);
  if (rc != BK_OK) return rc;
  CK(cudaEventRecord(s.ev1, s.stream));
  CK(cudaMemcpyAsync(s.h_res, (uint8_t*)s.scratch.buf.p + 64, sizeof(bk::DevResult), cudaMemcpyDeviceToHost,
                     s.stream));
  s.n = n;
  s.busy = true;
  return BK_OK;
}

namespace {

// enqueue the device tokenizer for one mate's text (already on the device): rec_off[0..n], TokResult in tok->p
int enqueue_tokenize(bk_ctx* ctx, cudaStream_t st, const uint8_t* d_text, uint64_t len, bool no_eol, uint32_t n,
                     uint32_t* d_rec_off, DevBuf* tok) {
  const uint32_t nblocks = (uint32_t)((len + bk::kTokChunk - 1) / bk::kTokChunk);
  int rc = ensure(ctx, tok, bk::tok_desc_words(nblocks) * 8 + 128);
  if (rc != BK_OK) return rc;
  bk::TokResult* d_res = (bk::TokResult*)tok->p;
  unsigned int* ticket = (unsigned int*)((uint8_t*)tok->p + 64);
  unsigned long long* desc = (unsigned long long*)((uint8_t*)tok->p + 128);
  bk::tok_init_kernel<<<(unsigned)((bk::tok_desc_words(nblocks) + 255) / 256), 256, 0, st>>>(d_res, desc, nblocks, ticket);
  bk::tok_fused_kernel<<<nblocks, bk::kTokThreads, 0, st>>>(d_text, (uint32_t)len, d_rec_off, n, desc, ticket, nblocks, d_res);
  if (no_eol) {  // the text's last record ends with the text: its end offset and its check come from here
    bk::tok_set_end_kernel<<<1, 32, 0, st>>>(d_rec_off, n, (uint32_t)len);
    bk::tok_check_kernel<<<1, 32, 0, st>>>(d_text, d_rec_off, n - 1, n, (uint32_t)len, d_res);
  }
  CK(cudaGetLastError());
  return BK_OK;
}

int enqueue_raw_extract(bk_ctx* ctx, Slot& s) {
  bk_mate_in dev[2];
  const int nm = ctx->paired ? 2 : 1;
  for (int m = 0; m < nm; ++m) {
    dev[m].text = (const uint8_t*)s.text[m].p;
    dev[m].rec_off = (const uint32_t*)s.off[m].p;
    dev[m].text_len = s.raw_len[m];
    dev[m].max_rec_bytes = s.raw_maxrec[m];
    dev[m].reserved = 0;
  }
  CK(cudaEventRecord(s.ev0, s.stream));
  int rc = enqueue_extract(ctx, &s, s.stream, s.n, &dev[0], ctx->paired ? &dev[1] : nullptr, (uint8_t*)s.out[0].p,
                           s.out_cap[0], (uint8_t*)s.out[1].p, s.out_cap[1], nullptr, nullptr, &s.launches);
  if (rc != BK_OK) return rc;
  CK(cudaEventRecord(s.ev1, s.stream));
  CK(cudaMemcpyAsync(s.h_res, (uint8_t*)s.scratch.buf.p + 64, sizeof(bk::DevResult), cudaMemcpyDeviceToHost, s.stream));
  return BK_OK;
}

}  // namespace

int bk_submit_text(bk_ctx* ctx, int slot, uint32_t n, const uint8_t* text1, uint64_t len1, const uint8_t* text2,
                   uint64_t len2, uint32_t max_rec_hint) {
  if (!ctx || slot < 0 || slot >= kSlots || !text1) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[slot];
  if (s.busy) { ctx->last_error = "slot is busy: call bk_wait first"; return BK_E_ARG; }
  if (ctx->paired != (text2 != nullptr)) { ctx->last_error = "mate count does not match the context mode"; return BK_E_ARG; }
  if (n == 0) { ctx->last_error = "empty batch"; return BK_E_ARG; }
  const uint8_t* text[2] = {text1, text2};
  const uint64_t len[2] = {len1, len2};
  const int nm = ctx->paired ? 2 : 1;
  if (!s.h_tok) CK(cudaHostAlloc((void**)&s.h_tok, 2 * sizeof(bk::TokResult), cudaHostAllocDefault));
  s.raw = true;
  s.passthrough = false;
  s.h2d_bytes = 0;
  s.launches = 0;
  bool need_sync = false;
  for (int m = 0; m < nm; ++m) {
    if (len[m] == 0 || len[m] >= (1ull << 32) - 65536) { ctx->last_error = "batch text must be 1 byte .. 4 GiB - 64 KiB per mate"; return BK_E_CAPACITY; }
    int rc;
    if ((rc = ensure(ctx, &s.text[m], len[m] + 64)) != BK_OK) return rc;
    if ((rc = ensure(ctx, &s.off[m], ((size_t)n + 2) * 4)) != BK_OK) return rc;
    const uint64_t ob = bk_out_bound(ctx, m + 1, n, len[m]);
    if ((rc = ensure(ctx, &s.out[m], ob)) != BK_OK) return rc;
    s.out_cap[m] = ob;
    s.raw_len[m] = len[m];
    s.raw_no_eol[m] = text[m][len[m] - 1] != '\n';
    s.h2d_bytes += len[m];
    CK(cudaMemcpyAsync(s.text[m].p, text[m], len[m], cudaMemcpyHostToDevice, s.stream));
    if ((rc = enqueue_tokenize(ctx, s.stream, (const uint8_t*)s.text[m].p, len[m], s.raw_no_eol[m], n, (uint32_t*)s.off[m].p,
                               &s.tok[m])) != BK_OK) return rc;
    // (The tokenizer's result is fetched at the END of the slot's work, not here: a device -> host copy between the
    // two mates' uploads queues behind the other slot's output download on the copy engine, and the second upload,
    // ordered behind it in this stream, then waits for all of that: 36 instead of 28 ms per 2 M pairs.)
    // The launch plan (tile size) needs the longest record before the tokenizer has answered: the caller's hint, else
    // what earlier batches of this context showed (with headroom).  A batch whose records turn out longer is launched
    // again by bk_wait with the exact figure; the first batch of a context without a hint waits for the tokenizer here.
    uint32_t mr = std::max(max_rec_hint, ctx->maxrec_seen[m] + ctx->maxrec_seen[m] / 8);
    if (mr == 0) need_sync = true;
    s.raw_maxrec[m] = mr;
  }
  s.n = n;
  if (need_sync) {
    for (int m = 0; m < nm; ++m)
      CK(cudaMemcpyAsync(&s.h_tok[m], s.tok[m].p, sizeof(bk::TokResult), cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    for (int m = 0; m < nm; ++m) {
      s.raw_maxrec[m] = std::max(s.raw_maxrec[m], std::max(s.h_tok[m].max_rec_bytes, 1u));
      ctx->maxrec_seen[m] = std::max(ctx->maxrec_seen[m], s.h_tok[m].max_rec_bytes);
    }
  }
  int rc = enqueue_raw_extract(ctx, s);
  if (rc != BK_OK) return rc;
  for (int m = 0; m < nm; ++m)
    CK(cudaMemcpyAsync(&s.h_tok[m], s.tok[m].p, sizeof(bk::TokResult), cudaMemcpyDeviceToHost, s.stream));
  s.busy = true;
  return BK_OK;
}

int bk_wait(bk_ctx* ctx, int slot, uint8_t* out1, uint64_t cap1, uint8_t* out2, uint64_t cap2, bk_result* res) {
  if (!ctx || slot < 0 || slot >= kSlots) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[slot];
  if (!s.busy) { ctx->last_error = "slot has no batch in flight"; return BK_E_ARG; }
  // On every path out of here -- errors included -- whatever is still enqueued on the slot's stream (the
  // output copies below read the slot's device buffers) has finished before the slot counts as free again.
  struct Drain {
    Slot& s;
    ~Drain() { cudaStreamSynchronize(s.stream); s.busy = false; }
  } drain{s};
  CK(cudaStreamSynchronize(s.stream));
  if (s.raw) {
    // what the device tokenizer found: every mate must hold exactly n well-formed records
    bool relaunch = false;
    for (int m = 0; m < (ctx->paired ? 2 : 1); ++m) {
      const bk::TokResult t = s.h_tok[m];
      const unsigned long long want = 4ull * s.n - (s.raw_no_eol[m] ? 1 : 0);
      if (t.newlines != want || t.first_bad != 0xFFFFFFFFu) {
        ctx->last_error = "malformed FASTQ text";
        return BK_E_FORMAT;
      }
      ctx->maxrec_seen[m] = std::max(ctx->maxrec_seen[m], t.max_rec_bytes);
      if (t.max_rec_bytes > s.raw_maxrec[m]) { s.raw_maxrec[m] = t.max_rec_bytes; relaunch = true; }
    }
    if (relaunch && (s.h_res->error & 2u)) {  // tiles did not fit the staging ring that the hint sized: once more, exactly
      int rc2 = enqueue_raw_extract(ctx, s);
      if (rc2 != BK_OK) return rc2;
      CK(cudaStreamSynchronize(s.stream));
    }
  }
  {
    const int rcc = compact_redo(ctx, s);
    if (rcc != BK_OK) return rcc;
  }
  bk::DevResult r = *s.h_res;
  int rc = check_dev_result(ctx, r);
  if (rc != BK_OK) return rc;
  if (s.passthrough) {
    const int pt = s.pt, pm = 1 - pt;
    uint8_t* const out[2] = {out1, out2};
    const uint64_t cap[2] = {cap1, cap2};
    if (!out1 || !out2) { ctx->last_error = "out1 and out2 are required in paired mode"; return BK_E_ARG; }
    if (r.out_len[0] > cap[pm]) { ctx->last_error = "output buffer too small"; return BK_E_CAPACITY; }
    // the device's output of the pattern's mate comes back while the host assembles the other mate's
    if (r.out_len[0]) CK(cudaMemcpyAsync(out[pm], s.out[pm].p, r.out_len[0], cudaMemcpyDeviceToHost, s.stream));
    timespec ta, tb;
    clock_gettime(CLOCK_MONOTONIC, &ta);
    const uint64_t got = bk::gather_kept(s.host_pt.text, s.host_pt.rec_off, s.n, s.h_match, out[pt], cap[pt], ctx->host_threads);
    clock_gettime(CLOCK_MONOTONIC, &tb);
    const double gather_ms = (tb.tv_sec - ta.tv_sec) * 1e3 + (tb.tv_nsec - ta.tv_nsec) / 1e6;
    float h2d_ms = 0;
    CK(cudaEventElapsedTime(&h2d_ms, s.evh0, s.evh1));
    // what the bus would have needed for this mate's text, at the rate the other mate's just moved
    const double bus_ms = h2d_ms * (double)s.host_pt.text_len / (double)std::max<uint64_t>(s.h2d_bytes, 1);
    if (!ctx->pass_pinned && s.n >= kPassMinRecords) {
      ctx->pass_slow = gather_ms > h2d_ms + bus_ms ? ctx->pass_slow + 1 : 0;
      if (ctx->pass_slow >= 2) ctx->pass_mate = -1;  // (slots in flight finish as they were submitted)
    }
    if (ctx->flags & BK_FLAG_VERBOSE)
      fprintf(stderr, "[bk] host gather: %u records, %llu bytes, %d threads, %.2f ms; bus: %.2f ms for the other mate, "
              "%.2f ms more for this one%s\n", s.n, (unsigned long long)got, ctx->host_threads, gather_ms, h2d_ms, bus_ms,
              ctx->pass_mate < 0 ? " -> both mates go to the device from now on" : "");
    CK(cudaStreamSynchronize(s.stream));
    if (got == UINT64_MAX) { ctx->last_error = "output buffer too small"; return BK_E_CAPACITY; }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
    bk::DevResult rr = r;  // single-end result -> this context's mates
    rr.out_len[pm] = r.out_len[0];
    rr.out_len[pt] = got;
    rr.n_match[pm] = r.n_match[0];
    rr.n_match[pt] = 0;
    fill_result(res, rr, s.n, ms, s.launches);
    if (res) {
      res->h2d_bytes = s.h2d_bytes;
      res->d2h_bytes = sizeof(bk::DevResult) + (uint64_t)s.n * 4 + r.out_len[0];
    }
    return BK_OK;
  }
  if (r.out_len[0] > cap1 || (ctx->paired && r.out_len[1] > cap2)) {
    ctx->last_error = "output buffer too small";
    return BK_E_CAPACITY;
  }
  if (r.out_len[0]) CK(cudaMemcpyAsync(out1, s.out[0].p, r.out_len[0], cudaMemcpyDeviceToHost, s.stream));
  if (ctx->paired && r.out_len[1]) {
    if (!out2) { ctx->last_error = "out2 is required in paired mode"; return BK_E_ARG; }
    CK(cudaMemcpyAsync(out2, s.out[1].p, r.out_len[1], cudaMemcpyDeviceToHost, s.stream));
  }
  CK(cudaStreamSynchronize(s.stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
  fill_result(res, r, s.n, ms, s.launches);
  if (res) {
    res->h2d_bytes = s.h2d_bytes;
    res->d2h_bytes = sizeof(bk::DevResult) + r.out_len[0] + (ctx->paired ? r.out_len[1] : 0);
  }
  return BK_OK;
}

int bk_extract_host(bk_ctx* ctx, uint32_t n, const bk_mate_in* in1, const bk_mate_in* in2, uint8_t* out1,
                    uint64_t cap1, uint8_t* out2, uint64_t cap2, bk_result* res) {
  int rc = bk_submit(ctx, 0, n, in1, in2);
  if (rc != BK_OK) return rc;
  return bk_wait(ctx, 0, out1, cap1, out2, cap2, res);
}

int bk_dev_alloc(bk_ctx* ctx, uint64_t bytes, void** dptr) {
  if (!ctx || !dptr) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMalloc(dptr, bytes ? bytes : 16));
  return BK_OK;
}
int bk_dev_free(bk_ctx* ctx, void* dptr) {
  if (!ctx) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaFree(dptr));
  return BK_OK;
}
int bk_memcpy_h2d(bk_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return BK_OK;
}
int bk_memcpy_d2h(bk_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return BK_OK;
}

int bk_extract_device(bk_ctx* ctx, uint32_t n, const bk_mate_in* d_in1, const bk_mate_in* d_in2, uint8_t* d_out1,
                      uint64_t cap1, uint8_t* d_out2, uint64_t cap2, int32_t* d_match1, int32_t* d_match2,
                      int iters, bk_result* res) {
  if (!ctx || !d_in1 || iters < 1) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[0];
  if (s.busy) { ctx->last_error = "slot 0 is busy"; return BK_E_ARG; }
  uint32_t launches = 0;
  CK(cudaEventRecord(s.ev0, s.stream));
  for (int it = 0; it < iters; ++it) {
    int rc = enqueue_extract(ctx, &s, s.stream, n, d_in1, d_in2, d_out1, cap1, d_out2, cap2, d_match1, d_match2,
                             &launches);
    if (rc != BK_OK) return rc;
  }
  CK(cudaEventRecord(s.ev1, s.stream));
  CK(cudaMemcpyAsync(s.h_res, (uint8_t*)s.scratch.buf.p + 64, sizeof(bk::DevResult), cudaMemcpyDeviceToHost,
                     s.stream));
  CK(cudaStreamSynchronize(s.stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, s.ev0, s.ev1));
  {
    const int rcc = compact_redo(ctx, s);
    if (rcc != BK_OK) return rcc;
  }
  bk::DevResult r = *s.h_res;
  fill_result(res, r, n, ms / iters, launches);
  return check_dev_result(ctx, r);
}

int bk_debug_events(bk_ctx* ctx, uint32_t ntiles, uint32_t first_tile, uint32_t count, uint64_t* out) {
  if (!ctx || !out) return BK_E_ARG;
#ifdef BK_PROFILE_PHASES
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[0];
  if (!s.scratch.buf.p || first_tile + count > ntiles) return BK_E_ARG;
  CK(cudaStreamSynchronize(s.stream));
  CK(cudaMemcpy(out, (uint8_t*)s.scratch.buf.p + events_offset(ntiles) + (size_t)first_tile * 128, (size_t)count * 128,
                cudaMemcpyDeviceToHost));
  return BK_OK;
#else
  ctx->last_error = "library was not built with -DBK_PROFILE_PHASES";
  return BK_E_ARG;
#endif
}

int bk_tokenize_device(bk_ctx* ctx, const uint8_t* d_text, uint64_t len, int final_chunk, uint32_t* d_rec_off,
                       uint32_t cap_records, uint32_t* n_out, uint64_t* consumed, uint32_t* max_rec_bytes) {
  if (!ctx || !d_rec_off || !n_out || (len && !d_text)) return BK_E_ARG;
  if (((uintptr_t)d_text & 15) != 0) { ctx->last_error = "batch text must be 16-byte aligned"; return BK_E_ARG; }
  if (len >= (1ull << 32) - 65536) { ctx->last_error = "batch text must be < 4 GiB - 64 KiB"; return BK_E_CAPACITY; }  // 32-bit positions, 16 KiB chunks
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->slot[0].stream;
  *n_out = 0;
  if (consumed) *consumed = 0;
  if (max_rec_bytes) *max_rec_bytes = 0;
  CK(cudaMemsetAsync(d_rec_off, 0, 4, st));
  // Blank lines after the last record of the input are accepted and dropped (as bk_tokenize does): the
  // scan ends behind the first newline of the text's trailing run of line ends.
  uint64_t eff = len;
  if (final_chunk && len) {
    uint8_t tail[4096];
    const uint64_t tn = len < sizeof tail ? len : sizeof tail;
    CK(cudaMemcpyAsync(tail, d_text + (len - tn), tn, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    uint64_t k = tn;
    while (k > 0 && (tail[k - 1] == '\n' || tail[k - 1] == '\r')) --k;
    if (k == 0 && tn == len) {  // nothing but line ends
      if (consumed) *consumed = len;
      return BK_OK;
    }
    if (k > 0) {  // (a longer run than the window: leave it to the format check)
      uint64_t j = k;
      while (j < tn && tail[j] != '\n') ++j;
      eff = len - tn + (j < tn ? j + 1 : tn);
    }
  }
  if (eff == 0) return BK_OK;
  const uint32_t nblocks = (uint32_t)((eff + bk::kTokChunk - 1) / bk::kTokChunk);
  int rc = ensure(ctx, &ctx->tok, bk::tok_desc_words(nblocks) * 8 + 128);
  if (rc != BK_OK) return rc;
  bk::TokResult* d_res = (bk::TokResult*)ctx->tok.p;
  unsigned int* ticket = (unsigned int*)((uint8_t*)ctx->tok.p + 64);
  unsigned long long* desc = (unsigned long long*)((uint8_t*)ctx->tok.p + 128);
  bk::tok_init_kernel<<<(unsigned)((bk::tok_desc_words(nblocks) + 255) / 256), 256, 0, st>>>(d_res, desc, nblocks, ticket);
  bk::tok_fused_kernel<<<nblocks, bk::kTokThreads, 0, st>>>(d_text, (uint32_t)eff, d_rec_off, cap_records, desc, ticket, nblocks, d_res);
  CK(cudaGetLastError());
  bk::TokResult r;
  CK(cudaMemcpyAsync(&r, d_res, sizeof r, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  uint64_t n = r.newlines / 4;
  bool truncated = false;  // bytes behind the last complete record that do not make up a record
  uint32_t last_end = 0;
  bool have_last_end = false, have_after = false;
  uint32_t after_n = 0;  // rec_off[n] as the single pass left it
  if (n < cap_records) {
    uint32_t after = 0;  // end of the last record that ends with a newline
    CK(cudaMemcpyAsync(&after, d_rec_off + n, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    after_n = after;
    have_after = true;
    if (after < eff) {
      if (final_chunk && r.newlines % 4 == 3) {  // the final record's quality line has no EOL
        ++n;
        last_end = (uint32_t)eff;
        have_last_end = true;
        CK(cudaMemcpyAsync(d_rec_off + n, &last_end, 4, cudaMemcpyHostToDevice, st));
      } else if (final_chunk) {
        truncated = true;
      }
    }
  } else {
    n = cap_records;
  }
  if (!have_last_end && have_after) {
    last_end = after_n;  // (read above)
  } else if (!have_last_end) {
    CK(cudaMemcpyAsync(&last_end, d_rec_off + n, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  if (have_last_end) {  // the final record without EOL is the one record the single pass did not see the end of
    bk::tok_check_kernel<<<1, 32, 0, st>>>(d_text, d_rec_off, (uint32_t)n - 1, (uint32_t)n, last_end, d_res);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&r, d_res, sizeof r, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  rc = truncated ? BK_E_FORMAT : BK_OK;
  if (n && r.first_bad < n) {  // records before the first malformed one stand, as with bk_tokenize
    n = r.first_bad;
    CK(cudaMemcpyAsync(&last_end, d_rec_off + n, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    rc = BK_E_FORMAT;
  }
  if (rc == BK_E_FORMAT) ctx->last_error = "malformed FASTQ text";
  *n_out = (uint32_t)n;
  if (consumed) *consumed = (rc == BK_OK && final_chunk && eff != len && last_end == eff) ? len : last_end;
  if (max_rec_bytes) *max_rec_bytes = n ? r.max_rec_bytes : 0;
  return rc;
}

int bk_bus_probe(bk_ctx* ctx, uint64_t bytes, int iters, double* h2d_gbs, double* d2h_gbs, double* duplex_gbs) {
  if (!ctx || bytes == 0 || iters < 1) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->slot[0].busy || ctx->slot[1].busy) { ctx->last_error = "a slot is busy"; return BK_E_ARG; }
  void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
  auto cleanup = [&]() {
    if (h_in) cudaFreeHost(h_in);
    if (h_out) cudaFreeHost(h_out);
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
  };
  if (cudaHostAlloc(&h_in, bytes, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault) != cudaSuccess ||
      cudaMalloc(&d_in, bytes) != cudaSuccess || cudaMalloc(&d_out, bytes) != cudaSuccess) {
    ctx->last_error = std::string("bus probe: allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    cleanup();
    return BK_E_CUDA;
  }
  memset(h_in, 0x5a, bytes);
  memset(h_out, 0, bytes);
  cudaStream_t s0 = ctx->slot[0].stream, s1 = ctx->slot[1].stream;
  auto now = []() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + t.tv_nsec * 1e-9; };
  auto run = [&](bool up, bool down) -> double {  // seconds for `iters` copies per active direction
    cudaStreamSynchronize(s0);
    cudaStreamSynchronize(s1);
    const double t0 = now();
    for (int i = 0; i < iters; ++i) {
      if (up) cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s0);
      if (down) cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s1);
    }
    cudaStreamSynchronize(s0);
    cudaStreamSynchronize(s1);
    return now() - t0;
  };
  run(true, true);  // warm-up: page tables, link power state
  const double gb = (double)bytes * iters / 1e9;
  const double tu = run(true, false), td = run(false, true), tb = run(true, true);
  if (h2d_gbs) *h2d_gbs = gb / tu;
  if (d2h_gbs) *d2h_gbs = gb / td;
  if (duplex_gbs) *duplex_gbs = 2.0 * gb / tb;
  const cudaError_t e = cudaGetLastError();
  cleanup();
  if (e != cudaSuccess) { ctx->last_error = std::string("bus probe: ") + cudaGetErrorString(e); return BK_E_CUDA; }
  return BK_OK;
}

int bk_host_mem_probe(uint64_t bytes, int threads, int iters, double* gbs) {
  if (!gbs || threads < 1 || iters < 1 || bytes < 4096) return BK_E_ARG;
  std::vector<uint8_t*> src(threads, nullptr), dst(threads, nullptr);
  bool ok = true;
  for (int t = 0; t < threads; ++t) {
    src[t] = (uint8_t*)malloc(bytes);
    dst[t] = (uint8_t*)malloc(bytes);
    ok = ok && src[t] && dst[t];
  }
  double best = 0;
  if (ok) {
    auto pass = [&](int t) { memcpy(dst[t], src[t], bytes); };
    for (int it = 0; it <= iters; ++it) {  // the first pass faults the pages in
      if (it == 0)
        for (int t = 0; t < threads; ++t) memset(src[t], 1 + t, bytes);
      timespec a, b;
      clock_gettime(CLOCK_MONOTONIC, &a);
      std::vector<std::thread> th;
      for (int t = 0; t < threads; ++t) th.emplace_back(pass, t);
      for (auto& x : th) x.join();
      clock_gettime(CLOCK_MONOTONIC, &b);
      const double s = (b.tv_sec - a.tv_sec) + (b.tv_nsec - a.tv_nsec) / 1e9;
      if (it > 0 && s > 0) best = std::max(best, 2.0 * (double)bytes * threads / s / 1e9);
    }
  }
  for (int t = 0; t < threads; ++t) { free(src[t]); free(dst[t]); }
  if (!ok) return BK_E_CAPACITY;
  *gbs = best;
  return BK_OK;
}

int bk_reverse_complement(bk_ctx* ctx, const uint8_t* seq, size_t n, uint8_t* out) {
  if (!ctx || (n && (!seq || !out))) return BK_E_ARG;
  if (n == 0) return BK_OK;
  if (n >= (1ull << 32)) return BK_E_CAPACITY;
  CK(cudaSetDevice(ctx->device));
  Slot& s = ctx->slot[0];
  int rc;
  if ((rc = ensure(ctx, &s.text[0], n + 64)) != BK_OK) return rc;
  if ((rc = ensure(ctx, &s.out[0], n + 64)) != BK_OK) return rc;
  CK(cudaMemcpyAsync(s.text[0].p, seq, n, cudaMemcpyHostToDevice, s.stream));
  int blocks = (int)std::min<size_t>((n + 255) / 256, 1184);
  bk::rc_kernel<<<blocks, 256, 0, s.stream>>>((const uint8_t*)s.text[0].p, (uint8_t*)s.out[0].p, (uint32_t)n);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, s.out[0].p, n, cudaMemcpyDeviceToHost, s.stream));
  CK(cudaStreamSynchronize(s.stream));
  return BK_OK;
}

int bk_generate_device(bk_ctx* ctx, const bk_syn_config* cfg, int mate, uint64_t seed, uint64_t first_idx,
                       uint32_t n, uint8_t* d_text, uint32_t* d_rec_off) {
  if (!ctx || !cfg || mate < 1 || mate > 2 || !d_text || !d_rec_off) return BK_E_ARG;
  CK(cudaSetDevice(ctx->device));
  const uint64_t rl = 23 + 2ull * cfg->read_len + 6;
  if (rl * n >= (1ull << 32)) { ctx->last_error = "batch text must be < 4 GiB per mate"; return BK_E_CAPACITY; }
  if (cfg->plant_len[mate - 1] > 32) { ctx->last_error = "planted literal longer than 32"; return BK_E_ARG; }
  bk::GenParams G;
  memset(&G, 0, sizeof G);
  G.seed = seed;
  G.first_idx = first_idx;
  G.n = n;
  G.read_len = cfg->read_len;
  G.mate = (uint32_t)mate;
  G.plant_len = cfg->plant_len[mate - 1];
  G.plant_off = cfg->plant_off[mate - 1];
  G.plant_jitter = cfg->plant_jitter[mate - 1];
  memcpy(G.plant_lit, cfg->plant_lit[mate - 1], 32);
  Slot& s = ctx->slot[0];
  int threads = 128;
  int blocks = (int)((n + threads - 1) / threads);
  if (blocks < 1) blocks = 1;
  bk::gen_kernel<<<blocks, threads, 0, s.stream>>>(G, d_text, d_rec_off);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s.stream));
  return BK_OK;
}

}  // extern "C"
