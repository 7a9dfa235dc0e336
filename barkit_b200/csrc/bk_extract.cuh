// The extract kernel: warp-specialised tile pipeline with ordered compaction (decoupled look-back).
// Building blocks (parse / match / emit / PTX wrappers) are in bk_kernels.cuh.
#pragma once
#include "bk_kernels.cuh"

namespace bk {

constexpr int kStages = 4;  // input ring depth per mate (the loader runs at most this far ahead)
constexpr int kBackWarps = kEmitParts;  // back warps per mate of a paired launch: warp j writes part j of every record (emit_part)
// ... and of a single-end launch.  4: A+B | C | D | E.  6 or 7: the seq and the qual line are written by two warps each
// (the two long jobs were the back stage's critical path: the slot of a tile is held until its slowest part is done).
static_assert(BK_SE_BACKS == 4 || BK_SE_BACKS == 5 || BK_SE_BACKS == 6 || BK_SE_BACKS == 7, "emit_part knows 3, 4, 5, 6 or 7 parts");
__host__ __device__ constexpr int back_warps(bool paired) { return paired ? kBackWarps : BK_SE_BACKS; }

// ---- ordered compaction: two-level decoupled look-back ---------------------------------------------------
//
// Tile descriptors (8 bytes): the tile's own output bytes on stream 0 | stream 1, each half tagged
// with kTileReady; the two front warps publish their halves independently.
// Block descriptors (16 bytes, one per 32 consecutive tiles): {stream 0, stream 1}, each a 62-bit
// value tagged kFlagAgg (the block's own output bytes) or kFlagPrefix (inclusive prefix through the
// block's last tile).  The aggregate is published by whichever front warp adds the block's last
// tile to the block's accumulator (so it depends on no scan warp's progress), the prefix by the
// scan warp of the block's last tile.
//
// A tile's exclusive prefix = the aggregates of the earlier tiles of its own block (one window of
// tile descriptors) + a look-back over whole blocks.  With kStages tiles per CTA in flight, several
// hundred claimed tiles are unresolved at any time; a flat look-back walks all of them, 128 per L2
// round trip, and its latency in turn keeps more tiles unresolved.  Over blocks the same distance is
// one window, and both windows' loads are issued together: two round trips become one.

constexpr uint32_t kTileReady = 0x80000000u;

__device__ __forceinline__ ulonglong2 ld_desc(const unsigned long long* p) {
  ulonglong2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_desc(unsigned long long* p, unsigned long long x, unsigned long long y) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(x), "l"(y) : "memory");
}
// usable once both halves are there and of the same kind (a half-updated descriptor is simply read again)
template <bool two>
__device__ __forceinline__ bool desc_ready(const ulonglong2& d) {
  const unsigned f0 = (unsigned)(d.x >> 62), f1 = (unsigned)(d.y >> 62);
  return f0 != 0 && (!two || (f1 != 0 && (f0 == 2) == (f1 == 2)));
}
// descriptors only ever move up: nothing -> aggregate -> prefix (the tags order that way)
__device__ __forceinline__ void desc_raise(unsigned long long* p, unsigned long long v) {
  asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <bool two>
__device__ __forceinline__ bool tile_ready(unsigned long long t) {
  return ((uint32_t)t & kTileReady) && (!two || ((uint32_t)(t >> 32) & kTileReady));
}

// Exclusive prefixes of tile `tile` on both streams.  The tile's own aggregates (agg0, agg1) have
// already been published by the front warps.
template <bool two>
__device__ __noinline__ ulonglong2 lookback(unsigned long long* tdesc, unsigned long long* bdesc, uint32_t tile,
                                            uint32_t agg0, uint32_t agg1, int lane) {
  const uint32_t blk = tile >> 5, j = tile & 31u;
  // window 1: the earlier tiles of this block; window 2: the 32 blocks before it (lane 0 = nearest)
  const bool in_blk = (uint32_t)lane < j;
  const unsigned long long* tp = tdesc + (size_t)(blk << 5) + lane;
  int look = (int)blk - 1;
  unsigned long long t = in_blk ? ld_relaxed_u64(tp) : 0ull;
  ulonglong2 d = look - lane >= 0 ? ld_desc(bdesc + 2 * (size_t)(look - lane)) : make_ulonglong2(kFlagPrefix, kFlagPrefix);
  // both windows are polled together: whichever is late, the other's round trip is not added to it
  while (__any_sync(0xffffffffu, (in_blk && !tile_ready<two>(t)) || !desc_ready<two>(d))) {
    __nanosleep(32);  // a predecessor is still being parsed: do not burn issue slots and L2 bandwidth on it
    if (in_blk && !tile_ready<two>(t)) t = ld_relaxed_u64(tp);
    if (!desc_ready<two>(d)) d = ld_desc(bdesc + 2 * (size_t)(look - lane));
  }
  const uint32_t s0 = __reduce_add_sync(0xffffffffu, in_blk ? (uint32_t)t & ~kTileReady : 0u);
  const uint32_t s1 = two ? __reduce_add_sync(0xffffffffu, in_blk ? (uint32_t)(t >> 32) & ~kTileReady : 0u) : 0u;
  const bool last = j == 31u;  // this tile completes its block
  unsigned long long e0 = s0, e1 = s1;
  while (true) {  // blocks before 0 read as "prefix 0", so this ends at the latest there
    while (__any_sync(0xffffffffu, !desc_ready<two>(d))) {
      __nanosleep(64);
      if (!desc_ready<two>(d)) d = ld_desc(bdesc + 2 * (size_t)(look - lane));
    }
    const unsigned pm = __ballot_sync(0xffffffffu, (d.x >> 62) == 2);
    const int first = pm ? __ffs(pm) - 1 : 32;  // nearest block holding inclusive prefixes
    // a block's aggregate is at most 32 tiles' bytes (32 bits is plenty); only prefixes are 64-bit
    e0 += __reduce_add_sync(0xffffffffu, lane < first ? (uint32_t)d.x : 0u);
    if (two) e1 += __reduce_add_sync(0xffffffffu, lane < first ? (uint32_t)d.y : 0u);
    if (pm) {
      e0 += __shfl_sync(0xffffffffu, d.x & kValueMask, first);
      if (two) e1 += __shfl_sync(0xffffffffu, d.y & kValueMask, first);
      break;
    }
    look -= 32;
    d = look - lane >= 0 ? ld_desc(bdesc + 2 * (size_t)(look - lane)) : make_ulonglong2(kFlagPrefix, kFlagPrefix);
  }
  if (last && lane == 0) {
    desc_raise(bdesc + 2 * (size_t)blk, kFlagPrefix | (e0 + agg0));
    desc_raise(bdesc + 2 * (size_t)blk + 1, kFlagPrefix | (e1 + agg1));
  }
  return make_ulonglong2(e0, e1);
}

// A mate's lookup block in shared memory (layout: bk_kernels.cuh, kTabComp .. kTabSA), filled by one warp.
__device__ __forceinline__ void fill_tab(uint8_t* t, const bk_devprog& pg, int lane, bool general) {
  // byte -> "which classes contain it" table, so a class test is one shared-memory lookup
  for (int c = lane; c < 256; c += 32) {
    uint32_t mask = 0;
    if (c < 128)
      for (int k = 0; k < BK_MAX_CLASSES; ++k) mask |= ((pg.cls[k][c >> 5] >> (c & 31)) & 1u) << k;
    t[c] = (uint8_t)mask;
    t[kTabComp + c] = (uint8_t)complement((uint32_t)c);
  }
  // ... and the program behind them (layout: kTabHdr, kTabOps, kTabLit)
  uint32_t* const tw = reinterpret_cast<uint32_t*>(t);
  if (lane == 0) {
    tw[kTabHdr / 4 + 0] = pg.nops;
    tw[kTabHdr / 4 + 1] = pg.total_len;
    tw[kTabHdr / 4 + 2] = pg.anchor_start;
    tw[kTabHdr / 4 + 3] = pg.anchor_end;
  }
  for (int i = lane; i < 2 * BK_MAX_OPS; i += 32) tw[kTabOps / 4 + i] = reinterpret_cast<const uint32_t*>(pg.ops)[i];
  for (int i = lane; i < BK_MAX_LIT / 4; i += 32) tw[kTabLit / 4 + i] = reinterpret_cast<const uint32_t*>(pg.lit)[i];
  if (lane == 0) tw[kTabIns / 4] = pg.nins;  // parametric programs: the insertion points (match_par)
  if (lane < 2 * BK_MAX_INS) tw[kTabIns / 4 + 2 + lane] = reinterpret_cast<const uint32_t*>(pg.ins)[lane];
  if (general && lane < 8) {  // bit planes of the first literal (match_unanchored): bits 1 and 2 of byte `lane`
    const uint32_t c = pg.nops && pg.ops[0].kind == BK_OP_LIT ? (uint32_t)pg.lit[pg.ops[0].idx + lane] : 0u;
    tw[kTabSA / 4 + 2 * lane] = (c & 2u) ? 0xFFFFFFFFu : 0u;
    tw[kTabSA / 4 + 2 * lane + 1] = (c & 4u) ? 0xFFFFFFFFu : 0u;
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------
//
// One LOADER warp per CTA feeds a ring of kStages staged input tiles per mate (one mate for
// single-end, two for paired-end); per mate a FRONT warp, a SCAN warp and kBackWarps BACK warps
// drain it:
//
//   loader L  : wait ring slot free (both mates) -> claim tile (global counter, in order) -> record
//               offsets -> one 1-D TMA load per mate of the tile's contiguous text range
//   front F_m : wait tile landed -> parse -> match -> exchange match decisions with the other mate's
//               front warp (PE keep rule, run.rs:149-160) -> sizes + warp scan -> PUBLISH the tile
//               aggregate -> hand the slot on
//   scan  S_m : take slot -> look-back over earlier tiles -> publish the inclusive prefix -> pass the
//               tile's global output offset to the back warps
//   back  B_m : take slot -> assemble the tile's output in shared memory (header lines lane per
//               record, long runs warp-cooperatively) -> release slot -> one warp issues the TMA store
//
// Claim, offset fetch and load latency sit in the loader, off the front warp's critical path, and
// every role only ever waits on its own ring slot, so a claimed tile's aggregate is published
// after a bounded delay; that is what keeps every other tile's look-back from stalling.  Lane i of
// a front/back warp owns record i of the tile.

// Optional event trace (compile with -DBK_PROFILE_PHASES; tools/phases.py): one clock64() stamp per
// tile and pipeline event, all taken on the tile's own SM, so differences between them are valid.
#ifdef BK_PROFILE_PHASES
#define EVT(tile, idx) { if (lane == 0) P.prof[(size_t)(tile) * 16 + (idx)] = (unsigned long long)clock64(); }
#else
#define EVT(tile, idx)
#endif

// named barrier of one mate's back warps (ids 2 and 3; 0 is __syncthreads, 1 the front warps')
__device__ __forceinline__ void back_sync(int m, uint32_t nwarps) { nbar_sync(2u + (uint32_t)m, 32u * nwarps); }

// Named barriers (16 per CTA): 0 __syncthreads, 1 the front warps' exchange, 2-3 back_sync, then one
// per ring slot for fronts -> scan ("aggregates published") and scan -> backs ("offsets known").
// A slot's barrier is reused only after the slot went round the ring, i.e. after every consumer
// has left it.
// (16 ids: a paired CTA may have at most 5 ring slots, a single-end one, which has no exchange, 6)
template <int ST> __device__ __forceinline__ uint32_t pub_id(uint32_t s) { return 4u + s; }
template <int ST> __device__ __forceinline__ uint32_t ready_id(uint32_t s) { return 4u + ST + s; }
constexpr int kScanWarps = 2;  // scan warp j resolves the tiles of ring rounds k with k % 2 == j

constexpr int kSkipTile = -2147483647 - 1;
constexpr uint32_t kEndTile = 0xFFFFFFFFu, kEndTile2 = 0xFFFFFFFEu;
constexpr uint32_t kNotKept = 0xFFFFFFFFu;

struct TileMeta {  // hand-over between the roles, one per (mate, stage); SoA so lanes do not conflict
  // (every array here is shared memory the ring does not get: bounds as 33 starts, the two line lengths in one word)
  uint32_t beg[33];  // record bounds, shared-memory addresses: record i is [beg[i], beg[i + 1]) (loader)
  uint32_t head_len_fast[32], seq[32], lens[32], qual[32];  // head_len | fast << 31, seq_len | qual_len << 16 (front)
  int32_t ms[32];
  uint32_t ooff[32];  // record's offset inside the tile's output, kNotKept if dropped
  uint32_t tile;      // kEndTile terminates the consumers (loader)
  uint32_t nrec;      // records in the tile (loader)
  uint32_t too_big;   // tile does not fit the staging buffer: nothing was loaded (loader)
  uint32_t total;     // output bytes of the tile on this stream (front)
  unsigned long long gout;  // global output offset of the tile on this stream (scan)
};

#ifndef BK_SE_FRONTS
#define BK_SE_FRONTS 2
#endif
// front warps per mate: warp p takes the ring rounds k with k % front_pairs == p
__host__ __device__ constexpr int front_pairs(bool paired) { return paired ? 2 : BK_SE_FRONTS; }
// A ring slot's barriers must always meet the same scan warp and front pair.  With an odd ring depth
// round k and round k + kStages share their barrier ids but belong to different scan warps, and the
// one running ahead completes the other's barrier generation (observed: a 3-deep ring hangs).
__host__ __device__ constexpr int cta_warps(bool paired) { return (paired ? 2 : 1) * (front_pairs(paired) + back_warps(paired)) + kScanWarps + 1; }

// GENERAL = false: every pattern is anchored and there is no reverse-complement retry -- the common
// case, compiled without the search code, because every kilobyte of this kernel competes for the
// SM's instruction cache.
// VAR (only with GENERAL): a mate's pattern may be a list of fixed-length variants (bounded variable
// repetition, bk_pattern.cc); the matched variant travels with the match start (start | variant << 24) and
// supplies the capture offsets, the match length and the header growth of that record.
// TABLE (single-end launches only): nothing is matched here -- the read's match (start | variant << 24, or -1) was
// left in P.tab_in[0] by match_kernel (bk_match.cuh) earlier in the stream, so this instantiation carries no search
// code at all.  Independently of TABLE, a single-end launch that is one mate of a pair takes the OTHER mate's match
// table from P.tab_in[1] for the keep rule (run.rs:149-160).
template <bool PAIRED, bool GENERAL, bool VAR = false, int ST = kStages, bool TABLE = false>
__global__ void __launch_bounds__(cta_warps(PAIRED) * 32, PAIRED ? 2 : 4) extract_kernel(const __grid_constant__ KParams P) {
  static_assert(!VAR || GENERAL || TABLE, "variants are matched by the general kernel or by match_kernel");
  static_assert(!TABLE || !GENERAL, "a launch that takes matches from a table has no search code");
  constexpr int kFrontPairs = PAIRED ? 2 : BK_SE_FRONTS;
  static_assert(ST % kScanWarps == 0 && ST % kFrontPairs == 0 && 4 + 2 * ST + (PAIRED ? 1 : 0) <= 16, "ring depth: a multiple of the scan / front warp counts, and 16 named barriers");
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NM = PAIRED ? 2 : 1;
  __shared__ __align__(8) uint64_t bar_full[NM][ST], bar_empty[NM][ST];
  // (static shared memory counts against the ring: a single-end launch carries one lookup block and no exchange buffer,
  // 2.5 KB less per CTA, which is one more record per tile at four CTAs per SM)
  __shared__ __align__(16) uint8_t cls_tab[NM][kTabBytes];  // per mate: class bits per byte, complement table, program
  __shared__ int s_ms[PAIRED ? kFrontPairs : 1][2][PAIRED ? 2 : 1][PAIRED ? 32 : 1];  // [front pair][round parity][mate][lane]
  // search sharing (GENERAL, paired, exactly one mate unanchored): the searching mate's reads {seq, len, can}
  // for the other mate's front warp, and that warp's results for the upper half of the tile
  __shared__ uint32_t s_share[GENERAL && PAIRED ? kFrontPairs : 1][2][3][32];
  __shared__ int s_help[GENERAL && PAIRED ? kFrontPairs : 1][2][32];
  __shared__ int s_res[GENERAL && PAIRED ? kFrontPairs : 1][2][2][32];  // match_unanchored shared by two warps: [round parity][warp][lane]
  __shared__ TileMeta meta[NM][ST];
  __shared__ uint4 run_desc[NM][16];  // copy descriptors of the tile's runs of unchanged records (runs are at least one record apart: <= 16)

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // Warp roles, lowest warp ids first: back, scan, front, loader.
  constexpr int NB = NM * back_warps(PAIRED);
  const bool is_back = warp < NB;
  const bool is_scan = !is_back && warp < NB + kScanWarps;
  const bool is_front = !is_back && !is_scan && warp < NB + kScanWarps + NM * kFrontPairs;
  const bool is_loader = warp == NB + kScanWarps + NM * kFrontPairs;
  const int fw = warp - NB - kScanWarps;  // front warp index
  // Back warps per mate: four each, or five and three when only one mate's reads are rewritten (its
  // records cost more to assemble than the other mate's verbatim copies).
  const uint32_t nb0 = !PAIRED ? back_warps(false)
                       : (P.prog[0].present && !P.prog[1].present) ? kBackWarps + 1
                       : (!P.prog[0].present && P.prog[1].present) ? kBackWarps - 1 : kBackWarps;
  const int m = is_back ? (warp < (int)nb0 ? 0 : 1) : is_front ? fw % NM : 0;  // mate = output stream
  const int fp = is_front ? fw / NM : 0;  // front pair: which ring rounds this front warp takes
  const int bj = is_back ? (m ? warp - (int)nb0 : warp) : 0;  // back warp index within the mate
  const uint32_t nparts = m ? NB - nb0 : nb0;                 // back warps of this mate
  constexpr uint32_t kPubCount = 32 * (NM + 1), kReadyCount = 32 * (1 + NB);
  const bk_devprog& pg = P.prog[m];
  const bool trim = !(P.flags & 2u);
  const bool rc = (P.flags & 4u) != 0;
  const uint32_t R = P.tile_recs;

  // dynamic shared memory: [mate0 ring][mate1 ring][out0][out1]
  // (32-bit shared-space addresses; see the note at the top of bk_kernels.cuh)
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t out_s = smem_base + ST * (P.in_cap[0] + (PAIRED ? P.in_cap[1] : 0)) + (m ? P.out_cap[0] : 0);
  const uint32_t tab = smem_u32(&cls_tab[m][0]);

  if (is_front && fp == 0) {
    if (!TABLE || PAIRED) fill_tab(&cls_tab[m][0], pg, lane, GENERAL);
    if (lane == 0)
      for (int s = 0; s < ST; ++s) {
        mbar_init(&bar_full[m][s], 1);
        mbar_init(&bar_empty[m][s], m ? NB - nb0 : nb0);
      }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  if (is_loader) {
    // =============================== LOADER =====================================================
    for (uint32_t k = 0;; ++k) {
      const uint32_t s = k % ST, ph = (k / ST) & 1u;
      // slot free on every mate (passes immediately for the first ST tiles)
#pragma unroll
      for (int mm = 0; mm < NM; ++mm) mbar_wait(&bar_empty[mm][s], ph ^ 1u);
#ifdef BK_PROFILE_PHASES
      const long long t_free = clock64();
#endif
      // claim with the slot in hand: the load starts at once and the aggregate follows a bounded
      // time later
      uint32_t tile = 0;
      if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= P.ntiles) {
        // end markers in consecutive slots, one per front warp of a mate (>= one per scan warp: both take the ring
        // rounds in turn); the first one also stops the back warps
        for (uint32_t j = 0; j < (uint32_t)kFrontPairs; ++j) {
          const uint32_t sj = (k + j) % ST, phj = ((k + j) / ST) & 1u;
          if (j)
            for (int mm = 0; mm < NM; ++mm) mbar_wait(&bar_empty[mm][sj], phj ^ 1u);
          if (lane == 0)
            for (int mm = 0; mm < NM; ++mm) {
              meta[mm][sj].tile = j ? kEndTile2 : kEndTile;
              mbar_arrive(&bar_full[mm][sj]);
            }
        }
        break;
      }
#ifdef BK_PROFILE_PHASES
      if (lane == 0) P.prof[(size_t)tile * 16 + 0] = (unsigned long long)t_free;
#endif
      EVT(tile, 1)
      const uint32_t r0 = tile * R;
      const uint32_t nrec = min(R, P.n - r0);
      uint32_t o[NM], oe[NM];
#pragma unroll
      for (int mm = 0; mm < NM; ++mm) {  // both mates' offsets in flight together
        o[mm] = lane < (int)nrec ? P.off[mm][r0 + lane] : 0;
        oe[mm] = P.off[mm][r0 + nrec];
      }
#pragma unroll 1  // (rolled: costs ~0.9 k cycles between claim and load, saves 1 KB of the instruction cache)
      for (int mm = 0; mm < NM; ++mm) {
        TileMeta& tm = meta[mm][s];
        const uint32_t om = mm ? o[NM - 1] : o[0], oem = mm ? oe[NM - 1] : oe[0];
        uint32_t onext = __shfl_down_sync(0xffffffffu, om, 1);
        if (lane == (int)nrec - 1) onext = oem;
        const uint32_t gbase = __shfl_sync(0xffffffffu, om, 0) & ~15u;
        const uint32_t bytes = ((oem + 15u) & ~15u) - gbase;
        const bool too_big = bytes > P.in_cap[mm];  // host sized the ring from max_rec_bytes
        const uint32_t in_s = smem_base + (mm ? ST * P.in_cap[0] : 0) + s * P.in_cap[mm];
        if (lane < (int)nrec) tm.beg[lane] = in_s + (om - gbase);
        if (lane == (int)nrec - 1) tm.beg[nrec] = in_s + (onext - gbase);
        if (lane == 0) {
          tm.tile = tile;
          tm.nrec = nrec;
          tm.too_big = too_big;
        }
        __syncwarp();
        if (lane == 0) {
          if (too_big) {
            mbar_arrive(&bar_full[mm][s]);
          } else {
            mbar_arrive_expect_tx(&bar_full[mm][s], bytes);
            bulk_g2s(in_s, P.text[mm] + gbase, bytes, &bar_full[mm][s]);
          }
          // Warm L2 for the tile this CTA will most likely claim next (every CTA claims about one tile
          // per ring round): its record offsets, and its text where that would start if its records are
          // as long as this tile's.  Claim -> offsets -> load is a chain of three DRAM round trips
          // that every later tile's look-back waits on; two of them become L2 hits.
          const uint32_t nt = tile + gridDim.x;
          if (nt < P.ntiles) {
            bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(P.off[mm]) + (((size_t)nt * R * 4u) & ~size_t(15)), ((4u * R + 15u) & ~15u) + 16u);
            const unsigned long long est = (unsigned long long)gbase + (unsigned long long)gridDim.x * (oem - om);  // (lane 0: om is the tile's first offset)
            if (est + bytes + 256u <= P.text_len[mm]) bulk_prefetch_l2(P.text[mm] + (est & ~15ull), bytes + 256u);
          }
        }
      }
      EVT(tile, 2)
    }
  } else if (is_front) {
    // =============================== FRONT ======================================================
    // Two front warps per mate take alternate ring rounds: at one warp per mate the front was 80 % busy,
    // tiles queued behind one another between claim and publication, and every later tile's look-back
    // waited for the slowest of them.
    uint32_t n_kept = 0, n_matched = 0;
    // An unanchored search costs ten times an anchored test.  When exactly one mate needs one, the
    // other mate's front warp of the pair -- otherwise idle until the exchange -- searches the upper
    // half of the tile's reads for it.
    const bool un_me = pg.present && !pg.anchor_start && !pg.anchor_end;
    const bk_devprog& pgo = P.prog[PAIRED ? m ^ 1 : m];
    const bool un_other = PAIRED && pgo.present && !pgo.anchor_start && !pgo.anchor_end;
    const bool var_me = VAR && P.nvar[m] > 1u;
    // (by lane halves for the warp-per-read sweep; by windows of starts for the bit-parallel scan: match_unanchored)
    const bk_devprog& pgs = un_me ? pg : pgo;
    const bool scan_form = pgs.nops && bs_eligible(pgs.ops[0].kind, pgs.ops[0].len, pgs.ops[0].k);
    const bool share = GENERAL && !VAR && PAIRED && !rc && un_me != un_other;
    const bool split = share && !scan_form, split2 = share && scan_form;
    const uint32_t pair_bar = fp == 0 ? 1u : 4u + 2u * ST;
    const uint32_t tab_other = smem_u32(&cls_tab[PAIRED ? m ^ 1 : m][0]);
    for (uint32_t k = fp;; k += kFrontPairs) {
      const uint32_t s = k % ST, ph = (k / ST) & 1u;
      const uint32_t xb = (k / kFrontPairs) & 1u;  // exchange buffer of this pair's round
      TileMeta& tm = meta[m][s];
      mbar_wait(&bar_full[m][s], ph);  // offsets written, text landed
      const uint32_t tile = tm.tile;
      if (tile == kEndTile || tile == kEndTile2) {  // one end marker per front pair (and per scan warp)
        nbar_arrive(pub_id<ST>(s), kPubCount);
        break;
      }
      if (m == 0) { EVT(tile, 3) }
      const uint32_t r0 = tile * R;
      const uint32_t nrec = tm.nrec;
      const bool active = lane < (int)nrec;
      const bool too_big = tm.too_big != 0;
      int t_own = -1, t_oth = -1;  // match tables of earlier launches (issued here: the loads overlap the parse)
      if (!PAIRED) {
        if (TABLE && active) t_own = __ldg(P.tab_in[0] + r0 + lane);
        if (P.tab_in[1] && active) t_oth = __ldg(P.tab_in[1] + r0 + lane);
      } else if (TABLE) {  // paired: tab_in[m] is mate m's own table; a mate without one matches for itself (anchored)
        if (P.tab_in[m] && active) t_own = __ldg(P.tab_in[m] + r0 + lane);
      }
      const bool tabbed = TABLE && (!PAIRED || P.tab_in[m] != nullptr);
      RecView rv;
      rv.beg = rv.end = rv.fast = rv.head_len = rv.seq = rv.seq_len = rv.qual = rv.qual_len = 0;
      int ms = -1;
      if (!too_big) {
        // ---- parse + match (lane per record) ---------------------------------------------------
        if (active) rv = parse_record(tm.beg[lane], tm.beg[lane + 1]);
        __syncwarp();
        if (m == 0) { EVT(tile, 4) }
        const bool can = active && pg.present && rv.qual_len == rv.seq_len;  // unequal lengths never leave bk_tokenize
        if (tabbed) {
          ms = t_own;
        } else if (VAR && !TABLE && var_me) {  // (lane per read, bytewise: see match_variants)
          if (can) {
            ms = match_variants<false>(P.vprog[m], P.nvar[m], tab, rv.seq, rv.seq_len);
            if (ms < 0 && rc) {
              ms = match_variants<true>(P.vprog[m], P.nvar[m], tab, rv.seq, rv.seq_len);
              if (ms >= 0 && !caps_utf8_ok(P.vprog[m][ms >> 24], rv.seq, (uint32_t)ms & 0xFFFFFFu)) ms = -1;
            }
          }
        } else if (GENERAL && un_me) {
          if (split || split2) {  // the other mate's front warp takes the upper half of the reads / every other window
            s_share[fp][xb][0][lane] = rv.seq;
            s_share[fp][xb][1][lane] = rv.seq_len;
            s_share[fp][xb][2][lane] = can;
            nbar_sync(pair_bar, 64);
          }
          ms = match_unanchored(tab, rv.seq, rv.seq_len, can && !(split && lane >= 16), lane, 0, split2 ? 2 : 1,
                                &s_res[fp][0][0][0], pair_bar);
        } else if (can) {
          ms = GENERAL ? match_read_shared(tab, rv.seq, rv.seq_len) : match_read<false>(tab, rv.seq, rv.seq_len);  // lane per read
        }
        if (GENERAL && !TABLE && !(VAR && var_me) && can && ms < 0 && rc) {  // parse.rs:61-66
          ms = match_read<true>(tab, rv.seq, rv.seq_len);
          if (ms >= 0 && !caps_utf8_ok(pg, rv.seq, (uint32_t)ms)) ms = -1;
        }
      } else {
        if (lane == 0) atomicOr(&P.result->error, 2u);
        if (GENERAL && (split || split2) && un_me) {  // keep the pair's barriers in step: nothing to search
          s_share[fp][xb][2][lane] = 0;
          nbar_sync(pair_bar, 64);
        }
      }
      if (GENERAL && (split || split2) && !un_me) {  // the other mate searches: take our share of it
        nbar_sync(pair_bar, 64);
        const int h = match_unanchored(tab_other, s_share[fp][xb][0][lane], s_share[fp][xb][1][lane],
                                       s_share[fp][xb][2][lane] != 0 && !(split && lane < 16), lane, 1, split2 ? 2 : 1,
                                       &s_res[fp][0][0][0], pair_bar);
        s_help[fp][xb][lane] = h;
      }
      __syncwarp();
      if (m == 0) { EVT(tile, 5) }
      // ---- exchange match decisions with the other mate's front warp -----------------------------
      int ms_other = PAIRED ? -1 : t_oth;
      bool skip = too_big;
      if (PAIRED) {
        s_ms[fp][xb][m][lane] = too_big ? kSkipTile : ms;
        nbar_sync(fp == 0 ? 1u : 4u + 2u * ST, 64);  // the two mates' front warps of this pair
        ms_other = s_ms[fp][xb][m ^ 1][lane];
        skip = skip || (s_ms[fp][xb][m ^ 1][0] == kSkipTile);
        if (GENERAL && split && lane >= 16) {  // the upper half's search results come from the helper
          const int h = s_help[fp][xb][lane];
          if (un_me) { if (!too_big) ms = h; } else if (ms_other != kSkipTile) ms_other = h;
        }
      }
      if (!too_big && active && P.match[m]) P.match[m][r0 + lane] = (VAR && ms >= 0) ? (ms & 0xFFFFFF) : ms;
      if (!PAIRED && !too_big && active && P.hlen[0]) P.hlen[0][r0 + lane] = rv.head_len;  // (the other mate's compaction guesses with it)
      // run.rs:71-78 (SE: unmatched reads vanish), run.rs:149-160 (PE: keep iff either mate matched)
      const bool keep = !skip && active && (ms >= 0 || ms_other >= 0);
      if (!PAIRED && P.kmask) {  // which reads of this tile stay: all the other mate's compaction needs (bk_compact.cuh)
        const unsigned km = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) st_relaxed_u64(P.kmask + tile, (1ull << 63) | (too_big ? 1ull << 62 : 0ull) | km);
      }

      // ---- size + scan, publish the aggregate -------------------------------------------------------
      const bk_devprog& pe = (VAR && var_me && ms >= 0) ? P.vprog[m][ms >> 24] : pg;  // the record's variant
      uint32_t olen = keep ? out_bytes(pe, rv, ms, trim) : 0;
      if (VAR && TABLE && bk_free_shape(pg) && keep && ms >= 0) {  // a program without a fixed shape: this record's own spans
        const EmitProg ep = emit_prog(pg, P.xtab[m][r0 + lane]);
        olen = out_bytes(ep, rv, ms, trim);
      }
      uint32_t incl = olen;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long blk_old = 0;  // the block's accumulator before this tile: tiles << 48 | bytes
      if (lane == 0) {
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(P.desc[0] + tile) + m),
                     "r"(total | kTileReady)
                     : "memory");
        blk_old = atomicAdd(P.desc[2] + 2 * (size_t)(tile >> 5) + m, (1ull << 48) | (unsigned long long)total);
      }
      EVT(tile, m == 0 ? 6 : 14)

      n_kept += keep;  // statistics: per lane, summed once at the end
      n_matched += active && ms >= 0;

      // ---- hand the slot to the scan and back warps ---------------------------------------------------
      tm.head_len_fast[lane] = rv.head_len | (rv.fast << 31);
      tm.seq[lane] = rv.seq;
      tm.lens[lane] = rv.seq_len | (rv.qual_len << 16);  // (a staged record is < 64 KiB: the ring holds four tiles in 200 KiB)
      tm.qual[lane] = rv.qual;
      tm.ms[lane] = ms;
      tm.ooff[lane] = keep ? incl - olen : kNotKept;
      if (lane == 0) tm.total = total;
      __syncwarp();
      nbar_arrive(pub_id<ST>(s), kPubCount);
      if (m == 0) { EVT(tile, 7) }
      // whoever completes a block publishes its aggregate (the answer of the atomic is needed only now)
      if (lane == 0) {
        const uint32_t blk = tile >> 5;
        const uint32_t in_blk = min(32u, P.ntiles - (blk << 5));
        if ((uint32_t)(blk_old >> 48) + 1u == in_blk)
          desc_raise(P.desc[1] + 2 * (size_t)blk + m, kFlagAgg | ((blk_old & ((1ull << 48) - 1)) + total));
      }
    }
    {  // statistics of this warp's tiles
      const uint32_t nk = __reduce_add_sync(0xffffffffu, n_kept), nm = __reduce_add_sync(0xffffffffu, n_matched);
      if (lane == 0) {
        if (m == 0 && nk && !P.no_nout) atomicAdd(&P.result->n_out, (unsigned long long)nk);
        if (nm) atomicAdd(&P.result->n_match[m + P.rslot], (unsigned long long)nm);
      }
    }
  } else if (is_scan) {
    // =============================== SCAN =======================================================
    const int sj = warp - NB;
    for (uint32_t k = sj;; k += kScanWarps) {
      const uint32_t s = k % ST;
      nbar_sync(pub_id<ST>(s), kPubCount);
      const uint32_t tile = meta[0][s].tile;
      if (tile == kEndTile2) break;
      if (tile != kEndTile) {
        EVT(tile, 8)
#ifdef BK_EXP_NOLB  // experiment: what the kernel does without the look-back (output offsets wrong, timing only)
        const ulonglong2 g = make_ulonglong2((unsigned long long)tile * R * 400ull, (unsigned long long)tile * R * 400ull);
#else
        const ulonglong2 g = lookback<PAIRED>(P.desc[0], P.desc[1], tile, meta[0][s].total, PAIRED ? meta[NM - 1][s].total : 0u, lane);
#endif
        if (lane == 0) {
          meta[0][s].gout = g.x;
          if (PAIRED) meta[NM - 1][s].gout = g.y;
        }
        EVT(tile, 9)
      }
      __syncwarp();
      nbar_arrive(ready_id<ST>(s), kReadyCount);
      if (tile == kEndTile) break;
    }
  } else {
    // =============================== BACK =======================================================
    uint8_t* const gdst = P.out[m];
    for (uint32_t k = 0;; ++k) {
      const uint32_t s = k % ST;
      const TileMeta& tm = meta[m][s];
      nbar_sync(ready_id<ST>(s), kReadyCount);
      const uint32_t tile = tm.tile;
      if (tile == kEndTile) break;
      if (m == 0 && bj == 0) { EVT(tile, 10) }
      const uint32_t total = tm.total;
      const unsigned long long gout = tm.gout;
      if (lane == 0 && bj == 0 && tile == P.ntiles - 1) P.result->out_len[m + P.rslot] = gout + total;
      const bool fits = gout + total <= P.cap[m];
      if (!fits && lane == 0) atomicOr(&P.result->error, 1u);
      const uint32_t phase = (uint32_t)(gout & 15ull);
      // The output tile may be overwritten once the engine has read the previous tile out of it.  That
      // wait sits here, behind the wait for this tile's offset, rather than behind the store itself.
      if (bj == 0 && lane == 0) bulk_wait_read0();
      back_sync(m, nparts);
      if (fits && total) {
        // ---- assemble the tile's output in shared memory (same 16-byte phase as the global destination)
        const uint32_t ooff = tm.ooff[lane];
        const bool kept = ooff != kNotKept;
        RecView rv;
        rv.beg = tm.beg[lane];
        rv.end = tm.beg[lane + 1];
        const uint32_t hf = tm.head_len_fast[lane], ln = tm.lens[lane];
        rv.head_len = hf & 0x7FFFFFFFu;
        rv.fast = hf >> 31;
        rv.seq = tm.seq[lane];
        rv.seq_len = ln & 0xFFFFu;
        rv.qual = tm.qual[lane];
        rv.qual_len = ln >> 16;
        const int msv = tm.ms[lane];
        const int ms = (VAR && msv >= 0) ? (msv & 0xFFFFFF) : msv;
        const bk_devprog& pe = (VAR && msv >= 0 && P.nvar[m] > 1u) ? P.vprog[m][msv >> 24] : pg;
        const uint32_t dst = out_s + phase + ooff;
        // lane per record, warp per part: part bj of every rewritten record (and of a record in foreign
        // framing: CR line ends, text after '+') through the stream writer
        const bool unc = kept && rv.fast && ms < 0;
        // (staggered long runs when the output records of the tile sit a multiple of 32 bytes apart)
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, ooff, 1);
        const bool even = kept && lane < 31 && nxt != kNotKept && ((nxt - ooff) & 31u) == 0;
        const uint32_t rot = __popc(__ballot_sync(0xffffffffu, even)) >= 8 ? ((uint32_t)lane & 7u) + 1u : 0u;
        if (VAR && TABLE && bk_free_shape(pg)) {
          if (kept && !unc) {
            const EmitProg ep = emit_prog(pg, ms >= 0 ? P.xtab[m][tile * R + lane] : make_uint4(0, 0, 0, 0));
            emit_part(ep, rv, ms, trim, dst, (uint32_t)bj, nparts, rot, kEmitCp && (!PAIRED || BK_EMIT_CP > 1));
          }
        }
#ifndef BK_EXP_NOEMIT  // experiment: the kernel without the assembly of rewritten records (timing only)
        else if (kept && !unc) emit_part(pe, rv, ms, trim, dst, (uint32_t)bj, nparts, rot, kEmitCp && (!PAIRED || BK_EMIT_CP > 1));
#endif
        // warp per pass: maximal runs of consecutive unchanged records
        const unsigned um = __ballot_sync(0xffffffffu, unc);
        if (um) {
          // lane i learns its run [i0, i1]; the run's first lane files the copy descriptor (every back
          // warp of the mate writes the same list) and the runs' ragged ends go to warps 0 and 1
          const unsigned below = ~um & ((1u << lane) - 1u);
          const int i0 = unc ? (below ? 32 - __clz(below) : 0) : lane;
          const unsigned above = ~(um >> lane);  // first zero at or above this lane ends the run
          const int i1 = unc ? (above ? lane + __ffs(above) - 2 : 31) : lane;
          const uint32_t run_s = __shfl_sync(0xffffffffu, rv.beg, i0);
          const uint32_t run_d = __shfl_sync(0xffffffffu, dst, i0);
          const uint32_t run_n = __shfl_sync(0xffffffffu, rv.end, i1) - run_s;
          const bool first = unc && i0 == lane;
          const unsigned starts = __ballot_sync(0xffffffffu, first);
          if (first) run_desc[m][__popc(starts & ((1u << lane) - 1u))] = bulk_desc(run_s, run_d, run_n, 0u);
          if (unc && bj < 2) {
            uint32_t pad = (16u - (run_d & 15u)) & 15u;
            pad = pad < run_n ? pad : run_n;
            const uint32_t t0 = ((run_d + run_n) & ~15u) - run_d;
            const uint32_t tlo = (int32_t)t0 > (int32_t)pad ? t0 : pad;
            const bool head = bj == 0 && first && pad, tail = bj == 1 && i1 == lane && tlo < run_n;
            if (head || tail) {
              if (kEmitCp) {
                cp_run(head ? run_d : run_d + tlo, head ? run_s : run_s + tlo, head ? pad : run_n - tlo);
              } else {
                OutStream o;
                os_begin(o, head ? run_d : run_d + tlo);
                os_run(o, head ? run_s : run_s + tlo, head ? pad : run_n - tlo);
                os_end(o);
              }
            }
          }
          __syncwarp();
          const uint32_t nruns = __popc(starts);
          uint32_t t0 = (uint32_t)bj;  // our first pass of the run; rotates from run to run so that short runs spread
          for (uint32_t r = 0; r < nruns; ++r) {
            bulk16(run_desc[m][r], lane, t0, nparts);
            t0 = t0 == 0 ? nparts - 1u : t0 - 1u;
          }
        }
      }
      __syncwarp();
      if (m == 0 && bj == 0) { EVT(tile, 11) }
      if (m == 0 && bj == 3) { EVT(tile, 15) }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[m][s]);  // ring slot (input tile + meta) is free again
      if (fits && total) fence_proxy_async_smem();
      // all shares written -> one warp stores; nobody overwrites the tile until the store has read it
      back_sync(m, nparts);
      if (m == 0 && bj == 0) { EVT(tile, 12) }
      if (fits && total && bj == 0) {
        // ---- store: aligned middle by TMA, ragged edges by the lanes ---------------------------------
        const unsigned long long g0 = gout, g1 = gout + total;
        const unsigned long long a0 = (g0 + 15ull) & ~15ull, a1 = g1 & ~15ull;
        if (a0 >= a1) {
          for (unsigned long long x = g0 + lane; x < g1; x += 32)
            gdst[x] = (uint8_t)lds8(out_s + phase + (uint32_t)(x - g0));
        } else {
          if (lane == 0) {
            bulk_s2g(gdst + a0, out_s + phase + (uint32_t)(a0 - g0), (uint32_t)(a1 - a0));
            bulk_commit();
          }
          unsigned long long x = g0 + lane;
          if (x < a0) gdst[x] = (uint8_t)lds8(out_s + phase + (uint32_t)(x - g0));
          x = a1 + lane;
          if (x < g1) gdst[x] = (uint8_t)lds8(out_s + phase + (uint32_t)(x - g0));
        }
        __syncwarp();
      }
      if (m == 0 && bj == 0) { EVT(tile, 13) }
    }
    if (bj == 0 && lane == 0) bulk_wait_read0();  // the last store must have left shared memory before the CTA does
  }
}

}  // namespace bk
