// Device building blocks of the extract hot path (sm_100a); the kernel itself is in bk_extract.cuh.
//
// One launch = one batch = parse_se_reads / parse_pe_reads (run.rs:63-80, :163-195) + the writer's
// framing (fastq.rs:259-263) for every record of the batch, single pass over HBM:
//
//   tile  = 32 consecutive records (pairs); lane i of a warp owns record i
//   load  : the tile's contiguous FASTQ text range, one 1-D TMA bulk copy per mate -> shared memory
//   parse : lane finds its record's four lines
//   match : lane evaluates the lowered program (bk_program.h) at ascending starts -> leftmost match
//           (pattern.rs:225-230), optional reverse-complement retry (parse.rs:58-68)
//   size  : output bytes per record, warp scan, decoupled look-back across tiles -> exact global
//           output offset (ordered, gap-free: run.rs:79,194 + fastq.rs:265-271)
//   emit  : lane assembles `@head NAME:seq:qual...\nseq'\n+\nqual'\n` (parse.rs:92-171) in a shared
//           output tile laid out with the global offset's 16-byte phase
//   store : aligned middle by TMA bulk store, ragged head/tail bytes by the lanes
//
// All tile accesses go through explicit 32-bit shared-space addresses (ld.shared / st.shared): the
// compiler cannot always prove that a pointer derived from the dynamic shared array, carried through
// a struct or a non-inlined call, is shared, and then emits generic 64-bit loads/stores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bk_program.h"

namespace bk {

constexpr int kTileRecs = 32;  // records per tile = lanes per warp

struct DevResult {
  unsigned long long out_len[2];
  unsigned long long n_out;
  unsigned long long n_match[2];
  unsigned int error;  // bit0: output capacity exceeded, bit1: tile larger than staging buffer
  unsigned int pad;
};

struct KParams {
  bk_devprog prog[2];
  const uint8_t* text[2];
  const uint32_t* off[2];
  uint8_t* out[2];
  unsigned long long cap[2];
  int32_t* match[2];
  unsigned long long* desc[3];  // look-back state: tile descriptors, block descriptors, block accumulators
  unsigned int* tile_counter;
  DevResult* result;
  unsigned long long* prof;  // 24 phase-cycle sums (only written when built with BK_PROFILE_PHASES)
  uint32_t n;
  uint32_t ntiles;
  uint32_t tile_recs;  // <= 32 (smaller when records are too long for the staging buffer)
  uint32_t text_len[2];  // bytes of FASTQ text per mate
  uint32_t in_cap[2];  // shared-memory bytes for the staged input / assembled output per mate
  uint32_t out_cap[2];
  uint32_t flags;
  // Bounded variable repetition (bk_pattern.cc): a mate's pattern as nvar > 1 fixed-length programs in
  // priority order, in device memory (prog[m] is the first).  Only extract_kernel<.., true, true> reads them.
  uint32_t nvar[2];
  const bk_devprog* vprog[2];
  // Optional: the header length of every read (the compaction of a pattern-less mate uses its partner's as the guess
  // for its own, bk_compact.cuh).
  uint32_t* hlen[2];
  // Optional (single-end launches): per tile 1 << 63 | keep bits of its reads, written once the tile is matched --
  // the hand-over to a compact_kernel running next to this kernel; and the index under which this launch files its
  // mate's figures in `result` (0, or 1 when it works on the second mate of a pair).
  unsigned long long* kmask;
  uint32_t rslot;
  // Optional (single-end launches): match tables left by earlier launches on the stream -- [0] this mate's own
  // (extract_kernel<.., TABLE = true> matches nothing itself), [1] the other mate's of the pair (keep rule); and
  // whether this launch leaves the count of kept pairs to the other mate's launch.
  // A paired launch in TABLE form reads them differently: tab_in[m] is mate m's own table, null for a mate that matches
  // for itself.
  const int32_t* tab_in[2];
  // Programs whose match has no fixed shape (parametric programs, bk_program.h): the spans of every match of mate m as
  // match_kernel found them -- uint4 = 8 x u16: match length, then (offset from the match start, length) of up to three
  // captures, one spare.
  const uint4* xtab[2];
  uint32_t no_nout;
};

constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

// ---- PTX helpers ---------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v));
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v));
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y));
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking
  uint32_t done;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// Blocking wait with back-off.  A bare try_wait loop issues continuously, and so does the
// suspend-time-hint form (NANOSLEEP.SYNCS wakes on every transaction-count update of the CTA):
// either way the pollers took a third of the SM's issue slots from the warps doing the work.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(128);
  }
}
// Named barriers as one-way hand-overs between warps: the producer arrives and runs on, the
// consumers block in hardware (no polling, no issue slots) until the count is reached.
__device__ __forceinline__ void nbar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void nbar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// L2 prefetch of a byte range (16-byte aligned address and size) / of one line
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// 1-D TMA: shared -> global
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_saddr, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_saddr),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- record parsing --------------------------------------------------------------------------------
// Positions are absolute shared-memory byte addresses.

// First '\n' at or after `from`, or `end` when there is none.
__device__ __forceinline__ uint32_t find_nl(uint32_t from, uint32_t end) {
  uint32_t p = from & ~3u;
  uint32_t w = lds32(p) ^ 0x0A0A0A0Au;
  uint32_t m = (w - 0x01010101u) & ~w & 0x80808080u;
  m &= 0xFFFFFFFFu << ((from & 3u) * 8u);
  while (m == 0) {
    p += 4;
    if (p >= end) return end;
    w = lds32(p) ^ 0x0A0A0A0Au;
    m = (w - 0x01010101u) & ~w & 0x80808080u;
  }
  const uint32_t pos = p + ((__ffs(m) - 1) >> 3);
  return pos < end ? pos : end;
}

struct RecView {    // seq_io field definitions: lines without EOL and without a trailing CR
  uint32_t beg;     // '@'
  uint32_t end;     // one past the record's last byte
  uint32_t fast;    // layout is exactly `@head LF seq LF + LF qual LF` (the writer's own framing)
  uint32_t head_len;
  uint32_t seq;
  uint32_t seq_len;
  uint32_t qual;
  uint32_t qual_len;
};

// find_nl for the header line, four words per turn (independent loads: one shared-memory round trip
// per 16 bytes instead of one per word)
__device__ __forceinline__ uint32_t find_nl_head(uint32_t from, uint32_t end) {
  uint32_t p = from & ~3u;
  uint32_t keep = 0xFFFFFFFFu << ((from & 3u) * 8u);
#pragma unroll 1
  while (p < end) {
    uint32_t m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w = lds32(p + 4u * i) ^ 0x0A0A0A0Au;
      m[i] = (w - 0x01010101u) & ~w & 0x80808080u;  // lowest set bit is exact; higher ones may be borrows
    }
    m[0] &= keep;
    keep = 0xFFFFFFFFu;
    uint32_t pos = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 3; i >= 0; --i) pos = m[i] ? p + 4u * i + ((__ffs(m[i]) - 1) >> 3) : pos;
    if (pos != 0xFFFFFFFFu) return pos < end ? pos : end;
    p += 16u;
  }
  return end;
}

// The same searches on text that stays in global memory (a mate that is never staged on chip, see
// extract_kernel's BYP form): positions are byte offsets into `t`, whose base is 16-byte aligned and whose
// allocation extends 16 bytes past the text (include/barkit_b200.h, bk_mate_in).  16 bytes per load.
__device__ __forceinline__ uint32_t nl_mask16(const uint4 v, uint32_t first_valid, uint32_t base) {
  // -> position (base + byte index) of the first '\n' among bytes >= first_valid of the chunk, or 0xFFFFFFFF
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t pos = 0xFFFFFFFFu;
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    const uint32_t x = w[i] ^ 0x0A0A0A0Au;
    uint32_t m = (x - 0x01010101u) & ~x & 0x80808080u;
    const uint32_t lo = 4u * i;  // bytes of this word below first_valid are masked off
    if (first_valid > lo) m &= first_valid - lo >= 4u ? 0u : 0xFFFFFFFFu << (8u * (first_valid - lo));
    pos = m ? base + lo + ((__ffs(m) - 1) >> 3) : pos;
  }
  return pos;
}
__device__ __noinline__ uint32_t find_nl_g(const uint8_t* t, uint32_t from, uint32_t end) {
  uint32_t p = from & ~15u, fv = from & 15u;
#pragma unroll 1
  while (p < end) {
    const uint32_t pos = nl_mask16(__ldg(reinterpret_cast<const uint4*>(t + p)), fv, p);
    if (pos != 0xFFFFFFFFu) return pos < end ? pos : end;
    fv = 0;
    p += 16u;
  }
  return end;
}
// header line: the first 64 bytes in one round trip (four independent loads), longer headers chunk by chunk
__device__ __forceinline__ uint32_t find_nl_head_g(const uint8_t* t, uint32_t from, uint32_t end) {
  const uint32_t p = from & ~15u;
  uint4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = p + 16u * i < end ? __ldg(reinterpret_cast<const uint4*>(t + p + 16u * i)) : make_uint4(0, 0, 0, 0);
  uint32_t pos = 0xFFFFFFFFu;
#pragma unroll
  for (int i = 3; i >= 0; --i) {
    const uint32_t q = nl_mask16(v[i], i == 0 ? (from & 15u) : 0u, p + 16u * i);
    pos = q != 0xFFFFFFFFu ? q : pos;
  }
  if (pos != 0xFFFFFFFFu) return pos < end ? pos : end;
  return p + 64u < end ? find_nl_g(t, p + 64u, end) : end;
}

// where a record's bytes are read from: the staged tile (absolute shared-memory addresses) or global text
struct MemS {
  __device__ __forceinline__ uint32_t b(uint32_t a) const { return lds8(a); }
  __device__ __forceinline__ uint32_t nl(uint32_t from, uint32_t end) const { return find_nl(from, end); }
  __device__ __forceinline__ uint32_t nl_head(uint32_t from, uint32_t end) const { return find_nl_head(from, end); }
};
struct MemG {
  const uint8_t* t;
  __device__ __forceinline__ uint32_t b(uint32_t a) const { return __ldg(t + a); }
  __device__ __forceinline__ uint32_t nl(uint32_t from, uint32_t end) const { return find_nl_g(t, from, end); }
  __device__ __forceinline__ uint32_t nl_head(uint32_t from, uint32_t end) const { return find_nl_head_g(t, from, end); }
};

template <class M>
__device__ __forceinline__ uint32_t strip_cr(const M& mem, uint32_t b, uint32_t e) {
  return (e > b && mem.b(e - 1) == '\r') ? e - 1 : e;
}

// the general case: any line ends, text after '+', a last line without EOL (rare; kept out of line so that
// the common path below stays small in the instruction cache)
template <class M>
__device__ __noinline__ RecView parse_record_slow(const M mem, uint32_t b, uint32_t e, uint32_t nl1) {
  RecView r;
  r.beg = b;
  r.end = e;
  r.fast = 0;
  r.head_len = strip_cr(mem, b + 1, nl1 > b ? nl1 : b + 1) - (b + 1);
  const uint32_t sb = nl1 + 1 < e ? nl1 + 1 : e;
  const uint32_t nl2 = mem.nl(sb, e);
  r.seq = sb;
  r.seq_len = strip_cr(mem, sb, nl2) - sb;
  const uint32_t pb = nl2 + 1 < e ? nl2 + 1 : e;
  const uint32_t nl3 = mem.nl(pb, e);
  const uint32_t qb = nl3 + 1 < e ? nl3 + 1 : e;
  const uint32_t nl4 = mem.nl(qb, e);
  r.qual = qb;
  r.qual_len = strip_cr(mem, qb, nl4) - qb;
  return r;
}

template <class M>
__device__ __forceinline__ RecView parse_record_t(const M& mem, uint32_t b, uint32_t e) {
  const uint32_t nl1 = mem.nl_head(b, e);
  const uint32_t sb = nl1 + 1 < e ? nl1 + 1 : e;
  // Fast path for the writer's own framing `@h LF seq LF + LF qual LF`: then e - sb = 2L + 4.  A record
  // of any other shape (CR line ends, text after '+', missing final LF) cannot pass the probes with
  // the L this formula gives (newlines only exist at the line ends), so it takes the scanning path.
  // All probes are fetched together: one memory round trip, not one per condition.
  const uint32_t rem = e - sb;
  const uint32_t L = (rem - 4u) >> 1;
  if (rem >= 4u && !(rem & 1u) && nl1 > b) {
    const uint32_t c0 = mem.b(sb + L), c1 = mem.b(sb + L + 1u), c2 = mem.b(sb + L + 2u), c3 = mem.b(e - 1u);
    const uint32_t c4 = mem.b(sb + L - 1u), c5 = mem.b(e - 2u), c6 = mem.b(nl1 - 1u);  // (c4 is the header's LF when L == 0)
    const bool framed = (c0 == '\n') & (c1 == '+') & (c2 == '\n') & (c3 == '\n');
    const bool cr = (c4 == '\r') | (c5 == '\r');
    if (framed && (L == 0 || !cr)) {
      RecView r;
      r.beg = b;
      r.end = e;
      const bool hcr = nl1 > b + 1u && c6 == '\r';  // CR on the header line: stripped, but then not the writer's framing
      r.head_len = nl1 - (b + 1u) - (hcr ? 1u : 0u);
      r.seq = sb;
      r.seq_len = L;
      r.qual = sb + L + 3u;
      r.qual_len = L;
      r.fast = !hcr;
      return r;
    }
  }
  return parse_record_slow<M>(mem, b, e, nl1);
}
__device__ __noinline__ RecView parse_record(uint32_t b, uint32_t e) { return parse_record_t(MemS{}, b, e); }
__device__ __noinline__ RecView parse_record_g(const uint8_t* t, uint32_t b, uint32_t e) { return parse_record_t(MemG{t}, b, e); }

// ---- matching --------------------------------------------------------------------------------------

// parse.rs:12-32 TRANSLATION_TABLE
__device__ __forceinline__ uint32_t complement(uint32_t c) {
  switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'G': return 'C';
    case 'C': return 'G';
    case 'R': case 'Y': case 'S': case 'W': case 'K': case 'M':
    case 'B': case 'D': case 'H': case 'V': case 'N':
      return c;
    default: return 'A';
  }
}

// `tab` is this mate's 512-byte lookup block in shared memory: [0,256) class-membership bits per
// byte, [256,512) the complement table (a switch here was a quarter of the kernel's code).
// Behind the two byte tables sits a copy of the matching program itself: read through a reference
// into the kernel's parameter space from a non-inlined function, every field access was a
// generic-address load with global-memory latency.
constexpr uint32_t kTabComp = 256, kTabHdr = 512, kTabOps = 528, kTabLit = kTabOps + 8u * BK_MAX_OPS;
constexpr uint32_t kTabSA = kTabLit + BK_MAX_LIT;  // 8 x 2 words: bit planes of the program's first literal (match_unanchored)
constexpr uint32_t kTabIns = kTabSA + 64u;  // nins, pad, then BK_MAX_INS x bk_ins (8 bytes each): parametric programs (match_par)
constexpr uint32_t kTabBytes = kTabIns + 16u + 8u * BK_MAX_INS;  // (a multiple of 16: the blocks of the two mates sit back to back)
static_assert(kTabBytes % 16u == 0, "lookup blocks are read with 16-byte loads");
struct MOp { uint32_t off, len, kind, k, idx; };
__device__ __forceinline__ MOp tab_op(uint32_t tab, uint32_t o) {  // bk_op, little endian
  const uint2 v = lds64(tab + kTabOps + 8u * o);
  MOp op;
  op.off = v.x & 0xFFFFu; op.len = v.x >> 16;
  op.kind = v.y & 0xFFu; op.k = (v.y >> 8) & 0xFFu; op.idx = v.y >> 16;
  return op;
}

template <bool RC>
__device__ __forceinline__ uint32_t fetch(uint32_t tab, uint32_t seq, uint32_t L, uint32_t j) {
  if (RC) return lds8(tab + kTabComp + lds8(seq + L - 1 - j));
  return lds8(seq + j);
}

// All compare runs of the program at start s.  Branch-free inside a run so that the lanes of a warp
// (one read each) stay in lock step: a data-dependent early exit would split the warp and replay
// the remaining runs once per split.  `tab[c]` = bit k set iff class k contains byte c (0 for
// c >= 0x80).  EARLY_OUT (unanchored search, lanes already diverge on s) stops at the first
// failed run.
// Forward strand, four bytes at a time: the run is read as unaligned words (two aligned loads and
// a funnel shift each; the loads do not depend on data, so they pipeline), a literal run is XORed
// with the program's word-packed literal and the non-zero bytes counted (= Hamming distance,
// pattern.rs:40-79), a class run looks its bytes up in `tab`.  A mismatching byte sits in a
// wildcard slot of one of the alternatives, i.e. under the regex `.`, which must be ASCII here
// ('\n' cannot occur inside a line).
#ifndef BK_MATCH_UNROLL
#define BK_MATCH_UNROLL 1
#endif
constexpr int kMatchUnroll = BK_MATCH_UNROLL;  // words of a compare run per turn of the matcher's loops
template <bool EARLY_OUT>
__device__ __forceinline__ bool match_at_fwd(uint32_t tab, uint32_t seq, uint32_t s) {
  bool ok = true;
  const uint32_t nops = lds32(tab + kTabHdr);
  for (uint32_t o = 0; o < nops; ++o) {
    const MOp op = tab_op(tab, o);
    const uint32_t base = seq + s + op.off;
    const uint32_t p = base & ~3u, sh = (base & 3u) * 8u;
    const uint32_t nw = (op.len + 3u) >> 2;
    const uint32_t tail = op.len & 3u;  // valid bytes of the last word (0 = all four)
    uint32_t prev = lds32(p);
    if (op.kind == BK_OP_LIT) {
      uint32_t mism = 0, bad = 0;
      const uint32_t lit = tab + kTabLit + op.idx;  // literals start on words (bk_pattern.cc)
#pragma unroll kMatchUnroll
      for (uint32_t i = 0; i < nw; ++i) {
        const uint32_t nxt = lds32(p + 4u * i + 4u);
        const uint32_t v = __funnelshift_r(prev, nxt, sh);
        prev = nxt;
        uint32_t x = v ^ lds32(lit + 4u * i);
        if (i + 1 == nw && tail) x &= (1u << (8u * tail)) - 1u;
        const uint32_t nz = (((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;  // 0x80 per differing byte
        mism += __popc(nz);
        bad |= nz & v;  // differing byte >= 0x80
      }
      ok = ok && (mism <= op.k) && !bad;
    } else {
      uint32_t acc = 0xFFu;
#pragma unroll kMatchUnroll
      for (uint32_t i = 0; i < nw; ++i) {
        const uint32_t nxt = lds32(p + 4u * i + 4u);
        uint32_t v = __funnelshift_r(prev, nxt, sh);
        prev = nxt;
        if (i + 1 == nw && tail) v = __byte_perm(v, 0, tail == 1 ? 0x0000 : tail == 2 ? 0x0010 : 0x0210);  // repeat byte 0
        acc &= lds8(tab + (v & 0xFFu)) & lds8(tab + ((v >> 8) & 0xFFu)) & lds8(tab + ((v >> 16) & 0xFFu)) &
               lds8(tab + (v >> 24));
      }
      ok = ok && ((acc >> op.idx) & 1u);
    }
    if (EARLY_OUT && !ok) return false;
  }
  return ok;
}

template <bool RC, bool EARLY_OUT>
__device__ __forceinline__ bool match_at(uint32_t tab, uint32_t seq, uint32_t L, uint32_t s) {
  if (!RC) return match_at_fwd<EARLY_OUT>(tab, seq, s);
  bool ok = true;
  const uint32_t nops = lds32(tab + kTabHdr);
  for (uint32_t o = 0; o < nops; ++o) {
    const MOp op = tab_op(tab, o);
    const uint32_t base = s + op.off;
    if (op.kind == BK_OP_LIT) {
      uint32_t mism = 0, bad = 0;
#pragma unroll 1
      for (uint32_t i = 0; i < op.len; ++i) {
        const uint32_t c = fetch<RC>(tab, seq, L, base + i);
        const uint32_t ne = (c != lds8(tab + kTabLit + op.idx + i));
        mism += ne;
        // a wildcard slot is the regex `.`: any ASCII byte except '\n' (pattern.rs:10,71)
        bad |= ne & (uint32_t)((c >= 0x80u) | (c == '\n'));
      }
      ok = ok && (mism <= op.k) && !bad;
    } else {
      uint32_t acc = 0xFFu;
#pragma unroll 1
      for (uint32_t i = 0; i < op.len; ++i) acc &= lds8(tab + fetch<RC>(tab, seq, L, base + i));
      ok = ok && ((acc >> op.idx) & 1u);
    }
    if (EARLY_OUT && !ok) return false;
  }
  return ok;
}

// One out-of-line copy of the forward matcher for the general kernels, whose hot code is past the instruction
// cache: the anchored mate's match and the verification step of the unanchored search
// share it instead of carrying an inlined copy each.
__device__ __noinline__ bool match_at_fwd_shared(uint32_t tab, uint32_t seq, uint32_t s) {
  return match_at_fwd<false>(tab, seq, s);
}
// match_read<false> on top of it
__device__ __noinline__ int match_read_shared(uint32_t tab, uint32_t seq, uint32_t L) {
  const uint4 h = lds128(tab + kTabHdr);  // nops, total_len, anchor_start, anchor_end
  const uint32_t T = h.y;
  if (L < T || !(h.z || h.w)) return -1;
  if (h.z && h.w && L != T) return -1;
  const uint32_t s = h.w ? L - T : 0;
  return match_at_fwd_shared(tab, seq, s) ? (int)s : -1;
}

// leftmost match start or -1, one lane per read.  Forward strand:
// anchored patterns only (one candidate start); the unanchored forward search is match_unanchored.
template <bool RC>
__device__ __noinline__ int match_read(uint32_t tab, uint32_t seq, uint32_t L) {
  const uint4 h = lds128(tab + kTabHdr);  // nops, total_len, anchor_start, anchor_end
  const uint32_t T = h.y;
  if (L < T) return -1;
  if (h.z || h.w) {  // exactly one candidate start
    if (h.z && h.w && L != T) return -1;
    const uint32_t s = h.w ? L - T : 0;
    return match_at<RC, false>(tab, seq, L, s) ? (int)s : -1;
  }
  if (RC) {  // (reverse-complement retry of an unanchored pattern: rare, one lane per read)
    for (uint32_t s = 0; s <= L - T; ++s)
      if (match_at<RC, true>(tab, seq, L, s)) return (int)s;
  }
  return -1;
}

// parse.rs:166 `from_utf8(barcode_seq)?`: a capture that is not well-formed UTF-8 makes create_read_with_new_header
// fail, and the read counts as unmatched (parse.rs:74-79 `.ok()?`).  A forward match cannot capture such bytes (here
// a byte >= 0x80 matches nothing), but a reverse-complement match is found on the complemented text -- where every
// byte is a letter -- and its offsets are then applied to the FORWARD read (parse.rs:58-68), so it can.
// Well-formed = Unicode table 3-7; the capture is validated on its own, so a sequence cut off by its end is ill-formed.
__device__ __noinline__ bool span_utf8_ok(uint32_t b, uint32_t e) {
  {
    for (uint32_t i = b; i < e;) {
      const uint32_t x = lds8(i);
      if (x < 0x80u) { ++i; continue; }
      uint32_t need, lo = 0x80u, hi = 0xBFu;
      if (x >= 0xC2u && x <= 0xDFu) need = 1;
      else if (x == 0xE0u) { need = 2; lo = 0xA0u; }
      else if (x == 0xEDu) { need = 2; hi = 0x9Fu; }
      else if (x >= 0xE1u && x <= 0xEFu) need = 2;
      else if (x == 0xF0u) { need = 3; lo = 0x90u; }
      else if (x == 0xF4u) { need = 3; hi = 0x8Fu; }
      else if (x >= 0xF1u && x <= 0xF3u) need = 3;
      else return false;
      if (i + need >= e) return false;
      for (uint32_t j = 1; j <= need; ++j) {
        const uint32_t y = lds8(i + j);
        if (y < lo || y > hi) return false;
        lo = 0x80u; hi = 0xBFu;
      }
      i += need + 1u;
    }
  }
  return true;
}
__device__ __forceinline__ bool caps_utf8_ok(const bk_devprog& pg, uint32_t seq, uint32_t ms) {
  for (uint32_t c = 0; c < pg.ncaps; ++c) {
    const uint32_t b = seq + ms + pg.caps[c].off;
    if (!span_utf8_ok(b, b + pg.caps[c].len)) return false;
  }
  return true;
}
// capture c of a parametric program under the extras `ex` (16 bits per insertion): insertions in front of the group
// move it, insertions inside it lengthen it
__device__ __forceinline__ void par_cap(const bk_devprog& pg, uint32_t c, unsigned long long ex, uint32_t& off, uint32_t& len) {
  off = pg.caps[c].off;
  len = pg.caps[c].len;
  for (uint32_t i = 0; i < pg.nins; ++i) {
    const uint32_t e = (uint32_t)(ex >> (16u * i)) & 0xFFFFu;
    if (i < pg.cap_ins_before[c]) off += e;
    else if ((pg.ins[i].capmask >> c) & 1u) len += e;
  }
}
// what emit_part and out_bytes read of a program; for a parametric one, the record's own figures
struct EmitProg {
  uint32_t ncaps, hdr_growth, total_len;
  bk_cap caps[BK_MAX_CAPS];
};
// the spans of one match, packed as KParams::xtab holds them
__device__ __forceinline__ uint4 pack_spans(uint32_t len, const uint32_t* off, const uint32_t* clen) {
  return make_uint4(len | off[0] << 16, clen[0] | off[1] << 16, clen[1] | off[2] << 16, clen[2]);
}
__device__ __forceinline__ EmitProg emit_prog(const bk_devprog& pg, const uint4 sp) {
  EmitProg ep;
  ep.ncaps = pg.ncaps;
  ep.total_len = sp.x & 0xFFFFu;
  const uint32_t off[3] = {sp.x >> 16, sp.y >> 16, sp.z >> 16}, clen[3] = {sp.y & 0xFFFFu, sp.z & 0xFFFFu, sp.w & 0xFFFFu};
  ep.hdr_growth = 0;
#pragma unroll
  for (uint32_t c = 0; c < BK_MAX_CAPS; ++c) {
    ep.caps[c] = pg.caps[c];
    ep.caps[c].off = (uint16_t)off[c];
    ep.caps[c].len = (uint16_t)clen[c];
    if (c < pg.ncaps) ep.hdr_growth += 3u + pg.caps[c].name_len + 2u * clen[c];  // parse.rs:161-168
  }
  return ep;
}

// ---- variable repetition: a program given as a list of fixed-length variants ------------------------------
//
// Leftmost-first over the variants = smallest start at which any variant matches, and among the variants
// matching there the first in the list (the regex crate's backtracking order, bk_pattern.cc).  The variants
// live in global memory and are evaluated bytewise, one lane per read: this path widens what is accepted;
// it is not tuned.
template <bool RC>
__device__ __forceinline__ bool match_at_variant(const bk_devprog& pv, uint32_t tab, uint32_t seq, uint32_t L, uint32_t s) {
  for (uint32_t o = 0; o < pv.nops; ++o) {
    const bk_op op = pv.ops[o];
    const uint32_t base = s + op.off;
    if (op.kind == BK_OP_LIT) {
      uint32_t mism = 0, bad = 0;
      for (uint32_t i = 0; i < op.len; ++i) {
        const uint32_t c = fetch<RC>(tab, seq, L, base + i);
        const uint32_t ne = (c != pv.lit[op.idx + i]);
        mism += ne;
        bad |= ne & (uint32_t)((c >= 0x80u) | (c == '\n'));
      }
      if (mism > op.k || bad) return false;
    } else {
      for (uint32_t i = 0; i < op.len; ++i) {
        const uint32_t c = fetch<RC>(tab, seq, L, base + i);
        if (c >= 128u || !((pv.cls[op.idx][c >> 5] >> (c & 31u)) & 1u)) return false;
      }
    }
  }
  return true;
}

// -> match start (or -1) and the variant that matched, packed as start | variant << 24
template <bool RC>
__device__ __noinline__ int match_variants(const bk_devprog* vp, uint32_t nvar, uint32_t tab, uint32_t seq, uint32_t L) {
  uint32_t best = 0xFFFFFFFFu;
  for (uint32_t v = 0; v < nvar; ++v) {
    const bk_devprog& pv = vp[v];
    const uint32_t T = pv.total_len;
    if (L < T) continue;
    uint32_t s0 = 0, s1 = L - T;  // candidate starts
    if (pv.anchor_start && pv.anchor_end) { if (L != T) continue; }
    else if (pv.anchor_start) s1 = 0;
    else if (pv.anchor_end) s0 = s1;
    for (uint32_t s = s0; s <= s1 && s < (best & 0xFFFFFFu); ++s)  // a later variant only wins with a smaller start
      if (match_at_variant<RC>(pv, tab, seq, L, s)) { best = s | (v << 24); break; }
  }
  return (int)best;  // 0xFFFFFFFF = -1: no variant matches
}

// ---- variable repetition without a small bound: a parametric program (bk_program.h, bk_ins) ----------------
//
// The pattern is S_0 X_0{e_0} S_1 X_1{e_1} .. S_Q with fixed pieces S_j (the ops of segment j, laid out with every
// atom at its minimum count) and e_i in [0, max_extra_i] further copies of the one-byte atom X_i.  The regex crate's
// leftmost-first match is the backtracking one: the smallest start s, and for it the first (e_0, .., e_{Q-1}) in
// priority order -- e_0 outermost, each from its longest run down (greedy) or from 0 up (lazy) -- under which all
// pieces match (and the match ends at the read's end when the pattern is `$`-anchored).  One lane per read, bytewise
// (fetch<RC>: also on the reverse complement), depth-first with an explicit stack: a piece is compared once the
// extras in front of it are chosen, so a failing piece prunes everything behind it.  Returns the start or -1 and the
// chosen extras, 16 bits each.
template <bool RC>
__device__ __forceinline__ bool par_segment_ok(uint32_t tab, uint32_t seq, uint32_t L, uint32_t o0, uint32_t o1, uint32_t base) {
  for (uint32_t o = o0; o < o1; ++o) {
    const MOp op = tab_op(tab, o);
    const uint32_t b = base + op.off;
    if (op.kind == BK_OP_LIT) {
      uint32_t mism = 0, bad = 0;
      for (uint32_t i = 0; i < op.len; ++i) {
        const uint32_t c = fetch<RC>(tab, seq, L, b + i);
        const uint32_t ne = (c != lds8(tab + kTabLit + op.idx + i));
        mism += ne;
        bad |= ne & (uint32_t)((c >= 0x80u) | (c == '\n'));
      }
      if (mism > op.k || bad) return false;
    } else {
      for (uint32_t i = 0; i < op.len; ++i)
        if (!((lds8(tab + fetch<RC>(tab, seq, L, b + i)) >> op.idx) & 1u)) return false;
    }
  }
  return true;
}

// When the piece behind an insertion begins with a short literal (the usual shape: `[ACGT]{8,}atgccat`), trying every
// e costs a literal compare each -- up to the read's length of them.  Instead the bit-parallel filter of the unanchored
// search (bs_planes, below) marks, 32 text positions at a time, where that literal can sit within its mismatch budget,
// and only those e are tried: the nearest marked position at or below `hi` (greedy) or at or above `lo` (lazy).
// `lit0` = shared address of the text position the literal has with e = 0.  Forward strand only.
template <int NW>
__device__ __forceinline__ void bs_planes(uint32_t addr, uint32_t& p0, uint32_t& p1);
__device__ __forceinline__ bool bs_eligible(uint32_t kind, uint32_t len, uint32_t k);
constexpr uint32_t kNoExtra = 0xFFFFFFFFu;
__device__ __noinline__ uint32_t par_next_lit(uint32_t tab, uint32_t op_len, uint32_t op_k, uint32_t op_idx, uint32_t lit0,
                                              uint32_t lo, uint32_t hi, uint32_t lazy) {
  uint32_t w = lazy ? lo : (hi >= 31u && hi - 31u > lo ? hi - 31u : lo);  // first window: [w, w + 31]
  while (true) {
    uint32_t a0, a1, b0, b1;
    bs_planes<8>(lit0 + w, a0, a1);
    bs_planes<2>(lit0 + w + 32u, b0, b1);
    uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < op_len; ++i) {
      const uint32_t c = lds8(tab + kTabLit + op_idx + i);
      const uint32_t x0 = (c & 2u) ? 0xFFFFFFFFu : 0u, x1 = (c & 4u) ? 0xFFFFFFFFu : 0u;
      const uint32_t mm = (__funnelshift_r(a0, b0, i) ^ x0) | (__funnelshift_r(a1, b1, i) ^ x1);
      const uint32_t carry = c0 & mm;
      c0 ^= mm;
      c2 |= c1 & carry;
      c1 ^= carry;
    }
    uint32_t cand = ~c2 & (op_k == 0u ? ~(c0 | c1) : op_k == 1u ? ~c1 : op_k == 2u ? ~(c1 & c0) : 0xFFFFFFFFu);
    if (hi - w < 31u) cand &= (2u << (hi - w)) - 1u;  // (w >= lo always)
    if (cand) return lazy ? w + (uint32_t)__ffs(cand) - 1u : w + 31u - (uint32_t)__clz(cand);
    if (lazy) {
      if (hi - w < 32u) return kNoExtra;
      w += 32u;
    } else {
      if (w == lo) return kNoExtra;
      w = w >= 32u && w - 32u > lo ? w - 32u : lo;  // (windows may overlap at the low end: harmless)
    }
  }
}

template <bool RC>
__device__ __noinline__ int match_par(uint32_t tab, uint32_t seq, uint32_t L, unsigned long long* extras) {
  const uint4 h = lds128(tab + kTabHdr);  // nops, total_len (minimum), anchor_start, anchor_end
  const uint32_t nops = h.x, T = h.y, nins = lds32(tab + kTabIns);
  *extras = 0ull;
  if (L < T) return -1;
  const uint32_t s_last = h.z ? 0u : L - T;
  // Unanchored: a start can only match if the program's first short literal sits in reach of it -- between its offset
  // in the minimal layout (fA) and that plus everything the insertions in front of it may add (fB).  The literal's
  // possible positions come from the same bit-parallel filter, nearest first; starts out of reach are skipped, and
  // when no position is left the read has no match.
  bool filt = false;
  MOp fop;
  fop.len = fop.k = fop.idx = 0;
  uint32_t fA = 0, fB = 0, nextp = 0;
  bool havep = false;
  if (!RC && !h.z) {
    for (uint32_t o = 0; o < nops && !filt; ++o) {
      const MOp op = tab_op(tab, o);
      if (!bs_eligible(op.kind, op.len, op.k)) continue;
      uint32_t reach = 0;
      for (uint32_t i = 0; i < nins; ++i) {
        const uint2 in = lds64(tab + kTabIns + 8u + 8u * i);
        if ((in.y >> 24) <= o) reach += min(in.x >> 16, L);  // insertion i lies in front of op o
      }
      fop = op;
      fA = op.off;
      fB = op.off + min(reach, L);
      filt = true;
    }
  }
  for (uint32_t s = 0; s <= s_last; ++s) {
    if (filt) {
      if (!havep || nextp < s + fA) {
        if (s + fA + fop.len > L) return -1;
        nextp = par_next_lit(tab, fop.len, fop.k, fop.idx, seq, s + fA, L - fop.len, 1u);
        havep = true;
        if (nextp == kNoExtra) return -1;
      }
      if (nextp > s + fB) {
        s = nextp - fB;
        if (s > s_last) return -1;
      }
    }
    uint32_t e[BK_MAX_INS], r[BK_MAX_INS];
    uint32_t shift = 0, d = 0;
    if (!par_segment_ok<RC>(tab, seq, L, 0u, nins ? (lds32(tab + kTabIns + 8u + 4u) >> 24) : nops, s)) continue;
    bool enter = true, found = false;
    while (true) {
      if (enter) {
        if (d == nins) {
          if (!h.w || s + T + shift == L) { found = true; break; }
          enter = false;  // `$`: the match must end with the read
          if (d == 0) break;
          --d;
          shift -= e[d];
          continue;
        }
        // longest run of the atom at this insertion, as far as the rest of the minimum still fits the read
        const uint2 in = lds64(tab + kTabIns + 8u + 8u * d);  // pos | max_extra << 16, cls | lazy << 8 | capmask << 16 | op_end << 24
        const uint32_t a = s + (in.x & 0xFFFFu) + shift, room = L - (s + T + shift);
        const uint32_t lim = min(in.x >> 16, room), cls = in.y & 0xFFu;
        uint32_t run = 0;
        if (!RC) {  // four bytes per step while whole words are in the class (an unbounded run is most of the read)
          const uint32_t p0 = seq + a;
          while (run + 4u <= lim) {
            const uint32_t q = p0 + run, v = __funnelshift_r(lds32(q & ~3u), lds32((q & ~3u) + 4u), (q & 3u) * 8u);
            const uint32_t m4 = lds8(tab + (v & 0xFFu)) & lds8(tab + ((v >> 8) & 0xFFu)) & lds8(tab + ((v >> 16) & 0xFFu)) & lds8(tab + (v >> 24));
            if (!((m4 >> cls) & 1u)) break;
            run += 4u;
          }
        }
        while (run < lim && ((lds8(tab + fetch<RC>(tab, seq, L, a + run)) >> cls) & 1u)) ++run;
        r[d] = run;
        e[d] = ((in.y >> 8) & 0xFFu) ? 0u : run;
      }
      const uint32_t o0 = lds32(tab + kTabIns + 8u + 8u * d + 4u) >> 24;
      const uint32_t o1 = d + 1u < nins ? lds32(tab + kTabIns + 8u + 8u * (d + 1u) + 4u) >> 24 : nops;
      const uint32_t lazy = (lds32(tab + kTabIns + 8u + 8u * d + 4u) >> 8) & 0xFFu;
      // the choices at level d in priority order: every e in [0, run], or -- when the piece behind the insertion begins
      // with a short literal -- only those at which that literal can sit (par_next_lit)
      MOp lop;
      lop.kind = 0xFFu;
      if (!RC && o0 < o1) lop = tab_op(tab, o0);
      const bool guided = !RC && o0 < o1 && r[d] >= 8u && bs_eligible(lop.kind, lop.len, lop.k);
      bool have;
      if (guided) {
        const uint32_t lit0 = seq + s + shift + lop.off;
        uint32_t nx;
        if (enter) nx = par_next_lit(tab, lop.len, lop.k, lop.idx, lit0, 0u, r[d], lazy);
        else if (lazy) nx = e[d] < r[d] ? par_next_lit(tab, lop.len, lop.k, lop.idx, lit0, e[d] + 1u, r[d], 1u) : kNoExtra;
        else nx = e[d] > 0u ? par_next_lit(tab, lop.len, lop.k, lop.idx, lit0, 0u, e[d] - 1u, 0u) : kNoExtra;
        have = nx != kNoExtra;
        if (have) e[d] = nx;
      } else if (enter) {
        have = true;
      } else {  // next choice at level d
        have = lazy ? e[d] < r[d] : e[d] > 0u;
        if (have) e[d] = lazy ? e[d] + 1u : e[d] - 1u;
      }
      if (!have) {  // give the level up
        enter = false;
        if (d == 0) break;
        --d;
        shift -= e[d];
        continue;
      }
      // the piece behind insertion d under this choice
      if (par_segment_ok<RC>(tab, seq, L, o0, o1, s + shift + e[d])) {
        shift += e[d];
        ++d;
        enter = true;
      } else {
        enter = false;
      }
    }
    if (found) {
      unsigned long long x = 0;
      for (uint32_t i = 0; i < nins; ++i) x |= (unsigned long long)e[i] << (16u * i);
      *extras = x;
      return (int)s;
    }
  }
  return -1;
}

// Unanchored forward search (pattern.rs:225-230: the leftmost start at which the whole program passes).
//
// The program's first run is its most selective one (bk_pattern.cc puts literals first).  When it is a literal of up
// to 8 bytes with k <= 3, every lane scans ITS read bit-parallel, 32 candidate starts per step:
//   * the text is reduced to two bit planes, bit 1 and bit 2 of every byte (A, C, G, T differ in them; every other
//     byte falls onto one of the four), 32 bytes per 32-bit word: four bytes at a time, `(w >> 1) & 0x01010101`
//     gathered into a nibble with one multiply;
//   * for literal position i the planes shifted by i (a funnel shift over two words) are compared with the
//     literal's two bits, which gives the mismatch mask of position i for 32 starts at once;
//   * the masks are summed in three bit-sliced counter words (half adders, the top one sticky), and `count <= k` is
//     a word of candidate starts.
// The reduction can only lose mismatches, so the candidates are a superset of the literal's k-mismatch occurrences;
// each lane verifies its candidates against the whole program (match_at_fwd_shared), lowest first, all lanes of the
// warp together.  ~6.5 instructions per candidate start for the 32 reads of a warp.  The previous form -- one warp per
// read, 32 starts per step, ~35 instructions per step of ONE read -- was what bounded BASELINE config C4; a byte-serial
// shift-add scan (one 4-bit counter per alignment) was also built and measured: ~13 instructions per text byte, slower.
//
// Two warps can share a search (`nwarps` = 2: the front warps of the two mates, when only one mate searches): they take
// alternate windows of 32 starts, exchange their windows' results through `res` behind the pair's named barrier
// `bid`, and both go on to the next two windows while any read is undecided -- the leftmost match of a read is the
// one in its lowest window.  Programs whose first run is not such a literal are swept one warp per read as before.
// Called by all lanes together; `can` = this lane's read takes part; returns this lane's read's result.
template <int NW>
__device__ __forceinline__ void bs_planes(uint32_t addr, uint32_t& p0, uint32_t& p1) {  // bit planes of the 4 * NW bytes at addr
  const uint32_t base = addr & ~3u, sh = (addr & 3u) * 8u;
  uint32_t prev = lds32(base);
  p0 = p1 = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const uint32_t nxt = lds32(base + 4u * w + 4u);
    const uint32_t v = __funnelshift_r(prev, nxt, sh);
    prev = nxt;
    p0 |= ((((v >> 1) & 0x01010101u) * 0x01020408u) >> 24) << (4 * w);
    p1 |= ((((v >> 2) & 0x01010101u) * 0x01020408u) >> 24) << (4 * w);
  }
}

__device__ __forceinline__ bool bs_eligible(uint32_t kind, uint32_t len, uint32_t k) { return kind == BK_OP_LIT && len <= 8u && k <= 3u; }

__device__ __noinline__ int match_unanchored(uint32_t tab, uint32_t seq, uint32_t L, bool can, int lane, int which, int nwarps,
                                              int* res, uint32_t bid) {
  const uint32_t T = lds32(tab + kTabHdr + 4u);
  const MOp op0 = tab_op(tab, 0);
  int ms = -1;
  bool pend = can && L >= T;
  if (bs_eligible(op0.kind, op0.len, op0.k)) {
    const uint32_t last = pend ? L - T : 0u;  // this read's last candidate start
    const uint32_t smax = __reduce_max_sync(0xffffffffu, last);
    if (!__any_sync(0xffffffffu, pend)) return -1;
    const uint32_t nwin = smax / 32u + 1u;
    for (uint32_t r = 0; r * (uint32_t)nwarps < nwin; ++r) {
      const uint32_t w = r * (uint32_t)nwarps + (uint32_t)which;
      int mine = -1;
      if (w < nwin) {
        const uint32_t q = 32u * w, cb = seq + q + op0.off;
        // (unrolled: a rolled, out-of-line form is 1 KB smaller and measured slower, 6.35 against 6.15 ms on C4)
        uint32_t a0, a1, b0, b1;
        bs_planes<8>(cb, a0, a1);
        bs_planes<2>(cb + 32u, b0, b1);  // (the literal reaches at most 7 bytes into the next 32)
        uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < op0.len; ++i) {
          const uint2 x = lds64(tab + kTabSA + 8u * i);
          const uint32_t mm = (__funnelshift_r(a0, b0, i) ^ x.x) | (__funnelshift_r(a1, b1, i) ^ x.y);
          const uint32_t carry = c0 & mm;
          c0 ^= mm;
          c2 |= c1 & carry;
          c1 ^= carry;
        }
        uint32_t cand = ~c2 & (op0.k == 0u ? ~(c0 | c1) : op0.k == 1u ? ~c1 : op0.k == 2u ? ~(c1 & c0) : 0xFFFFFFFFu);
        if (!pend || q > last) cand = 0;
        else if (last - q < 31u) cand &= (2u << (last - q)) - 1u;
        while (__any_sync(0xffffffffu, cand != 0u)) {  // every lane's lowest candidate, verified together
          const bool have = cand != 0u;
          const uint32_t st = q + (have ? (uint32_t)__ffs(cand) - 1u : 0u);
          if (have && match_at_fwd_shared(tab, seq, st)) {
            mine = (int)st;
            cand = 0;
          } else if (have) {
            cand &= cand - 1u;
          }
        }
      }
      if (nwarps == 2) {  // the other warp's window of this round
        int* const rr = res + (r & 1u) * 64u;
        rr[which * 32 + lane] = mine;
        nbar_sync(bid, 64);
        const int ra = rr[lane], rb = rr[32 + lane];
        mine = ra >= 0 ? ra : rb;
      }
      if (pend && mine >= 0) {
        ms = mine;
        pend = false;
      }
      if (!__any_sync(0xffffffffu, pend)) break;
    }
    return ms;
  }
  // general first run: one warp per read, the 32 lanes test 32 consecutive starts against the whole program
  for (unsigned todo = __ballot_sync(0xffffffffu, pend); todo; todo &= todo - 1u) {
    const int r = __ffs(todo) - 1;
    const uint32_t rseq = __shfl_sync(0xffffffffu, seq, r), last = __shfl_sync(0xffffffffu, L, r) - T;
    int found = -1;
    for (uint32_t s0 = 0; s0 <= last; s0 += 32u) {
      const uint32_t s = s0 + (uint32_t)lane;
      const bool ok = s <= last && match_at_fwd_shared(tab, rseq, s);
      const unsigned hit = __ballot_sync(0xffffffffu, ok);
      if (hit) {
        found = (int)s0 + __ffs(hit) - 1;
        break;
      }
    }
    if (lane == r) ms = found;
  }
  return ms;
}

// ---- emit ----------------------------------------------------------------------------------------------

// output bytes of a kept record (qual_len == seq_len whenever ms >= 0, see emit_part)
template <class PG>
__device__ __forceinline__ uint32_t out_bytes(const PG& pg, const RecView& r, int ms, bool trim) {
  uint32_t sl = r.seq_len, ql = r.qual_len, g = 0;
  if (ms >= 0) {
    g = pg.hdr_growth;  // sum of " NAME:" seq ":" qual   (parse.rs:161-168)
    if (trim) {
      sl -= pg.total_len;  // the match lies inside seq
      ql -= pg.total_len;
    }
  }
  return 1 + r.head_len + g + 1 + sl + 3 + ql + 1;
}

// Per-lane output stream into the shared output tile.  Bytes are produced strictly in order; the
// lane keeps the (up to 3) bytes that do not yet fill a word in a register and stores whole aligned
// words.  A stream may start and end anywhere inside a word; those two words are shared with other
// writers, so only the stream's own bytes of them are stored (bytewise).  Everything is predicated
// rather than branched: the lanes of a warp (one record each) are at different alignments.
struct OutStream {
  uint32_t wpos;  // shared address of the word being filled (4-byte aligned)
  uint32_t lo;    // its bytes so far: low `nb` bytes meaningful (the first `skip` are placeholders), rest zero
  uint32_t nb;    // 0..3
  uint32_t skip;  // leading bytes of the word at wpos that belong to another writer (first word only)
  uint32_t rot;   // 0, or 1 + the 16-byte step at which this lane starts a long run (see os_run_f)
};

__device__ __forceinline__ void os_begin(OutStream& o, uint32_t d) {
  o.wpos = d & ~3u;
  o.nb = o.skip = d & 3u;
  o.lo = 0;
  o.rot = 0;
}

// store the completed word v at o.wpos if `en`; the stream's first word is stored without its skipped bytes
__device__ __forceinline__ void os_store_first(OutStream& o, uint32_t v, bool en) {
  const uint32_t sk = o.skip;
  if (en && sk == 0) sts32(o.wpos, v);
  if (en && sk == 1) sts8(o.wpos + 1, v >> 8);
  if (en && (sk == 1 || sk == 2)) sts8(o.wpos + 2, v >> 16);
  if (en && sk != 0) sts8(o.wpos + 3, v >> 24);
  o.skip = en ? 0u : sk;
}

// Append n (1..4*NW-4) bytes starting at shared address s (any alignment) to the stream (NW = 5: up to 16).
// One batch of independent loads, then register work, then stores: a single shared-memory round
// trip per chunk, which is what bounds a lane copying many short runs.  The NW aligned words read
// cover [s, s+4*NW-4) at any alignment (bytes past the run are read but never used); source
// misalignment a and the stream's pending bytes nb fold into one funnel shift.
template <int NW>
__device__ __forceinline__ void os_chunk(OutStream& o, uint32_t s, uint32_t n) {
  const uint32_t sp = s & ~3u, a = s & 3u;
  uint32_t w[NW + 1];
#pragma unroll
  for (int i = 0; i < NW; ++i) w[i] = lds32(sp + 4u * i);
  w[NW] = 0;
  // output word q holds stream bytes [4q, 4q+4); stream byte j >= nb is source byte (a - nb + j) of w[]
  int delta = (int)a - (int)o.nb;
  if (delta < 0) {  // shift the window by one word so that the funnel amount is non-negative
#pragma unroll
    for (int i = NW; i > 0; --i) w[i] = w[i - 1];
    w[0] = 0;
    delta += 4;
  }
  const uint32_t sh = 8u * (uint32_t)delta;
  uint32_t v[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
  const uint32_t keep = (1u << (8u * o.nb)) - 1u;  // nb <= 3
  v[0] = (v[0] & ~keep) | o.lo;
  const uint32_t tot = o.nb + n, nfull = tot >> 2, rem = tot & 3u;  // nfull <= NW - 1
  const uint32_t d = o.wpos;
  os_store_first(o, v[0], nfull > 0);
#pragma unroll
  for (int i = 1; i < NW - 1; ++i)
    if (nfull > (uint32_t)i) sts32(d + 4u * i, v[i]);
  uint32_t last = v[NW - 1];
#pragma unroll
  for (int i = NW - 2; i >= 0; --i) last = nfull == (uint32_t)i ? v[i] : last;
  o.wpos = d + 4u * nfull;
  o.lo = last & ((1u << (8u * rem)) - 1u);
  o.nb = rem;
}

// append [s, s+len) of shared memory (any alignment), len > 0.  Deliberately a real call (the
// stream state travels in registers) around a single copy of the chunk code: the kernel's hot code
// has to fit the SM's 32 KB instruction cache.
//   len <= 16 : one chunk
//   longer    : a first chunk of 13..16 bytes that word-aligns the source and puts the stream's
//               first (possibly shared) word out, then 16 bytes per step -- four independent loads,
//               one funnel shift and one whole-word store each, the pending-byte count constant --
//               and a last chunk of fewer than 16 bytes.
__device__ __noinline__ OutStream os_run_f(OutStream o, uint32_t s, uint32_t len) {
  uint32_t n = len > 16u ? 16u - (s & 3u) : len;
#pragma unroll 1
  while (true) {  // at most two turns; a loop so that the chunk code exists once
    os_chunk<5>(o, s, n);
    s += n;
    len -= n;
    if (len >= 16u && o.rot == 0) {
      const uint32_t sh = o.nb * 8u;
      uint32_t prev = sh ? (o.lo << (32u - sh)) : 0u;  // so that funnelshift_l(prev, w, sh) = lo | w << sh
      uint32_t dp = o.wpos;
      do {
        const uint32_t w0 = lds32(s), w1 = lds32(s + 4), w2 = lds32(s + 8), w3 = lds32(s + 12);
        sts32(dp, __funnelshift_l(prev, w0, sh));
        sts32(dp + 4, __funnelshift_l(w0, w1, sh));
        sts32(dp + 8, __funnelshift_l(w1, w2, sh));
        sts32(dp + 12, __funnelshift_l(w2, w3, sh));
        prev = w3;
        s += 16;
        dp += 16;
        len -= 16;
      } while (len >= 16u);
      o.lo = __funnelshift_l(prev, 0u, sh);
      o.wpos = dp;
    } else if (len >= 16u) {
      // The same steps, each reading the source word before it instead of carrying it over, so that
      // they can be taken in any order -- and every lane starts at a different one.  When all
      // records of a tile have the same length and that length is a multiple of 32 bytes, lanes
      // working at the same step hit the same shared-memory bank 8 to 32 ways (BASELINE config C2
      // on fixed-length reads); staggered starts spread them over the banks.
      const uint32_t nsteps = len >> 4, sh = o.nb * 8u, dp = o.wpos;
      uint32_t q = o.rot - 1u < nsteps ? o.rot - 1u : 0u;
#pragma unroll 1
      for (uint32_t it = 0; it < nsteps; ++it) {
        const uint32_t a = s + 16u * q, b = dp + 16u * q;
        const uint32_t wm = lds32(a - 4), w0 = lds32(a), w1 = lds32(a + 4), w2 = lds32(a + 8), w3 = lds32(a + 12);
        sts32(b, __funnelshift_l(wm, w0, sh));
        sts32(b + 4, __funnelshift_l(w0, w1, sh));
        sts32(b + 8, __funnelshift_l(w1, w2, sh));
        sts32(b + 12, __funnelshift_l(w2, w3, sh));
        q = q + 1u == nsteps ? 0u : q + 1u;
      }
      s += 16u * nsteps;
      len -= 16u * nsteps;
      o.wpos = dp + 16u * nsteps;
      o.lo = __funnelshift_l(lds32(s - 4), 0u, sh);
    }
    if (len == 0) break;
    n = len;
  }
  return o;
}
__device__ __forceinline__ void os_run(OutStream& o, uint32_t s, uint32_t len) { o = os_run_f(o, s, len); }
__device__ __forceinline__ void os_copy(OutStream& o, uint32_t s, uint32_t len) { os_run(o, s, len); }

__device__ __forceinline__ void os_end(const OutStream& o) {
#pragma unroll
  for (uint32_t b = 0; b < 3; ++b)
    if (b >= o.skip && b < o.nb) sts8(o.wpos + b, o.lo >> (8u * b));
}

// append the n (0..4) low bytes of register value v (bytes of v above n must be zero)
__device__ __noinline__ OutStream os_put_f(OutStream o, uint32_t v, uint32_t n) {
  const uint32_t sh = o.nb * 8u;
  const uint32_t word = o.lo | (v << sh);
  const uint32_t tot = o.nb + n;
  const bool full = tot >= 4u;
  os_store_first(o, word, full);
  o.wpos += full ? 4u : 0u;
  o.lo = full ? __funnelshift_l(v, 0u, sh) : word;  // v >> (32 - sh); 0 when sh == 0
  o.nb = tot & 3u;
  return o;
}
__device__ __forceinline__ void os_put(OutStream& o, uint32_t v, uint32_t n) { o = os_put_f(o, v, n); }

// ---- runs of unchanged records: warp-cooperative copy ------------------------------------------------
//
// Consecutive kept records that are copied verbatim (already in the writer's framing) are one
// contiguous run in the input tile and in the output.  All lanes of a warp copy it together:
// consecutive lanes own consecutive 16-byte aligned destination chunks, read the two aligned
// source chunks under them and funnel-shift by the run's relative misalignment, so every access of
// a pass covers 512 consecutive bytes -- no bank conflicts, a third of the shared-memory wavefronts
// that lane-per-record copying costs at a 329-byte record stride.  Only chunks completely inside
// the run are written; its ragged ends (< 16 bytes each) go through the stream writer.
//   x: first destination chunk (aligned up)   y: end of the last full chunk (aligned down)
//   z: source - destination                   w: passes of earlier runs of the tile
__device__ __forceinline__ uint4 bulk_desc(uint32_t s, uint32_t d, uint32_t n, uint32_t w) {
  return make_uint4((d + 15u) & ~15u, (d + n) & ~15u, s - d, w);
}
// descriptor warp-uniform; the warp does passes t0, t0 + tstep, ... of 32 chunks each
__device__ __forceinline__ void bulk16(const uint4 r, int lane, uint32_t t0, uint32_t tstep) {
  const uint32_t delta = r.z;
  const uint32_t off = delta & ~15u, sh = (delta & 3u) * 8u;
  const bool q1 = (delta & 4u) != 0, q2 = (delta & 8u) != 0;  // word part of the misalignment, warp-uniform
  for (uint32_t D = r.x + 16u * (32u * t0 + (uint32_t)lane); D < r.y; D += 512u * tstep) {
    const uint4 x = lds128(D + off), y = lds128(D + off + 16u);
    const uint32_t a0 = q2 ? x.z : x.x, a1 = q2 ? x.w : x.y, a2 = q2 ? y.x : x.z, a3 = q2 ? y.y : x.w;
    const uint32_t a4 = q2 ? y.z : y.x, a5 = q2 ? y.w : y.y;
    const uint32_t b0 = q1 ? a1 : a0, b1 = q1 ? a2 : a1, b2 = q1 ? a3 : a2, b3 = q1 ? a4 : a3, b4 = q1 ? a5 : a4;
    sts128(D, __funnelshift_r(b0, b1, sh), __funnelshift_r(b1, b2, sh), __funnelshift_r(b2, b3, sh),
           __funnelshift_r(b3, b4, sh));
  }
}

// ---- stand-alone run copy ------------------------------------------------------------------------------------------
//
// Every field of a rewritten record has a destination that is a closed form of the record's header length, read
// length and the program's constants (emit_part), so a field does not have to travel through the stream writer: it
// can be copied on its own.  cp_run copies [s, s + n) to [d, d + n), n >= 1, shared addresses at any alignment, and
// writes exactly those bytes (the words at both ends may belong to other writers as well): up to three bytes bring the
// destination to a word boundary, whole words follow -- 16 bytes per step, each word an aligned load and a funnel
// shift by the source's misalignment -- and up to three bytes end the run.  Against os_run_f this leaves out the two
// os_chunk passes (60-65 instructions each) at the ends of every run.  Reads up to 19 bytes past the run: inside the
// staged tile or the buffer behind it, never used.
// Used by every launch, and for the ragged ends of the unchanged records' runs as well, so that the stream writer
// (os_*) is not part of the default build's kernels at all: they are 4 KB smaller for it (the anchored single-end
// kernel 41.2 -> 36.9 KB).  Measured on one box, 10 M units, BK_DEFINES="BK_EMIT_CP=0" (the stream writer) against the
// default: C1 1.657 / 1.647 -> 1.624 / 1.646 ms, C3 2.857 / 2.844 -> 2.816 / 2.820 ms, C2 2.897 / 2.889 -> 2.897 / 2.893 ms,
// C4 4.839 / 4.842 -> 4.786 / 4.794 ms.  (BK_EMIT_CP=1, cp_run in the single-end launches only: the paired kernels then
// carry both writers, C4 4.814-4.886 ms.)
#ifndef BK_CP_ROT
#define BK_CP_ROT 0
#endif
#ifndef BK_EMIT_CP
#define BK_EMIT_CP 2  // 0: the stream writer everywhere, 1: cp_run in the single-end launches only, 2: everywhere
#endif
constexpr bool kEmitCp = BK_EMIT_CP != 0;
// `rot` (0, or 1 + a step): the 16-byte steps of a long run are taken from that one on, wrapping round -- lanes whose
// runs sit a multiple of 32 bytes apart then work on different banks (see os_run_f).
__device__ __noinline__ void cp_run(uint32_t d, uint32_t s, uint32_t n, uint32_t rot = 0u) {
  {
    const uint32_t h0 = (0u - d) & 3u;
    const uint32_t h = h0 < n ? h0 : n;
    const uint32_t sp = s & ~3u;
    const uint32_t v = __funnelshift_r(lds32(sp), lds32(sp + 4u), (s & 3u) * 8u);  // source bytes 0..3
    if (h > 0u) sts8(d, v);
    if (h > 1u) sts8(d + 1u, v >> 8);
    if (h > 2u) sts8(d + 2u, v >> 16);
    d += h;
    s += h;
    n -= h;
  }
  uint32_t sp = s & ~3u;
  const uint32_t sh = (s & 3u) * 8u;
  uint32_t prev = lds32(sp);
  uint32_t nw = n >> 2;
#if BK_CP_ROT
  // Bank-aware start steps.  A lane's loads at step t hit banks 4 * (group + t) + class + j (mod 32), group = bits 4..6
  // and class = bits 2..3 of its source address: lanes of different classes never meet, and the (up to eight) lanes of
  // one class that are dealt the start steps 0..7 relative to their groups never meet either -- the loads of a long run
  // become conflict-free whatever the records' lengths (lane-per-record copying otherwise runs 2-3 ways conflicted:
  // record starts are as good as random modulo 128).
  if (nw >= 16u) {
    const unsigned act = __activemask();
    const unsigned same = __match_any_sync(act, (sp >> 2) & 3u);
    const uint32_t rank = __popc(same & ((1u << (threadIdx.x & 31)) - 1u));
    rot = 1u + ((rank - (sp >> 4)) & 7u);
  }
#endif
  if (rot != 0u && nw >= 8u) {  // each step reads the source word before it instead of carrying it over: any order will do
    const uint32_t nsteps = nw >> 2;
    uint32_t q = rot - 1u < nsteps ? rot - 1u : 0u;
#pragma unroll 1
    for (uint32_t it = 0; it < nsteps; ++it) {
      const uint32_t x = sp + 16u * q, y = d + 16u * q;
      const uint32_t wm = lds32(x), a = lds32(x + 4u), b = lds32(x + 8u), c = lds32(x + 12u), e = lds32(x + 16u);
      sts32(y, __funnelshift_r(wm, a, sh));
      sts32(y + 4u, __funnelshift_r(a, b, sh));
      sts32(y + 8u, __funnelshift_r(b, c, sh));
      sts32(y + 12u, __funnelshift_r(c, e, sh));
      q = q + 1u == nsteps ? 0u : q + 1u;
    }
    sp += 16u * nsteps;
    d += 16u * nsteps;
    nw -= 4u * nsteps;
    prev = lds32(sp);
  }
#pragma unroll 1
  for (; nw >= 4u; nw -= 4u) {
    const uint32_t a = lds32(sp + 4u), b = lds32(sp + 8u), c = lds32(sp + 12u), e = lds32(sp + 16u);
    sts32(d, __funnelshift_r(prev, a, sh));
    sts32(d + 4u, __funnelshift_r(a, b, sh));
    sts32(d + 8u, __funnelshift_r(b, c, sh));
    sts32(d + 12u, __funnelshift_r(c, e, sh));
    prev = e;
    sp += 16u;
    d += 16u;
  }
  // the last 0..3 words and 0..3 bytes
  const uint32_t a = lds32(sp + 4u), b = lds32(sp + 8u), c = lds32(sp + 12u), e = lds32(sp + 16u);
  const uint32_t v0 = __funnelshift_r(prev, a, sh), v1 = __funnelshift_r(a, b, sh), v2 = __funnelshift_r(b, c, sh),
                 v3 = __funnelshift_r(c, e, sh);
  if (nw > 0u) sts32(d, v0);
  if (nw > 1u) sts32(d + 4u, v1);
  if (nw > 2u) sts32(d + 8u, v2);
  const uint32_t t = nw == 0u ? v0 : nw == 1u ? v1 : nw == 2u ? v2 : v3;
  d += 4u * nw;
  const uint32_t r = n & 3u;
  if (r > 0u) sts8(d, t);
  if (r > 1u) sts8(d + 1u, t >> 8);
  if (r > 2u) sts8(d + 2u, t >> 16);
}
// the n (1..4) low bytes of v at any address
__device__ __forceinline__ void put_bytes(uint32_t d, uint32_t v, uint32_t n) {
  sts8(d, v);
  if (n > 1u) sts8(d + 1u, v >> 8);
  if (n > 2u) sts8(d + 2u, v >> 16);
  if (n > 3u) sts8(d + 3u, v >> 24);
}

constexpr int kEmitParts = 4;  // back warps per mate; part j of every record is written by warp j
#ifndef BK_SE_BACKS
#define BK_SE_BACKS 4  // back warps of a single-end launch (bk_extract.cuh)
#endif
constexpr bool kSplitLines = BK_SE_BACKS >= 6;  // six or seven parts: the seq and the qual line are two parts each

// Part `part` of `nparts` (3, 4 or 5) of one record's output at shared address d.  The record is
//   '@' head { " NAME:" seq[cap] ":" qual[cap] }* "\n" seq[..s0) seq[e0..] "\n+\n" qual[..s0) qual[e0..] "\n"
// (parse.rs:101-113,138-148,161-168; fastq.rs:261) and every field's destination offset is a closed
// form of head_len, seq_len and the program's constants, so its five jobs
//   A: '@' head     B: first half of the captures     C: remaining captures + "\n"
//   D: trimmed seq + "\n+\n"                           E: trimmed qual + "\n"
// can be produced by different warps independently: one job each with five parts (the mate whose
// reads are rewritten, when the other mate has no pattern and lends a warp), A+B | C | D | E with
// four, A+C | D | E with three (a mate without a pattern: nothing is captured).  Parts meet at
// arbitrary byte positions; the words they share are written bytewise by both sides, never as
// whole words.  (A record that is copied unchanged and already has the writer's framing is not
// written here but as part of a run, see bulk16.)  Requires qual_len == seq_len for matched records
// (the front warp treats anything else as unmatched; bk_tokenize never produces it).
template <class PG>
__device__ __forceinline__ void emit_part(const PG& pg, const RecView& r, int ms, bool trim, uint32_t d,
                                          uint32_t part, uint32_t nparts, uint32_t rot, bool cp = false) {
  OutStream o;
  // jobs of this part, bit 0 = A ... bit 4 = E
  // (six or seven parts: the seq and the qual line are two parts each, `half` = 1 / 2)
  const uint32_t jobs = nparts == 7 ? (part < 3u ? 1u << part : part < 5u ? 8u : 16u)
                        : nparts == 6 ? (part == 0 ? 3u : part == 1 ? 4u : part < 4u ? 8u : 16u)
                        : nparts == 5 ? 1u << part : nparts == 4 ? (part == 0 ? 3u : 2u << part) : (part == 0 ? 5u : 4u << part);
  // (compiled in only when the build asks for that many back warps, BK_SE_BACKS in bk_extract.cuh: the default build's
  // hot code stays as it was)
  const uint32_t half = !kSplitLines ? 0u : nparts == 7 ? (part >= 3u ? ((part - 3u) & 1u) + 1u : 0u) : nparts == 6 ? (part >= 2u ? (part & 1u) + 1u : 0u) : 0u;
  const uint32_t ncaps = ms >= 0 ? pg.ncaps : 0u;
  const uint32_t csplit = (ncaps + 1u) >> 1;
  const uint32_t H = r.head_len, L = r.seq_len;
  if (jobs & 7u) {
    // header line: captures [c0, c1) of it belong to this part
    const uint32_t c0 = (jobs & 2u) ? 0u : csplit, c1 = (jobs & 4u) ? ncaps : (jobs & 2u) ? csplit : c0;
    uint32_t off = 1u + H;
    if (!(jobs & 3u)) {
#pragma unroll 1
      for (uint32_t c = 0; c < c0; ++c) off += 3u + pg.caps[c].name_len + 2u * pg.caps[c].len;
    }
    if (cp) {  // every field copied on its own (cp_run)
      uint32_t pos = d + ((jobs & 1u) ? 0u : off);
      if (jobs & 1u) {
        cp_run(pos, r.beg, 1u + H);  // '@' + head
        pos += 1u + H;
      }
#pragma unroll 1
      for (uint32_t c = c0; c < c1; ++c) {  // pattern order (parse.rs:101)
        const bk_cap cap = pg.caps[c];
        const uint32_t b3 = cap.name_len == 3 ? (uint32_t)(uint8_t)cap.name[2] : (uint32_t)':';
        put_bytes(pos, (uint32_t)' ' | ((uint32_t)(uint8_t)cap.name[0] << 8) | ((uint32_t)(uint8_t)cap.name[1] << 16) | (b3 << 24), 4u);
        if (cap.name_len == 3) sts8(pos + 4u, (uint32_t)':');  // " NAME:"  (parse.rs:166); names are 2 or 3 bytes
        pos += 2u + cap.name_len;
        if (cap.len) cp_run(pos, r.seq + ms + cap.off, cap.len);
        pos += cap.len;
        sts8(pos, (uint32_t)':');
        pos += 1u;
        if (cap.len) cp_run(pos, r.qual + ms + cap.off, cap.len);
        pos += cap.len;
      }
      if (jobs & 4u) sts8(pos, (uint32_t)'\n');
      return;
    }
    if (jobs & 1u) {
      os_begin(o, d);
      os_run(o, r.beg, 1u + H);  // '@' + head
    } else {
      os_begin(o, d + off);
    }
#pragma unroll 1
    for (uint32_t c = c0; c < c1; ++c) {  // pattern order (parse.rs:101)
      const bk_cap cap = pg.caps[c];
      const uint32_t b3 = cap.name_len == 3 ? (uint32_t)(uint8_t)cap.name[2] : (uint32_t)':';
      os_put(o, (uint32_t)' ' | ((uint32_t)(uint8_t)cap.name[0] << 8) | ((uint32_t)(uint8_t)cap.name[1] << 16) |
                    (b3 << 24), 4u);  // " NAME:"  (parse.rs:166); names are 2 or 3 bytes
      if (cap.name_len == 3) os_put(o, (uint32_t)':', 1u);
      if (cap.len) os_copy(o, r.seq + ms + cap.off, cap.len);
      os_put(o, (uint32_t)':', 1u);
      if (cap.len) os_copy(o, r.qual + ms + cap.off, cap.len);
    }
    if (jobs & 4u) os_put(o, (uint32_t)'\n', 1u);
    os_end(o);
    return;
  }
  // D and E are the same job on the seq and on the qual line (parse.rs:138-148: cut [s0, e0) out of both)
  const uint32_t G = ms >= 0 ? pg.hdr_growth : 0u;
  const bool cut = ms >= 0 && trim;
  const uint32_t cs = cut ? (uint32_t)ms : 0u, ce = cut ? (uint32_t)ms + pg.total_len : 0u;
  const bool isD = (jobs & 8u) != 0;
  const uint32_t ll = isD ? L : r.qual_len;  // == L for matched records; an unmatched one is not cut
  const uint32_t src = isD ? r.seq : r.qual;
  const uint32_t ps = cs < ll ? cs : ll, pe = ce < ll ? ce : ll;
  const uint32_t sep = isD ? (uint32_t)'\n' | ((uint32_t)'+' << 8) | ((uint32_t)'\n' << 16) : (uint32_t)'\n';
  const uint32_t nsep = isD ? 3u : 1u;
  const uint32_t dpos = d + 2u + H + G + (isD ? 0u : L - (ce - cs) + 3u);
  // the line end ("\n+\n" after seq, "\n" after qual) travels with the run behind the cut when the framing has it there
  const uint32_t tl = ll - pe + (r.fast ? nsep : 0u);
  if (half) {
    // Two warps per line: the longer of the two runs (in front of the cut / behind it) is split near its middle, at a
    // multiple of 16 bytes from its start; the halves meet at an arbitrary byte of the output, like any two parts.
    const bool in_tail = tl >= ps;
    const uint32_t n1 = ((in_tail ? tl : ps) >> 1) & ~15u;
    if (half == 1u) {
      os_begin(o, dpos);
      o.rot = rot;
      if (in_tail) {
        if (ps) os_run(o, src, ps);
        if (n1) os_run(o, src + pe, n1);
      } else if (n1) {
        os_run(o, src, n1);
      }
    } else {
      os_begin(o, dpos + (in_tail ? ps + n1 : n1));
      o.rot = rot;
      if (in_tail) {
        if (tl > n1) os_run(o, src + pe + n1, tl - n1);
      } else {
        os_run(o, src + n1, ps - n1);
        if (tl) os_run(o, src + pe, tl);
      }
      if (!r.fast) os_put(o, sep, nsep);
    }
    os_end(o);
    return;
  }
  if (cp) {
    if (ps) cp_run(dpos, src, ps, rot);
    if (tl) cp_run(dpos + ps, src + pe, tl, rot);
    if (!r.fast) put_bytes(dpos + ps + tl, sep, nsep);
    return;
  }
  os_begin(o, dpos);
  o.rot = rot;
  if (ps) os_run(o, src, ps);
  if (tl) os_run(o, src + pe, tl);
  if (!r.fast) os_put(o, sep, nsep);
  os_end(o);
}

// A record of the never-staged mate, parsed with probes instead of a scan.  `H` is a guess of its header length
// (the other mate's, which is the same in every FASTQ pair the common tools write).  If the record is in the
// writer's framing with that header length, its bytes at four positions are line ends:
//   '@' H bytes LF | L bytes LF | '+' LF | L bytes LF          with L = (e - b - H - 6) / 2.
// A record holds exactly four line ends (its bounds come from bk_tokenize: every fourth newline), so four
// distinct ones found are all of them and the lines are as assumed.  Anything else (another header length, CR
// line ends, text after '+', no final EOL) returns false and takes the scanning parser.  One round trip.
__device__ __forceinline__ bool parse_probe_g(const uint8_t* t, uint32_t b, uint32_t e, uint32_t H, RecView& r) {
  if (H > e - b) return false;  // (the guess may be anything)
  const uint32_t sb = b + H + 2u;
  if (e < sb + 4u || ((e - sb) & 1u)) return false;
  const uint32_t L = (e - sb - 4u) >> 1;
  const uint32_t ch = __ldg(t + sb - 1u), c0 = __ldg(t + sb + L), c1 = __ldg(t + sb + L + 1u), c2 = __ldg(t + sb + L + 2u);
  const uint32_t c3 = __ldg(t + e - 1u), c4 = __ldg(t + sb + L - 1u), c5 = __ldg(t + e - 2u), c6 = __ldg(t + sb - 2u);
  const bool framed = (ch == '\n') & (c0 == '\n') & (c1 == '+') & (c2 == '\n') & (c3 == '\n');
  const bool cr = ((L != 0) & ((c4 == '\r') | (c5 == '\r'))) | ((H != 0) & (c6 == '\r'));
  if (!framed || cr) return false;
  r.beg = b;
  r.end = e;
  r.fast = 1;
  r.head_len = H;
  r.seq = sb;
  r.seq_len = L;
  r.qual = sb + L + 3u;
  r.qual_len = L;
  return true;
}

// a kept record of such a mate in foreign framing (CR line ends, text after '+', no final EOL): re-framed
// bytewise by its lane (rare)
__device__ __noinline__ void emit_record_g(const uint8_t* t, uint32_t beg, uint32_t head_len, uint32_t seq, uint32_t seq_len,
                                           uint32_t qual, uint32_t qual_len, uint8_t* o) {
  *o++ = '@';
  for (uint32_t i = 0; i < head_len; ++i) *o++ = (uint8_t)__ldg(t + beg + 1u + i);
  *o++ = '\n';
  for (uint32_t i = 0; i < seq_len; ++i) *o++ = (uint8_t)__ldg(t + seq + i);
  *o++ = '\n'; *o++ = '+'; *o++ = '\n';
  for (uint32_t i = 0; i < qual_len; ++i) *o++ = (uint8_t)__ldg(t + qual + i);
  *o++ = '\n';
}

// ---- unanchored search with the windows dealt out again every round (match_kernel) ----------------------------------
//
// match_unanchored gives every lane the windows of ITS read, one per round, and the warp runs until its slowest read is
// done: on BASELINE config C4 nine reads in ten match in their first window, but nearly every tile holds a read
// without a match, so all tiles run all four rounds with three lanes in thirty-two still working.  Here the warp deals
// the open windows out again each round: with P reads undecided, every one of them gets g = min(32 / P, windows
// left) consecutive windows, one lane each (the lane fetches that read's bounds by shuffle), and the read's owner takes
// the lowest hit of its g lanes -- the leftmost match is the one in the lowest window.  The first round is the old
// one (P = 32, g = 1); the second usually finishes the tile.
__device__ __noinline__ int match_unanchored_rd(uint32_t tab, uint32_t seq, uint32_t L, bool can, int lane) {
  const uint32_t T = lds32(tab + kTabHdr + 4u);
  const MOp op0 = tab_op(tab, 0);
  if (!bs_eligible(op0.kind, op0.len, op0.k)) return match_unanchored(tab, seq, L, can, lane, 0, 1, nullptr, 0u);
  bool pend = can && L >= T;
  const uint32_t last = pend ? L - T : 0u;  // this read's last candidate start
  const uint32_t nwin = last / 32u + 1u;
  uint32_t wnext = 0;  // this read's first window not searched yet
  int ms = -1;
  while (true) {
    const unsigned pm = __ballot_sync(0xffffffffu, pend);
    if (!pm) break;
    const uint32_t P = (uint32_t)__popc(pm);
    const uint32_t maxrem = __reduce_max_sync(0xffffffffu, pend ? nwin - wnext : 0u);
    const uint32_t g = min(32u / P, maxrem);  // >= 1
    // this lane's share: window wi of the j-th undecided read
    const uint32_t j = (uint32_t)lane / g, wi = (uint32_t)lane - j * g;
    const bool work = j < P;
    // the lane of the j-th undecided read = the largest t with fewer than j + 1 set bits of pm below it (binary search;
    // the __fns intrinsic is a 32-step loop and was 12 % of this kernel's instructions)
    uint32_t r = 0;
#pragma unroll
    for (uint32_t step = 16u; step; step >>= 1)
      if ((uint32_t)__popc(pm & ((1u << (r + step)) - 1u)) <= j) r += step;
    if (!work) r = 0;
    const uint32_t rseq = __shfl_sync(0xffffffffu, seq, (int)r), rlast = __shfl_sync(0xffffffffu, last, (int)r);
    const uint32_t q = 32u * (__shfl_sync(0xffffffffu, wnext, (int)r) + wi);
    uint32_t cand = 0;
    if (work && q <= rlast) {
      const uint32_t cb = rseq + q + op0.off;
      uint32_t a0, a1, b0, b1;
      bs_planes<8>(cb, a0, a1);
      bs_planes<2>(cb + 32u, b0, b1);  // (the literal reaches at most 7 bytes into the next 32)
      uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 1
      for (uint32_t i = 0; i < op0.len; ++i) {
        const uint2 x = lds64(tab + kTabSA + 8u * i);
        const uint32_t mm = (__funnelshift_r(a0, b0, i) ^ x.x) | (__funnelshift_r(a1, b1, i) ^ x.y);
        const uint32_t carry = c0 & mm;
        c0 ^= mm;
        c2 |= c1 & carry;
        c1 ^= carry;
      }
      cand = ~c2 & (op0.k == 0u ? ~(c0 | c1) : op0.k == 1u ? ~c1 : op0.k == 2u ? ~(c1 & c0) : 0xFFFFFFFFu);
      if (rlast - q < 31u) cand &= (2u << (rlast - q)) - 1u;
    }
    uint32_t mine = 0xFFFFFFFFu;
    while (__any_sync(0xffffffffu, cand != 0u)) {  // every lane's lowest candidate, verified together
      const bool have = cand != 0u;
      const uint32_t st = q + (have ? (uint32_t)__ffs(cand) - 1u : 0u);
      if (have && match_at_fwd_shared(tab, rseq, st)) {
        mine = st;
        cand = 0;
      } else if (have) {
        cand &= cand - 1u;
      }
    }
    // the owner of an undecided read collects its g lanes
    const uint32_t myj = (uint32_t)__popc(pm & ((1u << lane) - 1u));
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < g; ++i) best = min(best, __shfl_sync(0xffffffffu, mine, (int)((myj * g + i) & 31u)));
    if (pend) {
      if (best != 0xFFFFFFFFu) {
        ms = (int)best;
        pend = false;
      } else {
        wnext += g;
        pend = wnext < nwin;
      }
    }
  }
  return ms;
}

// ---- unanchored search over a list of variants (match_kernel) ---------------------------------------------------
//
// Bounded repetition in front of or behind a literal -- `(?P<CB>[ATGCN]{14,16})atgccat` -- gives variants that all carry
// the SAME first literal (bytes, length, mismatch budget), only at different offsets.  Every lane scans its read as
// match_unanchored does: the bit planes of 64 text bytes from the window's first start + the smallest literal
// offset are made once, and per variant only the counter loop runs, on the planes shifted by that variant's offset;
// that leaves one word of candidate starts per variant.  The candidates are then verified in the regex crate's
// order -- ascending start, and at one start the variants in list order -- against the variant's whole program.
// Requires (checked by the caller): no anchors, every variant's ops[0] an eligible literal equal to variant 0's,
// offsets within 24 bytes of each other.  `voff` = the variants' literal offsets (shared memory, one byte each).
__device__ __noinline__ int match_variants_scan(const bk_devprog* vp, uint32_t nvar, uint32_t tab, uint32_t voff, uint32_t off_min,
                                                 uint32_t seq, uint32_t L, bool can) {
  const MOp op0 = tab_op(tab, 0);
  uint32_t tmin = 0xFFFFFFFFu;  // the shortest variant bounds the candidate starts
  for (uint32_t v = 0; v < nvar; ++v) tmin = min(tmin, vp[v].total_len);
  bool pend = can && L >= tmin;
  const uint32_t last = pend ? L - tmin : 0u;
  const uint32_t smax = __reduce_max_sync(0xffffffffu, last);
  if (!__any_sync(0xffffffffu, pend)) return -1;
  int ms = -1;
  for (uint32_t q = 0; q <= smax; q += 32u) {
    uint32_t cand[BK_MAX_VARIANTS];
    uint32_t any = 0;
    if (pend && q <= last) {
      const uint32_t cb = seq + q + off_min;
      uint32_t a0, a1, b0, b1;
      bs_planes<8>(cb, a0, a1);
      bs_planes<8>(cb + 32u, b0, b1);
      for (uint32_t v = 0; v < nvar; ++v) {
        const uint32_t dlt = lds8(voff + v) - off_min;  // <= 24
        uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 1
        for (uint32_t i = 0; i < op0.len; ++i) {
          const uint2 x = lds64(tab + kTabSA + 8u * i);
          const uint32_t mm = (__funnelshift_r(a0, b0, i + dlt) ^ x.x) | (__funnelshift_r(a1, b1, i + dlt) ^ x.y);
          const uint32_t carry = c0 & mm;
          c0 ^= mm;
          c2 |= c1 & carry;
          c1 ^= carry;
        }
        uint32_t cv = ~c2 & (op0.k == 0u ? ~(c0 | c1) : op0.k == 1u ? ~c1 : op0.k == 2u ? ~(c1 & c0) : 0xFFFFFFFFu);
        const uint32_t T = vp[v].total_len;  // this variant's own last start
        if (L < T || q > L - T) cv = 0;
        else if (L - T - q < 31u) cv &= (2u << (L - T - q)) - 1u;
        cand[v] = cv;
        any |= cv;
      }
    }
    while (any) {  // (per lane: candidates are a few per read)
      const uint32_t b = (uint32_t)__ffs(any) - 1u;
      for (uint32_t v = 0; v < nvar && ms < 0; ++v)
        if (((cand[v] >> b) & 1u) && match_at_variant<false>(vp[v], tab, seq, L, q + b)) ms = (int)((q + b) | (v << 24));
      if (ms >= 0) break;
      any &= any - 1u;
    }
    if (ms >= 0) pend = false;
    if (!__any_sync(0xffffffffu, pend)) break;
  }
  return ms;
}

// parse::get_reverse_complement as a stand-alone op (API parity; the extract kernel never materialises it)
__global__ void rc_kernel(const uint8_t* in, uint8_t* out, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = complement(in[n - 1 - i]);
}

// ---- the general form: a backtracking program (bk_program.h: bk_vmop; match_kernel<VM = true>) ------------------------
//
// One lane per read, the regex crate's leftmost-first semantics by construction (pattern.rs:225-230): starts in ascending
// order, at every SPLIT the preferred branch first, the other one on a stack in local memory.  An entry is 8 bytes:
// kind << 30 | rem << 12 | pc and a text position -- BRANCH (continue at pc, pos), RESTORE (capture slot pc gets its old
// value back), STAR_G / STAR_L (a repeated one-byte class: ONE entry for all its lengths; greedy gives a byte back per
// visit, lazy takes one more).
// Plain backtracking is polynomial in the read length (one degree per repetition) where the reference's engine is
// linear, so the search is memoised the way a bounded backtracker is: the instructions at which paths can meet again --
// SPLIT, STAR and what follows a STAR -- carry a mark (bk_vmop::c), and a (mark, position) state is entered once per READ:
// a second arrival, from whichever start, can only fail again (no back-references, and no loop body matches the empty
// string).  That bounds the work by marks x (length + 1) states.  The visited bits (kVmVisWords words of local memory)
// must hold them, and kVmStack entries the stack; a read that needs more raises error bit 8 and the batch is refused
// (BK_E_UNSUPPORTED) -- giving a different answer instead is not an option.
constexpr uint32_t kVmStack = 384, kVmVisWords = 256, kVmSteps = 32;
constexpr uint32_t kVmBranch = 0u, kVmRestore = 1u, kVmStarG = 2u, kVmStarL = 3u;

// -> the match's start; -1: no match; -2: a match in which a named group took no part (the read is dropped)
template <bool RC>
__device__ __noinline__ int match_vm(uint32_t tab, uint32_t vmc, uint32_t vm_len, uint32_t ncaps, uint32_t seq, uint32_t L,
                                     uint4* spans, uint32_t* gave_up) {
  const uint4 h = lds128(tab + kTabHdr);  // marked instructions, shortest match, anchored at the start, -
  if (L < h.y) return -1;
  const uint32_t nstates = h.x * (L + 1u);
  if (L > 0xFFFFu || nstates > 32u * kVmVisWords) { *gave_up = 1u; return -1; }
  const uint32_t last = h.z ? 0u : L - h.y;
  // (a greedy run gives its bytes back one visit at a time even when every state behind them is known: up to a
  // read's length of one-step visits per STAR state)
  uint32_t budget = kVmSteps * vm_len * (L + 1u) + nstates * (L + 1u);
  uint32_t run_pc = 0xFFFFFFFFu, run_from = 0, run_to = 0;  // the last greedy STAR's run of class bytes: [run_from, run_to)
  uint32_t cap[2 * BK_MAX_CAPS];
  uint32_t vis[kVmVisWords];
  uint2 stk[kVmStack];
  for (uint32_t i = 0; i <= nstates >> 5; ++i)
    if (i < kVmVisWords) vis[i] = 0u;
  for (uint32_t s = 0; s <= last; ++s) {
    uint32_t sp = 0, pc = 0, pos = s;
#pragma unroll
    for (uint32_t c = 0; c < 2 * BK_MAX_CAPS; ++c) cap[c] = 0xFFFFFFFFu;
    for (;;) {
      if (budget-- == 0u) { *gave_up = 1u; return -1; }
      const uint2 w = lds64(vmc + 8u * pc);  // bk_vmop, little endian
      const uint32_t op = w.x & 0xFFu, k = (w.x >> 8) & 0xFFu, a = w.x >> 16, b = w.y & 0xFFFFu, mark = w.y >> 16;
      bool ok = true;
      uint2 push = make_uint2(0xFFFFFFFFu, 0u);  // (x = all ones: nothing to push)
      if (mark) {  // been in this state before: it failed then, it fails now
        const uint32_t bit = (mark - 1u) * (L + 1u) + pos, m = 1u << (bit & 31u), v = vis[bit >> 5];
        ok = !(v & m);
        vis[bit >> 5] = v | m;
      }
      if (!ok) {
      } else if (op == BK_VM_CLS) {
        ok = pos + b <= L;
        if (ok) {
          uint32_t acc = 0xFFu;
          for (uint32_t i = 0; i < b; ++i) acc &= lds8(tab + fetch<RC>(tab, seq, L, pos + i));
          ok = (acc >> a) & 1u;
          pos += b;
        }
        ++pc;
      } else if (op == BK_VM_LIT) {
        ok = pos + b <= L;
        if (ok) {
          uint32_t mism = 0, bad = 0;
          for (uint32_t i = 0; i < b; ++i) {
            const uint32_t c = fetch<RC>(tab, seq, L, pos + i);
            const uint32_t ne = (c != lds8(tab + kTabLit + a + i));
            mism += ne;
            bad |= ne & (uint32_t)((c >= 0x80u) | (c == '\n'));  // a wildcard slot is the regex `.` (pattern.rs:10,71)
          }
          ok = mism <= k && !bad;
          pos += b;
        }
        ++pc;
      } else if (op == BK_VM_SPLIT) {
        push = make_uint2(kVmBranch << 30 | b, pos);
        pc = a;
      } else if (op == BK_VM_JMP) {
        pc = a;
      } else if (op == BK_VM_SAVE) {
        push = make_uint2(kVmRestore << 30 | a, cap[a]);
        cap[a] = pos;
        ++pc;
      } else if (op == BK_VM_BOL) {
        ok = pos == 0u;
        ++pc;
      } else if (op == BK_VM_EOL) {
        ok = pos == L;
        ++pc;
      } else if (op == BK_VM_STAR) {
        ++pc;
        if (k == 0u) {  // greedy: the whole run, then a byte less per visit
          // (an unanchored search arrives here again one byte further on: the run it finds ends where the last did)
          uint32_t n = run_pc == pc && pos >= run_from && pos <= run_to ? run_to - pos : 0u;
          if (run_pc != pc || pos < run_from || pos > run_to) {
            while (pos + n < L && ((lds8(tab + fetch<RC>(tab, seq, L, pos + n)) >> a) & 1u)) ++n;
            run_pc = pc; run_from = pos; run_to = pos + n;
          }
          n = min(n, b);
          if (n) push = make_uint2(kVmStarG << 30 | (n - 1u) << 12 | pc, pos + n - 1u);
          pos += n;
        } else if (b) {  // lazy: nothing, then a byte more per visit
          push = make_uint2(kVmStarL << 30 | b << 12 | (pc - 1u), pos);
        }
      } else {  // BK_VM_MATCH
        uint32_t off[3] = {0, 0, 0}, clen[3] = {0, 0, 0};
        bool all = true;
#pragma unroll
        for (uint32_t c = 0; c < BK_MAX_CAPS; ++c)
          if (c < ncaps) {
            // a named group that took no part in the match: get_barcode_match_positions fails and the read is dropped
            // (parse.rs:127-136, 82 `.ok()?`)
            if (cap[2 * c] == 0xFFFFFFFFu || cap[2 * c + 1] == 0xFFFFFFFFu) all = false;
            off[c] = cap[2 * c] - s;
            clen[c] = cap[2 * c + 1] - cap[2 * c];
          }
        if (!all) return -2;  // (a match all the same: no reverse-complement retry, parse.rs:61)
        *spans = pack_spans(pos - s, off, clen);
        return (int)s;
      }
      if (push.x != 0xFFFFFFFFu) {
        if (sp == kVmStack) { *gave_up = 1u; return -1; }
        stk[sp++] = push;
      }
      if (ok) continue;
      // ---- backtrack: the most recent alternative
      bool resumed = false;
      while (sp) {
        const uint2 e = stk[--sp];
        const uint32_t kind = e.x >> 30, epc = e.x & 0xFFFu, rem = (e.x >> 12) & 0xFFFFu;
        if (kind == kVmRestore) { cap[epc] = e.y; continue; }
        if (kind == kVmBranch) { pc = epc; pos = e.y; resumed = true; break; }
        if (kind == kVmStarG) {
          // the lengths left are e.y, e.y - 1 .. e.y - rem; those whose continuation state has been entered before can
          // only fail again: skip to the longest that has not, a word of visited bits at a time
          const uint32_t m0 = (lds32(vmc + 8u * epc + 4u) >> 16) - 1u;  // (what follows a STAR is marked)
          const uint32_t hi = m0 * (L + 1u) + e.y, lo = hi - rem;
          uint32_t bit = hi, found = 0xFFFFFFFFu;
          for (;;) {
            const uint32_t free_bits = ~vis[bit >> 5] & (0xFFFFFFFFu >> (31u - (bit & 31u)));
            if (free_bits) { found = (bit & ~31u) + 31u - (uint32_t)__clz(free_bits); break; }
            if ((bit & ~31u) <= lo) break;
            bit = (bit & ~31u) - 1u;
          }
          if (found == 0xFFFFFFFFu || found < lo) continue;  // nothing left of this run
          pc = epc; pos = e.y - (hi - found);
          if (found > lo) stk[sp++] = make_uint2(kVmStarG << 30 | (found - lo - 1u) << 12 | epc, pos - 1u);  // (the slot just freed)
          resumed = true;
          break;
        }
        // kVmStarL: one more byte of the class, if there is one
        const uint32_t cls = lds32(vmc + 8u * epc) >> 16;
        if (e.y < L && ((lds8(tab + fetch<RC>(tab, seq, L, e.y)) >> cls) & 1u)) {
          pc = epc + 1u; pos = e.y + 1u;
          if (rem > 1u) stk[sp++] = make_uint2(kVmStarL << 30 | (rem - 1u) << 12 | epc, e.y + 1u);
          resumed = true;
          break;
        }
      }
      if (!resumed) break;  // no match from this start
    }
  }
  return -1;
}

// ---- bksyn-v1 generator (bench / tests; oracle/bksyn.py is the normative statement) -------------------------

struct GenParams {
  uint64_t seed;
  uint64_t first_idx;
  uint32_t n;
  uint32_t read_len;
  uint32_t mate;
  uint32_t plant_len;
  uint32_t plant_off;
  uint32_t plant_jitter;
  char plant_lit[32];
};

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ int acgt_index(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }

// one thread per record; records are fixed-size so offsets are arithmetic
__global__ void gen_kernel(const __grid_constant__ GenParams G, uint8_t* text, uint32_t* rec_off) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t L = G.read_len;
  const uint32_t rl = 23 + 2 * L + 6;
  if (r == 0) rec_off[G.n] = G.n * rl;
  if (r >= G.n) return;
  rec_off[r] = r * rl;
  uint8_t* p = text + (size_t)r * rl;
  const uint64_t idx = G.first_idx + r;
  const uint64_t ctr = 8 * idx + 4 * (G.mate - 1);
  const uint64_t k0 = splitmix64(G.seed ^ splitmix64(ctr + 0));
  const uint64_t k1 = splitmix64(G.seed ^ splitmix64(ctr + 1));
  const uint64_t k2 = splitmix64(G.seed ^ splitmix64(ctr + 2));
  const uint64_t k3 = splitmix64(G.seed ^ splitmix64(ctr + 3));
  // head "bk.%012d %d:N:0:1"
  *p++ = '@'; *p++ = 'b'; *p++ = 'k'; *p++ = '.';
  {
    uint64_t v = idx;
    for (int d = 11; d >= 0; --d) { p[d] = (uint8_t)('0' + v % 10); v /= 10; }
    p += 12;
  }
  *p++ = ' '; *p++ = (uint8_t)('0' + G.mate);
  *p++ = ':'; *p++ = 'N'; *p++ = ':'; *p++ = '0'; *p++ = ':'; *p++ = '1';
  *p++ = '\n';
  uint8_t* bases = p;
  uint64_t wb = 0, wn = 0, wq = 0;
  for (uint32_t i = 0; i < L; ++i) {
    if ((i & 31) == 0) wb = splitmix64(k0 + (i >> 5));
    if ((i & 7) == 0) wn = splitmix64(k2 + (i >> 3));
    uint8_t b = "ACGT"[(wb >> (2 * (i & 31))) & 3];
    if (((wn >> (8 * (i & 7))) & 0xFF) == 0) b = 'N';
    bases[i] = b;
  }
  if (G.plant_len) {
    const uint32_t u = (uint32_t)(splitmix64(k3 + 0) & 0xFF);
    if (u < 250) {
      const uint32_t nsub = u < 205 ? 0 : u < 230 ? 1 : u < 243 ? 2 : 3;
      const uint32_t pos0 = G.plant_off + (uint32_t)(splitmix64(k3 + 5) % (G.plant_jitter + 1));
      const uint32_t ll = G.plant_len;
      if (pos0 + ll <= L) {
        for (uint32_t i = 0; i < ll; ++i) bases[pos0 + i] = (uint8_t)G.plant_lit[i];
        uint32_t used[3];
        for (uint32_t t = 0; t < nsub; ++t) {
          const uint64_t w = splitmix64(k3 + 1 + t);
          uint32_t q = (uint32_t)(w % ll);
          for (bool again = true; again;) {
            again = false;
            for (uint32_t v = 0; v < t; ++v)
              if (used[v] == q) { q = (q + 1) % ll; again = true; }
          }
          used[t] = q;
          const int orig = acgt_index(G.plant_lit[q]);
          bases[pos0 + q] = "ACGT"[(orig + 1 + (int)((w >> 8) % 3)) % 4];
        }
      }
    }
  }
  p += L;
  *p++ = '\n'; *p++ = '+'; *p++ = '\n';
  for (uint32_t i = 0; i < L; ++i) {
    if ((i & 7) == 0) wq = splitmix64(k1 + (i >> 3));
    p[i] = (uint8_t)(33 + ((wq >> (8 * (i & 7))) & 0xFF) % 41);
  }
  p[L] = '\n';
}

}  // namespace bk
