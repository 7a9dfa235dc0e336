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
  unsigned long long* desc[2];  // look-back descriptors, one per tile and output stream
  unsigned int* tile_counter;
  DevResult* result;
  unsigned long long* prof;  // 24 phase-cycle sums (only written when built with BK_PROFILE_PHASES)
  uint32_t n;
  uint32_t ntiles;
  uint32_t tile_recs;  // <= 32 (smaller when records are too long for the staging buffer)
  uint32_t in_cap[2];  // shared-memory bytes for the staged input / assembled output per mate
  uint32_t out_cap[2];
  uint32_t flags;
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D TMA: global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
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

__device__ __forceinline__ uint32_t strip_cr(uint32_t b, uint32_t e) {
  return (e > b && lds8(e - 1) == '\r') ? e - 1 : e;
}

__device__ __forceinline__ RecView parse_record(uint32_t b, uint32_t e) {
  RecView r;
  const uint32_t nl1 = find_nl(b, e);
  r.beg = b;
  r.end = e;
  r.fast = 0;
  r.head_len = strip_cr(b + 1, nl1 > b ? nl1 : b + 1) - (b + 1);
  const uint32_t sb = nl1 + 1 < e ? nl1 + 1 : e;
  // Fast path for the common layout `seq LF + LF qual LF`: then e - sb = 2L + 4.  A record of any
  // other shape (CR line ends, text after '+', missing final LF) cannot pass the three probes
  // with the L this formula gives (newlines only exist at the three line ends), so it falls
  // through to the scanning path.
  {
    const uint32_t rem = e - sb;
    if (rem >= 4 && !(rem & 1u)) {
      const uint32_t L = (rem - 4) >> 1;
      if (lds8(sb + L) == '\n' && lds8(sb + L + 1) == '+' && lds8(sb + L + 2) == '\n' && lds8(e - 1) == '\n' &&
          (L == 0 || (lds8(sb + L - 1) != '\r' && lds8(e - 2) != '\r'))) {
        r.seq = sb;
        r.seq_len = L;
        r.qual = sb + L + 3;
        r.qual_len = L;
        r.fast = (b + 1 + r.head_len == nl1);  // and no CR on the header line
        return r;
      }
    }
  }
  const uint32_t nl2 = find_nl(sb, e);
  r.seq = sb;
  r.seq_len = strip_cr(sb, nl2) - sb;
  const uint32_t pb = nl2 + 1 < e ? nl2 + 1 : e;
  const uint32_t nl3 = find_nl(pb, e);
  const uint32_t qb = nl3 + 1 < e ? nl3 + 1 : e;
  const uint32_t nl4 = find_nl(qb, e);
  r.qual = qb;
  r.qual_len = strip_cr(qb, nl4) - qb;
  return r;
}

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

template <bool RC>
__device__ __forceinline__ uint32_t fetch(uint32_t seq, uint32_t L, uint32_t j) {
  if (RC) return complement(lds8(seq + L - 1 - j));
  return lds8(seq + j);
}

// All compare runs of the program at start s.  Branch-free inside a run so that the lanes of a warp
// (one read each) stay in lock step: a data-dependent early exit would split the warp and replay
// the remaining runs once per split.  `tab[c]` = bit k set iff class k contains byte c (0 for
// c >= 0x80).  EARLY_OUT (unanchored search, lanes already diverge on s) stops at the first
// failed run.
template <bool RC, bool EARLY_OUT>
__device__ __forceinline__ bool match_at(const bk_devprog& pg, uint32_t tab, uint32_t seq, uint32_t L, uint32_t s) {
  bool ok = true;
  for (uint32_t o = 0; o < pg.nops; ++o) {
    const bk_op op = pg.ops[o];
    const uint32_t base = s + op.off;
    if (op.kind == BK_OP_LIT) {
      uint32_t mism = 0, bad = 0;
#pragma unroll 4
      for (uint32_t i = 0; i < op.len; ++i) {
        const uint32_t c = fetch<RC>(seq, L, base + i);
        const uint32_t ne = (c != pg.lit[op.idx + i]);
        mism += ne;
        // a wildcard slot is the regex `.`: any ASCII byte except '\n' (pattern.rs:10,71)
        bad |= ne & (uint32_t)((c >= 0x80u) | (c == '\n'));
      }
      ok = ok && (mism <= op.k) && !bad;
    } else {
      uint32_t acc = 0xFFu;
      uint32_t i = 0;
      for (; i + 4 <= op.len; i += 4) {  // 4 independent byte loads, then 4 independent table loads
        const uint32_t c0 = fetch<RC>(seq, L, base + i), c1 = fetch<RC>(seq, L, base + i + 1);
        const uint32_t c2 = fetch<RC>(seq, L, base + i + 2), c3 = fetch<RC>(seq, L, base + i + 3);
        acc &= lds8(tab + c0) & lds8(tab + c1) & lds8(tab + c2) & lds8(tab + c3);
      }
      for (; i < op.len; ++i) acc &= lds8(tab + fetch<RC>(seq, L, base + i));
      ok = ok && ((acc >> op.idx) & 1u);
    }
    if (EARLY_OUT && !ok) return false;
  }
  return ok;
}

// leftmost match start or -1
template <bool RC>
__device__ __forceinline__ int match_read(const bk_devprog& pg, uint32_t tab, uint32_t seq, uint32_t L) {
  const uint32_t T = pg.total_len;
  if (L < T) return -1;
  if (pg.anchor_start || pg.anchor_end) {  // exactly one candidate start
    if (pg.anchor_start && pg.anchor_end && L != T) return -1;
    const uint32_t s = pg.anchor_end ? L - T : 0;
    return match_at<RC, false>(pg, tab, seq, L, s) ? (int)s : -1;
  }
  for (uint32_t s = 0; s <= L - T; ++s)
    if (match_at<RC, true>(pg, tab, seq, L, s)) return (int)s;
  return -1;
}

// ---- emit ----------------------------------------------------------------------------------------------

// output bytes of a kept record
__device__ __forceinline__ uint32_t out_bytes(const bk_devprog& pg, const RecView& r, int ms, bool trim) {
  uint32_t sl = r.seq_len, ql = r.qual_len, g = 0;
  if (ms >= 0) {
    g = pg.hdr_growth;
    if (trim) {
      const uint32_t T = pg.total_len;
      sl -= T;  // match lies inside seq
      const uint32_t qs = (uint32_t)ms < ql ? (uint32_t)ms : ql;
      const uint32_t qe = (uint32_t)ms + T < ql ? (uint32_t)ms + T : ql;
      ql -= (qe - qs);
    }
  }
  return 1 + r.head_len + g + 1 + sl + 3 + ql + 1;
}

// Per-lane output stream into the shared output tile.  A record's bytes are produced strictly in
// order, so the lane keeps the (up to 3) bytes that do not yet fill a word in a register and only
// ever stores whole aligned words; the caller writes the bytes in front of the first word boundary
// itself (they share a word with the previous record, which the neighbouring lane writes) and
// os_end writes the last partial word bytewise for the same reason.  Inside a run the pending-byte
// count is constant, so a run's inner loop is one load, one funnel shift and one store per word,
// with all loads independent of the stores.  No data-dependent branches: lanes stay converged.
struct OutStream {
  uint32_t wpos;  // shared address (multiple of 4) of the next word to store
  uint32_t lo;    // pending bytes: low `nb` bytes valid, the rest zero
  uint32_t nb;    // 0..3
};

// append the n (0..4) low bytes of v; bytes of v above n must be zero
__device__ __forceinline__ void os_put(OutStream& o, uint32_t v, uint32_t n) {
  const uint32_t sh = o.nb * 8u;
  const uint32_t word = o.lo | (v << sh);
  const uint32_t tot = o.nb + n;
  const bool full = tot >= 4u;
  if (full) sts32(o.wpos, word);                       // predicated store, not a branch
  o.wpos += full ? 4u : 0u;
  o.lo = full ? __funnelshift_l(v, 0u, sh) : word;     // v >> (32 - sh); 0 when sh == 0
  o.nb = tot & 3u;
}

// append [s, s+len) of shared memory (any alignment)
__device__ __forceinline__ void os_run(OutStream& o, uint32_t s, uint32_t len) {
  uint32_t sp = s & ~3u;
  const uint32_t a = s & 3u;
  uint32_t n0 = 4u - a;
  n0 = n0 < len ? n0 : len;
  // first (partial) source word; the load is harmless when len == 0
  uint32_t w = lds32(sp) >> (8u * a);
  w = n0 < 4u ? (w & ((1u << (8u * n0)) - 1u)) : w;
  os_put(o, w, n0);
  len -= n0;
  sp += 4;
  const uint32_t nw = len >> 2;
  if (nw) {
    const uint32_t sh = o.nb * 8u;
    uint32_t prev = sh ? (o.lo << (32u - sh)) : 0u;  // so that funnelshift_l(prev, w, sh) = lo | w << sh
    uint32_t dp = o.wpos;
    uint32_t i = 0;
    for (; i + 8 <= nw; i += 8) {
      const uint32_t w0 = lds32(sp), w1 = lds32(sp + 4), w2 = lds32(sp + 8), w3 = lds32(sp + 12);
      const uint32_t w4 = lds32(sp + 16), w5 = lds32(sp + 20), w6 = lds32(sp + 24), w7 = lds32(sp + 28);
      sts32(dp, __funnelshift_l(prev, w0, sh));
      sts32(dp + 4, __funnelshift_l(w0, w1, sh));
      sts32(dp + 8, __funnelshift_l(w1, w2, sh));
      sts32(dp + 12, __funnelshift_l(w2, w3, sh));
      sts32(dp + 16, __funnelshift_l(w3, w4, sh));
      sts32(dp + 20, __funnelshift_l(w4, w5, sh));
      sts32(dp + 24, __funnelshift_l(w5, w6, sh));
      sts32(dp + 28, __funnelshift_l(w6, w7, sh));
      prev = w7;
      sp += 32;
      dp += 32;
    }
    for (; i < nw; ++i) {
      const uint32_t wi = lds32(sp);
      sts32(dp, __funnelshift_l(prev, wi, sh));
      prev = wi;
      sp += 4;
      dp += 4;
    }
    o.lo = __funnelshift_l(prev, 0u, sh);
    o.wpos = dp;
  }
  const uint32_t r = len & 3u;
  if (r) os_put(o, lds32(sp) & ((1u << (8u * r)) - 1u), r);
}

__device__ __forceinline__ void os_end(const OutStream& o) {
  if (o.nb > 0) sts8(o.wpos, o.lo);
  if (o.nb > 1) sts8(o.wpos + 1, o.lo >> 8);
  if (o.nb > 2) sts8(o.wpos + 2, o.lo >> 16);
}

// One record of output at shared address `d` (any alignment).  `in`-side positions in r are shared
// addresses too.
__device__ __forceinline__ void emit_record(const bk_devprog& pg, const RecView& r, int ms, bool trim, uint32_t d) {
  // The first k0 <= 3 output bytes sit in a word shared with the previous record: write them
  // bytewise, then stream from the word boundary.  They are the first bytes of '@'+head, or of the
  // whole record when it is copied unchanged; a header too short to cover them (< 2 bytes) takes
  // the plain bytewise path below.
  const bool unchanged = ms < 0 && r.fast;
  const uint32_t k0 = (4u - (d & 3u)) & 3u;
  OutStream o;
  uint32_t first_len = unchanged ? r.end - r.beg : 1u + r.head_len;
  uint32_t src = r.beg;
  if (first_len >= 3u) {
    if (k0 > 0) sts8(d, lds8(src));
    if (k0 > 1) sts8(d + 1, lds8(src + 1));
    if (k0 > 2) sts8(d + 2, lds8(src + 2));
    o.wpos = d + k0;
    o.lo = 0;
    o.nb = 0;
    src += k0;
    first_len -= k0;
  } else {
    // degenerate header (rare): plain bytewise path
    uint32_t pos = d;
    auto put_byte = [&](uint32_t c) { sts8(pos++, c); };
    put_byte('@');
    for (uint32_t i = 0; i < r.head_len; ++i) put_byte(lds8(r.beg + 1 + i));
    if (ms >= 0) {
      for (uint32_t c = 0; c < pg.ncaps; ++c) {
        const bk_cap cap = pg.caps[c];
        put_byte(' ');
        for (uint32_t i = 0; i < cap.name_len; ++i) put_byte((uint8_t)cap.name[i]);
        put_byte(':');
        for (uint32_t i = 0; i < cap.len; ++i) put_byte(lds8(r.seq + ms + cap.off + i));
        put_byte(':');
        for (uint32_t i = 0; i < cap.len; ++i)
          put_byte(ms + cap.off + i < r.qual_len ? lds8(r.qual + ms + cap.off + i) : (uint32_t)'!');
      }
    }
    put_byte('\n');
    const uint32_t T = (ms >= 0 && trim) ? pg.total_len : 0u;
    const uint32_t cut = (ms >= 0 && trim) ? (uint32_t)ms : 0u;
    for (uint32_t i = 0; i < r.seq_len; ++i)
      if (i < cut || i >= cut + T) put_byte(lds8(r.seq + i));
    put_byte('\n'); put_byte('+'); put_byte('\n');
    for (uint32_t i = 0; i < r.qual_len; ++i)
      if (i < cut || i >= cut + T) put_byte(lds8(r.qual + i));
    put_byte('\n');
    return;
  }
  // `fast` records already carry the writer's framing, so neighbouring fields travel with their
  // separators as one run ("\n+\n" after seq, "\n" after qual) and an unchanged record is one run.
  os_run(o, src, first_len);
  if (unchanged) {
    os_end(o);
    return;
  }
  if (ms >= 0) {
    for (uint32_t c = 0; c < pg.ncaps; ++c) {  // pattern order (parse.rs:101)
      const bk_cap cap = pg.caps[c];
      // " NAME:"  (parse.rs:166): names are 2 or 3 bytes
      const uint32_t b3 = cap.name_len == 3 ? (uint32_t)(uint8_t)cap.name[2] : (uint32_t)':';
      os_put(o, (uint32_t)' ' | ((uint32_t)(uint8_t)cap.name[0] << 8) | ((uint32_t)(uint8_t)cap.name[1] << 16) |
                    (b3 << 24), 4u);
      os_put(o, cap.name_len == 3 ? (uint32_t)':' : 0u, cap.name_len == 3 ? 1u : 0u);
      os_run(o, r.seq + ms + cap.off, cap.len);
      os_put(o, (uint32_t)':', 1u);
      // qual shorter than seq cannot come out of bk_tokenize; clamp keeps the kernel memory-safe
      const uint32_t qs = ms + cap.off < r.qual_len ? ms + cap.off : r.qual_len;
      const uint32_t qe = ms + cap.off + cap.len < r.qual_len ? ms + cap.off + cap.len : r.qual_len;
      os_run(o, r.qual + qs, qe - qs);
      for (uint32_t i = qe - qs; i < cap.len; ++i) os_put(o, (uint32_t)'!', 1u);
    }
  }
  os_put(o, (uint32_t)'\n', 1u);
  const uint32_t sep = r.fast ? 3u : 0u, eol = r.fast ? 1u : 0u;  // separators copied with the run
  const uint32_t kSep = (uint32_t)'\n' | ((uint32_t)'+' << 8) | ((uint32_t)'\n' << 16);
  if (ms >= 0 && trim) {  // parse.rs:138-148: cut [start0, end0) out of seq and qual
    const uint32_t T = pg.total_len;
    os_run(o, r.seq, ms);
    os_run(o, r.seq + ms + T, r.seq_len - ms - T + sep);
    os_put(o, r.fast ? 0u : kSep, r.fast ? 0u : 3u);
    const uint32_t qs = (uint32_t)ms < r.qual_len ? (uint32_t)ms : r.qual_len;
    const uint32_t qe = (uint32_t)ms + T < r.qual_len ? (uint32_t)ms + T : r.qual_len;
    os_run(o, r.qual, qs);
    os_run(o, r.qual + qe, r.qual_len - qe + eol);
  } else if (r.fast) {
    os_run(o, r.seq, r.end - r.seq);  // seq LF + LF qual LF
  } else {
    os_run(o, r.seq, r.seq_len);
    os_put(o, kSep, 3u);
    os_run(o, r.qual, r.qual_len);
  }
  os_put(o, r.fast ? 0u : (uint32_t)'\n', r.fast ? 0u : 1u);
  os_end(o);
}

// parse::get_reverse_complement as a stand-alone op (API parity; the extract kernel never materialises it)
__global__ void rc_kernel(const uint8_t* in, uint8_t* out, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = complement(in[n - 1 - i]);
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
