// Device code of the extract hot path (sm_100a).
//
// One launch = one batch = parse_se_reads / parse_pe_reads (run.rs:63-80, :163-195) + the writer's
// framing (fastq.rs:259-263) for every record of the batch, single pass over HBM:
//
//   tile  = 32 consecutive records (pairs); one CTA owns a tile, one warp per mate, lane i owns record i
//   load  : the tile's contiguous FASTQ text range, one 1-D TMA bulk copy per mate -> shared memory
//   parse : lane finds its record's four lines (SWAR newline scan in shared memory)
//   match : lane evaluates the lowered program (bk_program.h) at ascending starts -> leftmost match
//           (pattern.rs:225-230), optional reverse-complement retry (parse.rs:58-68)
//   size  : output bytes per record, warp scan, decoupled look-back across tiles -> exact global
//           output offset (ordered, gap-free: run.rs:79,194 + fastq.rs:265-271)
//   emit  : lane assembles `@head NAME:seq:qual...\nseq'\n+\nqual'\n` (parse.rs:92-171) in a shared
//           output tile laid out with the global offset's 16-byte phase
//   store : aligned middle by TMA bulk store, ragged head/tail bytes by the lanes
// This header holds the device building blocks; the kernel itself is in bk_extract.cuh.
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 1-D TMA: shared -> global
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
               "r"(smem_u32(src)), "r"(bytes)
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

// First '\n' at or after `from` (shared-memory byte offsets), or `end` when there is none.
__device__ __forceinline__ uint32_t find_nl(const uint8_t* s, uint32_t from, uint32_t end) {
  uint32_t p = from & ~3u;
  uint32_t w = *reinterpret_cast<const uint32_t*>(s + p) ^ 0x0A0A0A0Au;
  uint32_t m = (w - 0x01010101u) & ~w & 0x80808080u;
  m &= 0xFFFFFFFFu << ((from & 3u) * 8u);
  while (m == 0) {
    p += 4;
    if (p >= end) return end;
    w = *reinterpret_cast<const uint32_t*>(s + p) ^ 0x0A0A0A0Au;
    m = (w - 0x01010101u) & ~w & 0x80808080u;
  }
  uint32_t pos = p + ((__ffs(m) - 1) >> 3);
  return pos < end ? pos : end;
}

struct RecView {    // seq_io field definitions: lines without EOL and without a trailing CR
  uint32_t beg;     // '@'
  uint32_t end;     // one past the record's last byte
  uint32_t fast;    // layout is exactly `@head LF seq LF + LF qual LF` (the writer's own framing)
  uint32_t head;    // first byte after '@'
  uint32_t head_len;
  uint32_t seq;
  uint32_t seq_len;
  uint32_t qual;
  uint32_t qual_len;
};

__device__ __forceinline__ uint32_t strip_cr(const uint8_t* s, uint32_t b, uint32_t e) {
  return (e > b && s[e - 1] == '\r') ? e - 1 : e;
}

__device__ __forceinline__ RecView parse_record(const uint8_t* s, uint32_t b, uint32_t e) {
  RecView r;
  uint32_t nl1 = find_nl(s, b, e);
  r.beg = b;
  r.end = e;
  r.fast = 0;
  r.head = b + 1;
  r.head_len = strip_cr(s, b + 1, nl1 > b ? nl1 : b + 1) - (b + 1);
  uint32_t sb = nl1 + 1 < e ? nl1 + 1 : e;
  // Fast path for the common layout `seq LF + LF qual LF`: then e - sb = 2L + 4.  A record of any
  // other shape (CR line ends, text after '+', missing final LF) cannot pass the three probes
  // with the L this formula gives (newlines only exist at the three line ends), so it falls
  // through to the scanning path.
  {
    const uint32_t rem = e - sb;
    if (rem >= 4 && !(rem & 1u)) {
      const uint32_t L = (rem - 4) >> 1;
      if (s[sb + L] == '\n' && s[sb + L + 1] == '+' && s[sb + L + 2] == '\n' && s[e - 1] == '\n' &&
          (L == 0 || (s[sb + L - 1] != '\r' && s[e - 2] != '\r'))) {
        r.seq = sb;
        r.seq_len = L;
        r.qual = sb + L + 3;
        r.qual_len = L;
        r.fast = (b + 1 + r.head_len == nl1);  // and no CR on the header line
        return r;
      }
    }
  }
  uint32_t nl2 = find_nl(s, sb, e);
  r.seq = sb;
  r.seq_len = strip_cr(s, sb, nl2) - sb;
  uint32_t pb = nl2 + 1 < e ? nl2 + 1 : e;
  uint32_t nl3 = find_nl(s, pb, e);
  uint32_t qb = nl3 + 1 < e ? nl3 + 1 : e;
  uint32_t nl4 = find_nl(s, qb, e);
  r.qual = qb;
  r.qual_len = strip_cr(s, qb, nl4) - qb;
  return r;
}

// ---- matching --------------------------------------------------------------------------------------

// parse.rs:12-32 TRANSLATION_TABLE
__device__ __forceinline__ uint8_t complement(uint8_t c) {
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
__device__ __forceinline__ uint8_t fetch(const uint8_t* seq, uint32_t L, uint32_t j) {
  if (RC) return complement(seq[L - 1 - j]);
  return seq[j];
}

// All compare runs of the program at start s.  Branch-free inside a run so that the lanes of a warp
// (one read each) stay in lock step: a data-dependent early exit would split the warp and replay
// the remaining runs once per split.  `tab[c]` = bit k set iff class k contains byte c (0 for
// c >= 0x80).  EARLY_OUT (unanchored search, lanes already diverge on s) stops at the first
// failed run.
template <bool RC, bool EARLY_OUT>
__device__ __forceinline__ bool match_at(const bk_devprog& pg, const uint8_t* tab, const uint8_t* seq,
                                         uint32_t L, uint32_t s) {
  bool ok = true;
  for (uint32_t o = 0; o < pg.nops; ++o) {
    const bk_op op = pg.ops[o];
    const uint32_t base = s + op.off;
    if (op.kind == BK_OP_LIT) {
      uint32_t mism = 0, bad = 0;
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
      for (uint32_t i = 0; i < op.len; ++i) acc &= tab[fetch<RC>(seq, L, base + i)];
      ok = ok && ((acc >> op.idx) & 1u);
    }
    if (EARLY_OUT && !ok) return false;
  }
  return ok;
}

// leftmost match start or -1
template <bool RC>
__device__ __forceinline__ int match_read(const bk_devprog& pg, const uint8_t* tab, const uint8_t* seq,
                                          uint32_t L) {
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

// Byte-granular shared->shared copy by ONE lane: ragged head until the destination is word aligned,
// then aligned destination words funnel-shifted out of aligned source words (4 independent loads
// in flight per step), then a ragged tail.  Reads at most 7 bytes past the source run, which stays
// inside the staged tile (+ slack).  `dst` and `src` never alias (input tile vs output tile).
__device__ __forceinline__ uint32_t copy_run(uint8_t* __restrict__ dst, uint32_t d,
                                             const uint8_t* __restrict__ src, uint32_t s, uint32_t len) {
  const uint32_t end = d + len;
  uint32_t h = (4u - (d & 3u)) & 3u;
  h = h < len ? h : len;
  for (uint32_t i = 0; i < h; ++i) dst[d + i] = src[s + i];
  d += h;
  s += h;
  len -= h;
  const uint32_t nw = len >> 2;
  if (nw) {
    const uint32_t* __restrict__ sp = reinterpret_cast<const uint32_t*>(src + (s & ~3u));
    uint32_t* __restrict__ dp = reinterpret_cast<uint32_t*>(dst + d);
    const uint32_t sh = (s & 3u) * 8u;
    uint32_t lo = sp[0];
    uint32_t i = 0;
    for (; i + 4 <= nw; i += 4) {
      const uint32_t w1 = sp[i + 1], w2 = sp[i + 2], w3 = sp[i + 3], w4 = sp[i + 4];
      dp[i] = __funnelshift_r(lo, w1, sh);
      dp[i + 1] = __funnelshift_r(w1, w2, sh);
      dp[i + 2] = __funnelshift_r(w2, w3, sh);
      dp[i + 3] = __funnelshift_r(w3, w4, sh);
      lo = w4;
    }
    for (; i < nw; ++i) {
      const uint32_t hi = sp[i + 1];
      dp[i] = __funnelshift_r(lo, hi, sh);
      lo = hi;
    }
  }
  d += nw * 4;
  s += nw * 4;
  const uint32_t t = len & 3u;
  for (uint32_t i = 0; i < t; ++i) dst[d + i] = src[s + i];
  return end;
}

// output bytes of a kept record
__device__ __forceinline__ uint32_t out_bytes(const bk_devprog& pg, const RecView& r, int ms, bool trim) {
  uint32_t sl = r.seq_len, ql = r.qual_len, g = 0;
  if (ms >= 0) {
    g = pg.hdr_growth;
    if (trim) {
      uint32_t T = pg.total_len;
      sl -= T;  // match lies inside seq
      uint32_t qs = (uint32_t)ms < ql ? (uint32_t)ms : ql;
      uint32_t qe = (uint32_t)ms + T < ql ? (uint32_t)ms + T : ql;
      ql -= (qe - qs);
    }
  }
  return 1 + r.head_len + g + 1 + sl + 3 + ql + 1;
}

__device__ __forceinline__ void emit_record(const bk_devprog& pg, const uint8_t* __restrict__ in,
                                            const RecView& r, int ms, bool trim,
                                            uint8_t* __restrict__ out, uint32_t d) {
  // `fast` records already carry the writer's framing, so neighbouring fields travel with their
  // separators as one run ("\n+\n" after seq, "\n" after qual) and an unchanged record is one run.
  if (ms < 0 && r.fast) {
    copy_run(out, d, in, r.beg, r.end - r.beg);
    return;
  }
  d = copy_run(out, d, in, r.beg, 1 + r.head_len);  // '@' + head
  if (ms >= 0) {
    for (uint32_t c = 0; c < pg.ncaps; ++c) {  // pattern order (parse.rs:101)
      const bk_cap cap = pg.caps[c];
      out[d++] = ' ';
      for (uint32_t i = 0; i < cap.name_len; ++i) out[d++] = cap.name[i];
      out[d++] = ':';
      d = copy_run(out, d, in, r.seq + ms + cap.off, cap.len);
      out[d++] = ':';
      // qual shorter than seq cannot come out of bk_tokenize; clamp keeps the kernel memory-safe
      uint32_t qs = ms + cap.off < r.qual_len ? ms + cap.off : r.qual_len;
      uint32_t qe = ms + cap.off + cap.len < r.qual_len ? ms + cap.off + cap.len : r.qual_len;
      d = copy_run(out, d, in, r.qual + qs, qe - qs);
      for (uint32_t i = qe - qs; i < cap.len; ++i) out[d++] = '!';
    }
  }
  out[d++] = '\n';
  const uint32_t sep = r.fast ? 3u : 0u, eol = r.fast ? 1u : 0u;  // separators copied with the run
  if (ms >= 0 && trim) {  // parse.rs:138-148: cut [start0, end0) out of seq and qual
    const uint32_t T = pg.total_len;
    d = copy_run(out, d, in, r.seq, ms);
    d = copy_run(out, d, in, r.seq + ms + T, r.seq_len - ms - T + sep);
    if (!r.fast) { out[d++] = '\n'; out[d++] = '+'; out[d++] = '\n'; }
    const uint32_t qs = (uint32_t)ms < r.qual_len ? (uint32_t)ms : r.qual_len;
    const uint32_t qe = (uint32_t)ms + T < r.qual_len ? (uint32_t)ms + T : r.qual_len;
    d = copy_run(out, d, in, r.qual, qs);
    d = copy_run(out, d, in, r.qual + qe, r.qual_len - qe + eol);
  } else if (r.fast) {
    d = copy_run(out, d, in, r.seq, r.end - r.seq);  // seq LF + LF qual LF
  } else {
    d = copy_run(out, d, in, r.seq, r.seq_len);
    out[d++] = '\n'; out[d++] = '+'; out[d++] = '\n';
    d = copy_run(out, d, in, r.qual, r.qual_len);
  }
  if (!r.fast) out[d++] = '\n';
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
