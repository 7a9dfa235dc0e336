// Record-boundary scan of plain FASTQ text on the device: the part of
// seq_io 0.3.2's reader that FastqReader::read_record_set (fastq.rs:166-172) relies on, so that a
// batch can be shipped as raw text and the host tokenizer (bk_fastq.cc, same contract) drops out of
// the end-to-end path.  A record is four lines; its start is the byte after every fourth newline.
//
// One pass over the text (tok_fused_kernel): newlines per 64 KiB chunk (SWAR zero-byte test), their
// numbers in the whole text by a decoupled look-back over the chunks, rec_off[g / 4 + 1] = position + 1
// of every newline g with g % 4 == 3, and the record checks -- '@' and '+' line starts, seq / qual
// lengths equal after CR stripping (the conditions bk_tokenize reports as BK_E_FORMAT), longest record.
// tok_check_kernel is the same check one thread per record, used for a final record without EOL.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bk_extract.cuh"  // lookback, descriptor helpers

namespace bk {

#ifndef BK_TOK_THREADS
#define BK_TOK_THREADS 256
#endif
#ifndef BK_TOK_W
#define BK_TOK_W 4
#endif
constexpr int kTokThreads = BK_TOK_THREADS;
constexpr int kTokW = BK_TOK_W;                                   // 64-byte units per thread (contiguous)
constexpr uint32_t kTokSpan = 64u * kTokW;                       // bytes per thread
constexpr uint32_t kTokChunk = kTokThreads * kTokSpan;           // bytes per block

struct TokResult {
  unsigned long long newlines;  // in the whole text
  unsigned int first_bad;       // smallest index of a malformed record, 0xFFFFFFFF if none
  unsigned int max_rec_bytes;
};

__device__ __forceinline__ uint32_t tok_block_sum(uint32_t v, uint32_t* excl) {  // block-wide sum; *excl = sum over lower threads
  __shared__ uint32_t warp_tot[kTokThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kTokThreads / 32; ++w) {
    const uint32_t t = warp_tot[w];
    if (w < warp) base += t;
    total += t;
  }
  __syncthreads();
  *excl = base + incl - v;
  return total;
}

// ---- single pass: count + scan + emit + check ---------------------------------------------------------
//
// One CTA per 64 KiB of text (256 threads x 256 contiguous bytes each: the block scan, the look-back hand-over and the
// record logic are paid per thread and per chunk, not per byte), taken in ticket order.  The chunk is read once: a warp
// reads its threads' 8 KiB with coalesced 16-byte loads, every lane turns its piece into 16 newline flags, and the
// flags change hands through shared memory so that every thread holds the masks of ITS 256 bytes; a block scan numbers
// the chunk's newlines, a decoupled look-back over the earlier chunks' counts (one 8-byte descriptor per chunk:
// aggregate, then inclusive prefix) gives the number of the chunk's first newline in the whole text, and the thread
// that owns newline g with g % 4 == 3 writes rec_off[g / 4 + 1] and validates record g / 4 -- its other line ends are
// found by walking the chunk's masks backwards in shared memory; the one record per chunk that began in an earlier
// chunk is checked by warp 0 from global memory.
// Geometry measured on 10 M records of 329 bytes (tools/ab.sh, wall time around bk_tokenize_device): 128 threads 3.57,
// 256 threads 3.83, 512 threads 3.86-3.92, 1024 threads 2.98 TB/s; 512 bytes per thread 3.03; 8 / 16 loads in flight
// per lane instead of 4: 3.52-3.69 / 3.26.  256 threads keeps 512 chunks in a 32 MiB batch of the file runtime.

// descriptor words: nblocks + 2 chunk descriptors, then 2 per group of 32 chunks (aggregate / prefix), then 2 per group (accumulators)
__host__ __device__ inline size_t tok_desc_words(uint32_t nblocks) { return (size_t)nblocks + 2 + 4 * ((size_t)nblocks / 32 + 1); }

__global__ void tok_init_kernel(TokResult* res, unsigned long long* desc, uint32_t nblocks, unsigned int* ticket) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < tok_desc_words(nblocks)) desc[i] = 0;
  if (i == 0) {
    res->newlines = 0;
    res->first_bad = 0xFFFFFFFFu;
    res->max_rec_bytes = 0;
    *ticket = 0;
  }
}

// the record [s, p4] with line ends p1 < p2 < p3 < p4: the BK_E_FORMAT conditions of bk_tokenize
// (text < 4 GiB: positions are 32-bit)
__device__ __forceinline__ bool tok_record_ok(const uint8_t* text, uint32_t s, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4) {
  uint32_t sl = p2 - (p1 + 1u), ql = p4 - (p3 + 1u);
  const uint32_t c0 = text[s], c1 = text[p2 + 1u], c2 = text[p2 - 1u], c3 = text[p4 - 1u];  // (the chunk was just read: L1 hits)
  if (sl && c2 == '\r') --sl;
  if (ql && c3 == '\r') --ql;
  return c0 == '@' && c1 == '+' && sl == ql;
}

// '\n' flags of a word's four bytes in bits 0..3 (junk above): exact zero-byte test of w ^ "\n\n\n\n", then ONE high
// multiply gathers bits 7, 15, 23, 31 into bits 32..35 of the product (2^25 + 2^18 + 2^11 + 2^4: no two partial products
// land on the same bit, so nothing carries)
__device__ __forceinline__ uint32_t tok_nib(uint32_t w) {
  const uint32_t x = w ^ 0x0A0A0A0Au;
  const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
  return __umulhi(z, 0x02040810u);
}
// the flags of 16 more bytes shifted in at the top of acc (a funnel shift takes the low nibble of every tok_nib)
__device__ __forceinline__ uint32_t tok_shift16(uint32_t acc, const uint4 v) {
  acc = __funnelshift_r(acc, tok_nib(v.x), 4);
  acc = __funnelshift_r(acc, tok_nib(v.y), 4);
  acc = __funnelshift_r(acc, tok_nib(v.z), 4);
  return __funnelshift_r(acc, tok_nib(v.w), 4);
}
// 16-bit mask of the '\n' bytes of a 16-byte piece
__device__ __forceinline__ uint32_t tok_mask16(const uint4 v) { return tok_shift16(0u, v) >> 16; }

constexpr int kTokM = 2 * kTokW;  // 32-byte units per thread, one 32-bit newline mask each
#ifndef BK_TOK_INFLIGHT
#define BK_TOK_INFLIGHT 4
#endif
#ifndef BK_TOK_MINBLOCKS
#define BK_TOK_MINBLOCKS (1536 / BK_TOK_THREADS)
#endif
constexpr int kTokInflight = BK_TOK_INFLIGHT;  // 16-byte loads a lane has in flight

__global__ void __launch_bounds__(kTokThreads, BK_TOK_MINBLOCKS) tok_fused_kernel(const uint8_t* text, uint32_t len, uint32_t* rec_off,
                                                                 uint32_t cap_records, unsigned long long* desc,
                                                                 unsigned int* ticket, uint32_t nblocks, TokResult* res) {
  __shared__ uint32_t s_bid;
  __shared__ uint32_t s_base;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_bid = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t bid = s_bid;
  const uint32_t cbase = bid * kTokChunk;           // (< 2^32: the text is)
  const uint32_t tpos0 = cbase + threadIdx.x * kTokSpan;  // this thread's contiguous bytes
  // The text allocation is 16-byte aligned with >= 16 bytes of slack behind len; pieces at or behind len are not read.
  // A thread's 256 contiguous bytes read by that thread alone are 32 different cache lines per load instruction of its
  // warp, and the L1's tag stage then bounds the kernel (2.9 TB/s).  So a warp reads its threads' 8 KiB together -- lane
  // l takes the 16-byte piece 32 j + l: 512 contiguous bytes = four lines per instruction --, every lane turns its
  // piece into 16 newline flags, and the flags change hands through shared memory (2 bytes per piece): thread t then
  // finds the masks of ITS bytes as 32 contiguous bytes there.
  __shared__ __align__(16) uint16_t pmask[kTokChunk / 16];
  constexpr int kPieces = (int)kTokSpan / 16;  // pieces per thread = load instructions per lane
  const uint32_t wfirst = (uint32_t)(threadIdx.x >> 5) * 32u * kPieces;  // the warp's first piece in the chunk
  const bool full = cbase + kTokChunk <= len;  // (all chunks but the last: no bounds to look at)
  if (full) {
#pragma unroll
    for (int j0 = 0; j0 < kPieces; j0 += kTokInflight) {
      uint4 v[kTokInflight];
#pragma unroll
      for (int j = 0; j < kTokInflight; ++j) v[j] = *reinterpret_cast<const uint4*>(text + cbase + 16u * (wfirst + 32u * (j0 + j) + lane));
#pragma unroll
      for (int j = 0; j < kTokInflight; ++j) pmask[wfirst + 32u * (j0 + j) + lane] = (uint16_t)tok_mask16(v[j]);
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < kPieces; ++j) {
      const uint32_t pos = cbase + 16u * (wfirst + 32u * j + lane);
      uint32_t m = 0;
      if (pos < len) {
        m = tok_mask16(*reinterpret_cast<const uint4*>(text + pos));
        if (pos + 16u > len) m &= (1u << (len - pos)) - 1u;  // only a piece straddling len needs it
      }
      pmask[wfirst + 32u * j + lane] = (uint16_t)m;
    }
  }
  __syncwarp();
  uint32_t mw[kTokM];
  uint32_t mine = 0;
  static_assert(kTokM % 4 == 0, "a thread fetches its masks with 16-byte loads");
#pragma unroll
  for (int i = 0; i < kTokM / 4; ++i) {
    const uint4 q = *reinterpret_cast<const uint4*>(&pmask[wfirst + (uint32_t)kPieces * lane + 8u * i]);
    mw[4 * i] = q.x; mw[4 * i + 1] = q.y; mw[4 * i + 2] = q.z; mw[4 * i + 3] = q.w;
    mine += __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
  }
  uint32_t excl;
  const uint32_t total = tok_block_sum(mine, &excl);
  // ---- number of the chunk's first newline: the two-level decoupled look-back of the extract kernel
  // (bk_extract.cuh: chunk descriptors within a group of 32, group descriptors across).  A flat look-back
  // walks every chunk still in flight -- ~1200 of them, 32 per L2 round trip -- and its own latency keeps
  // more chunks in flight.
  if (threadIdx.x < 32) {
    unsigned long long* const tdesc = desc;
    unsigned long long* const bdesc = desc + nblocks + 2;
    unsigned long long* const acc = bdesc + 2 * ((size_t)nblocks / 32 + 1);
    if (lane == 0) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(tdesc + bid)), "r"(total | kTileReady) : "memory");
      const unsigned long long old = atomicAdd(acc + 2 * (size_t)(bid >> 5), (1ull << 48) | (unsigned long long)total);
      const uint32_t in_grp = min(32u, nblocks - ((bid >> 5) << 5));
      if ((uint32_t)(old >> 48) + 1u == in_grp)  // whoever completes a group publishes its aggregate
        desc_raise(bdesc + 2 * (size_t)(bid >> 5), kFlagAgg | ((old & ((1ull << 48) - 1)) + total));
    }
    __syncwarp();
    const ulonglong2 g = lookback<false>(tdesc, bdesc, bid, total, 0u, lane);
    if (lane == 0) {
      s_base = (uint32_t)g.x;  // (< 2^32 newlines in < 4 GiB)
      if (bid == nblocks - 1) res->newlines = g.x + total;
      if (bid == 0) rec_off[0] = 0;
    }
  }
  // ---- record ends --------------------------------------------------------------------------------------------
  // The thread's newlines have the numbers base + excl ...; every fourth ends a record.  Such a newline is located in
  // the thread's masks, and the record's other line ends -- the four newlines in front of it -- by walking the chunk's
  // masks backwards in shared memory (pmask read as 32-byte words: a 329-byte record is ten of them).  Nothing is kept
  // per newline.
  __syncthreads();
  const uint32_t base = s_base;
  const uint32_t* const pm32 = reinterpret_cast<const uint32_t*>(pmask);
  uint32_t bad = 0xFFFFFFFFu, longest = 0;
  for (uint32_t i = (3u - (base + excl)) & 3u; i < mine; i += 4u) {
    // my i-th newline: its word, then its bit
    uint32_t skip = i, word = 0, uu = 0;
    bool have = false;
#pragma unroll
    for (int u = 0; u < kTokM; ++u) {
      const uint32_t c = __popc(mw[u]);
      if (!have && skip < c) { word = mw[u]; uu = (uint32_t)u; have = true; }
      if (!have) skip -= c;
    }
    for (; skip; --skip) word &= word - 1u;
    const uint32_t bit = (uint32_t)__ffs((int)word) - 1u;
    uint32_t widx = threadIdx.x * kTokM + uu;  // the word's index in the chunk
    const uint32_t p4 = cbase + 32u * widx + bit;
    const uint32_t g = base + excl + i, rec = g >> 2;
    if (rec >= cap_records) continue;
    rec_off[rec + 1u] = p4 + 1u;
    // the four newlines before it, nearest first: q3, q2, q1, q0 (chunk-relative)
    uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0, found = 0;
    uint32_t w = pm32[widx] & ((1u << bit) - 1u);
    for (;;) {
      while (w && found < 4u) {
        const uint32_t hb = 31u - (uint32_t)__clz(w);
        q3 = q2; q2 = q1; q1 = q0; q0 = 32u * widx + hb;
        w &= ~(1u << hb);
        ++found;
      }
      if (found == 4u || widx == 0u) break;
      w = pm32[--widx];
    }
    uint32_t s;
    if (found == 4u) {
      s = cbase + q0 + 1u;
    } else if (g == 3u && found == 3u) {  // the text's first record
      q3 = q2; q2 = q1; q1 = q0;
      s = 0u;
    } else {
      continue;  // the record began in an earlier chunk: warp 0 looks at it below
    }
    const bool ok = tok_record_ok(text, s, cbase + q1, cbase + q2, cbase + q3, p4);
    if (!ok && rec < bad) bad = rec;
    longest = max(longest, p4 + 1u - s);
  }
  // ---- the record that straddles the chunk's start: its end is the chunk's newline kb < 4, its other
  // newlines are the chunk's first kb and the last 4 - kb before the chunk.  Warp 0 finds those together,
  // 512 bytes of text per step (one L2 round trip), instead of one thread walking ~300 dependent bytes.
  if (threadIdx.x < 32 && bid > 0) {
    const uint32_t kb = (3u - (base & 3u)) & 3u;
    const uint32_t rec = (base + kb) >> 2;
    if (kb < total && rec < cap_records) {  // (warp-uniform)
      uint32_t first_nl[4];  // the chunk's first kb + 1 newlines (chunk-relative)
      {
        uint32_t n = 0, wi = 0, w = pm32[0];
        while (n <= kb) {
          while (!w) w = pm32[++wi];  // (total > kb: they exist)
          first_nl[n++] = 32u * wi + (uint32_t)__ffs((int)w) - 1u;
          w &= w - 1u;
        }
      }
      uint32_t before[4];  // the four newlines before the record's last one, nearest first
      uint32_t found = 0;
      for (; found < kb; ++found) before[found] = cbase + first_nl[kb - 1u - found];
      uint32_t wend = cbase;
      while (found < 4u && wend > 0) {
        const uint32_t wstart = wend >= 512u ? wend - 512u : 0u;  // cbase is a multiple of 512
        uint32_t mm = 0;
        const uint32_t pos = wstart + 16u * lane;
        if (pos < wend) mm = tok_mask16(*reinterpret_cast<const uint4*>(text + pos));
        for (int l = 31; l >= 0 && found < 4u; --l) {  // every lane keeps the same list
          uint32_t ml = __shfl_sync(0xffffffffu, mm, l);
          while (ml && found < 4u) {
            const uint32_t bit = 31u - (uint32_t)__clz(ml);
            before[found++] = wstart + 16u * l + bit;
            ml &= ~(1u << bit);
          }
        }
        wend = wstart;
      }
      if (lane == 0) {
        const uint32_t p4 = cbase + first_nl[kb];
        const uint32_t s = found == 4u ? before[3] + 1u : 0u;
        const bool ok = found == 4u && tok_record_ok(text, s, before[2], before[1], before[0], p4);
        if (!ok && rec < bad) bad = rec;
        longest = max(longest, p4 + 1u - s);
      }
    }
  }
  bad = __reduce_min_sync(0xffffffffu, bad);
  longest = __reduce_max_sync(0xffffffffu, longest);
  if (lane == 0) {
    if (bad != 0xFFFFFFFFu) atomicMin(&res->first_bad, bad);
    if (longest > *reinterpret_cast<volatile unsigned int*>(&res->max_rec_bytes)) atomicMax(&res->max_rec_bytes, longest);
  }
}

// rec_off[n] = end (a last record without line end: no newline writes its end offset)
__global__ void tok_set_end_kernel(uint32_t* rec_off, uint32_t n, uint32_t end) {
  if (threadIdx.x == 0 && blockIdx.x == 0) rec_off[n] = end;
}

// first '\n' in [b, e) or e
__device__ __forceinline__ uint32_t tok_find_nl(const uint8_t* text, uint32_t b, uint32_t e) {
  while (b < e && text[b] != '\n') ++b;
  return b;
}

// records [first, n) one thread each (the general form of the check: used for a final record without EOL)
__global__ void tok_check_kernel(const uint8_t* text, const uint32_t* rec_off, uint32_t first, uint32_t n, uint32_t last_end,
                                 TokResult* res) {
  const uint32_t r = first + blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < n;
  const uint32_t b = live ? rec_off[r] : 0, e = !live ? 0 : r + 1 == n ? last_end : rec_off[r + 1];
  bool ok = !live || (e > b && text[b] == '@');
  if (ok && live) {
    const uint32_t nl1 = tok_find_nl(text, b, e);
    const uint32_t sb = nl1 + 1 < e ? nl1 + 1 : e;
    // the writer's own framing `seq LF + LF qual LF` in closed form first, the general scan otherwise
    const uint32_t rem = e - sb, L = (rem - 4u) >> 1;
    const bool fast = rem >= 4u && !(rem & 1u) && text[sb + L] == '\n' && text[sb + L + 1] == '+' &&
                      text[sb + L + 2] == '\n' && text[e - 1] == '\n' &&
                      (L == 0 || (text[sb + L - 1] != '\r' && text[e - 2] != '\r'));
    if (!fast) {
      const uint32_t nl2 = tok_find_nl(text, sb, e);
      const uint32_t pb = nl2 + 1 < e ? nl2 + 1 : e;
      const uint32_t nl3 = tok_find_nl(text, pb, e);
      const uint32_t qb = nl3 + 1 < e ? nl3 + 1 : e;
      const uint32_t nl4 = tok_find_nl(text, qb, e);
      uint32_t sl = nl2 - sb, ql = nl4 - qb;
      if (sl && text[nl2 - 1] == '\r') --sl;
      if (ql && text[nl4 - 1] == '\r') --ql;
      ok = nl2 < e && nl3 < e && pb < e && text[pb] == '+' && sl == ql;
    }
  }
  // one atomic per warp, and only when it changes the result
  const uint32_t bad = __reduce_min_sync(0xffffffffu, ok ? 0xFFFFFFFFu : r);
  const uint32_t longest = __reduce_max_sync(0xffffffffu, e - b);
  if ((threadIdx.x & 31) == 0) {
    if (bad != 0xFFFFFFFFu) atomicMin(&res->first_bad, bad);
    if (longest > *reinterpret_cast<volatile unsigned int*>(&res->max_rec_bytes)) atomicMax(&res->max_rec_bytes, longest);
  }
}

}  // namespace bk
