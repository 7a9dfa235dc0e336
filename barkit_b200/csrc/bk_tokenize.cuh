// Record-boundary scan of plain FASTQ text on the device: the part of
// seq_io 0.3.2's reader that FastqReader::read_record_set (fastq.rs:166-172) relies on, so that a
// batch can be shipped as raw text and the host tokenizer (bk_fastq.cc, same contract) drops out of
// the end-to-end path.  A record is four lines; its start is the byte after every fourth newline.
//
//   count : newlines per 4 KiB block (16 bytes per thread, SWAR zero-byte test)
//   scan  : exclusive prefix of the block counts (one CTA)
//   emit  : newline number g (0-based, text order) with g % 4 == 3 ends record g / 4:
//           rec_off[g / 4 + 1] = its position + 1
//   check : one thread per record -- '@' and '+' line starts, seq / qual lengths equal after CR
//           stripping (the conditions bk_tokenize reports as BK_E_FORMAT), longest record
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bk {

constexpr int kTokThreads = 256;
constexpr uint32_t kTokChunk = kTokThreads * 16u;  // bytes per block

struct TokResult {
  unsigned long long newlines;  // in the whole text
  unsigned int first_bad;       // smallest index of a malformed record, 0xFFFFFFFF if none
  unsigned int max_rec_bytes;
};

// bit i set iff byte i of the 16 bytes at text[pos..] (clipped to len) is '\n'
__device__ __forceinline__ uint32_t tok_nl_mask(const uint8_t* text, uint64_t pos, uint64_t len) {
  if (pos >= len) return 0;
  const uint4 v = *reinterpret_cast<const uint4*>(text + pos);  // text is 16-byte aligned with >= 16 bytes of slack
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = w[i] ^ 0x0A0A0A0Au;
    const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;  // 0x80 where the byte is '\n' (exact)
    const uint32_t b = z >> 7;
    m |= ((b & 1u) | ((b >> 7) & 2u) | ((b >> 14) & 4u) | ((b >> 21) & 8u)) << (4 * i);
  }
  const uint64_t left = len - pos;
  return left >= 16 ? m : m & ((1u << left) - 1u);
}

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

__global__ void __launch_bounds__(kTokThreads) tok_count_kernel(const uint8_t* text, uint64_t len, uint32_t* block_count) {
  const uint64_t pos = (uint64_t)blockIdx.x * kTokChunk + threadIdx.x * 16u;
  uint32_t excl;
  const uint32_t total = tok_block_sum(__popc(tok_nl_mask(text, pos, len)), &excl);
  if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}

// in-place exclusive scan of block_count[0..nblocks); one CTA walks it 1024 entries at a time
__global__ void __launch_bounds__(1024) tok_scan_kernel(uint32_t* block_count, uint32_t nblocks, TokResult* res) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nblocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long v = i < nblocks ? block_count[i] : 0;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
    const unsigned long long carry = carry_s;
    // newline counts beyond 2^32 cannot occur: the text is < 4 GiB
    if (i < nblocks) block_count[i] = (uint32_t)(carry + wbase + incl - v);
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    res->newlines = carry_s;
    res->first_bad = 0xFFFFFFFFu;
    res->max_rec_bytes = 0;
  }
}

__global__ void __launch_bounds__(kTokThreads) tok_emit_kernel(const uint8_t* text, uint64_t len, const uint32_t* block_base,
                                                               uint32_t* rec_off, uint32_t cap_records) {
  const uint64_t pos = (uint64_t)blockIdx.x * kTokChunk + threadIdx.x * 16u;
  uint32_t m = tok_nl_mask(text, pos, len);
  uint32_t excl;
  tok_block_sum(__popc(m), &excl);
  uint32_t g = block_base[blockIdx.x] + excl;  // number of this thread's first newline
  if (blockIdx.x == 0 && threadIdx.x == 0) rec_off[0] = 0;
  for (; m; m &= m - 1u, ++g) {
    if ((g & 3u) == 3u) {
      const uint32_t rec = (g >> 2) + 1u;
      if (rec <= cap_records) rec_off[rec] = (uint32_t)(pos + (uint32_t)(__ffs(m) - 1) + 1u);
    }
  }
}

// first '\n' in [b, e) or e
__device__ __forceinline__ uint32_t tok_find_nl(const uint8_t* text, uint32_t b, uint32_t e) {
  while (b < e && text[b] != '\n') ++b;
  return b;
}

__global__ void tok_check_kernel(const uint8_t* text, const uint32_t* rec_off, uint32_t n, uint32_t last_end,
                                 TokResult* res) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t b = rec_off[r], e = r + 1 == n ? last_end : rec_off[r + 1];
  bool ok = e > b && text[b] == '@';
  if (ok) {
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
  if (!ok) atomicMin(&res->first_bad, r);
  atomicMax(&res->max_rec_bytes, e - b);
}

}  // namespace bk
