// compact_kernel: the pattern-less mate of a paired batch.
//
// In paired mode a mate without a pattern is only ever copied or dropped (run.rs:149-160: a pair is kept iff the
// other mate's pattern matched, and the unmatched side goes out as it came in, re-framed by the writer,
// fastq.rs:259-263).  That is an ordered compaction of whole records, and it needs nothing the extract kernel has
// on chip.  So a paired launch with exactly one pattern runs as two kernels on the same stream:
//
//   1. extract_kernel<PAIRED = false> on the pattern's mate (same code as a single-end batch), which also leaves
//      the match table (>= 0: matched, i.e. pair kept) and the header length of every read in device memory;
//   2. this kernel on the other mate: sizes from the record offsets, ordered compaction by the same two-level
//      decoupled look-back as the extract kernel (bk_extract.cuh: lookback), and a global -> global copy.
//
// Neither kernel reads the other mate's text, so DRAM traffic stays at one read of the input and one write of
// the output; the hand-over costs 8 bytes per pair.  The staged kernel no longer carries this mate through its
// shared-memory ring (half of its shared memory and a third of its shared-memory traffic), and this kernel is a
// plain streaming copy at full occupancy.
//
// Tile = 256 consecutive records, one CTA of 256 threads, taken in ticket order.
//   sizes : thread t owns record t.  A kept record in the writer's own framing `@h LF seq LF + LF qual LF` goes out
//           verbatim (e - b bytes); that is checked with eight byte probes (parse_probe_g) using the partner read's
//           header length as the guess, the scanning parser otherwise.  Foreign framing (CR line ends, text after
//           '+', no final EOL) is re-framed bytewise by the record's thread.
//   scan  : block scan of the sizes, look-back over the earlier tiles -> global offset of the tile's output.
//   copy  : maximal runs of consecutive verbatim records are contiguous on both sides.  A warp copies the runs of
//           its 32 records, all lanes together: consecutive lanes own consecutive 16-byte aligned destination
//           chunks, load the two aligned source chunks under them (the second is the next lane's first: L1 serves
//           it) and funnel-shift by the run's relative misalignment; the run's ragged ends (< 16 bytes each) are
//           copied bytewise, one lane per byte.  Eight CTAs per SM hide the load latency; nothing is unrolled.
#pragma once
#include "bk_extract.cuh"

namespace bk {

constexpr int kCompactRecs = 256;     // records per tile = threads per CTA
constexpr int kCompactThreads = 256;
#ifndef BK_COMPACT_UNROLL
#define BK_COMPACT_UNROLL 4
#endif
#ifndef BK_POLL_NS
#define BK_POLL_NS 500
#endif
constexpr int kCompactUnroll = BK_COMPACT_UNROLL;  // 16-byte chunks per lane and turn of the copy loop

struct CompactParams {
  const uint8_t* text;
  const uint32_t* off;
  const unsigned long long* kmask;  // the other mate's keep masks, one per tile of ITS kernel: 1 << 63 | bit i = read i stays
  uint32_t kmask_recs;      // reads per such tile
  const uint32_t* hlen;     // the other mate's header lengths (guesses only: read without synchronisation), nullable
  uint8_t* out;
  unsigned long long cap;
  unsigned long long* desc[3];  // tile descriptors, block descriptors, block accumulators (8 bytes per entry, one stream)
  unsigned int* ticket;
  DevResult* result;
  uint32_t n;
  uint32_t ntiles;
  uint32_t text_len;
  uint32_t stream;          // which output stream of the pair this mate is (0 / 1): where its length goes in `result`
};

__device__ __forceinline__ uint4 realign16(const uint4 x, const uint4 y, uint32_t sh) {
  const uint32_t bits = (sh & 3u) * 8u;
  const bool q1 = (sh & 4u) != 0, q2 = (sh & 8u) != 0;
  const uint32_t a0 = q2 ? x.z : x.x, a1 = q2 ? x.w : x.y, a2 = q2 ? y.x : x.z, a3 = q2 ? y.y : x.w;
  const uint32_t a4 = q2 ? y.z : y.x, a5 = q2 ? y.w : y.y;
  const uint32_t b0 = q1 ? a1 : a0, b1 = q1 ? a2 : a1, b2 = q1 ? a3 : a2, b3 = q1 ? a4 : a3, b4 = q1 ? a5 : a4;
  return make_uint4(__funnelshift_r(b0, b1, bits), __funnelshift_r(b1, b2, bits), __funnelshift_r(b2, b3, bits),
                    __funnelshift_r(b3, b4, bits));
}

// CAREFUL = false, the default: every kept record is taken to be in the writer's own framing, so its size is e - b and
// nothing of the text is touched before the copy itself reads it -- framing probes in front of the look-back pulled
// sectors into L2 that the copy, 10-20 us later and behind ~100 MB of other tiles' traffic, had to fetch from DRAM
// again (ncu: 4.06 GB read for 3.0 GB of kept text).  The probes run AFTER the copy, on lines the warp has just
// read; a record that fails them raises bit 2 of `error`, and the host runs the batch's compaction again with
// CAREFUL = true (sizes from the parsed records, foreign framing re-framed), and keeps doing so for that context.
template <bool CAREFUL>
__global__ void __launch_bounds__(kCompactThreads, 6) compact_kernel(const __grid_constant__ CompactParams P) {
  __shared__ uint32_t s_warp[kCompactThreads / 32];
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    // The hand-over from the other mate's extract kernel is its tiles' keep masks, one self-contained 8-byte message
    // per tile: this kernel depends on nothing else of that launch, and a mask that is not there yet is waited for
    // (so the two kernels may also overlap; one warp polls, the CTA's other warps block on the barrier).
    uint32_t tk = 0;
    if (lane == 0) tk = atomicAdd(P.ticket, 1u);
    tk = __shfl_sync(0xffffffffu, tk, 0);
    const uint32_t f = tk * kCompactRecs, l = min(f + (uint32_t)kCompactRecs, P.n) - 1u;
    const uint32_t k0 = f / P.kmask_recs, k1 = l / P.kmask_recs;
    for (uint32_t k = k0 + (uint32_t)lane; k <= k1; k += 32u)
      while (!(ld_relaxed_u64(P.kmask + k) >> 63)) __nanosleep(BK_POLL_NS);
    if (lane == 0) s_tile = tk;
  }
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t r0 = tile * kCompactRecs;
  const uint32_t nrec = min((uint32_t)kCompactRecs, P.n - r0);
  const uint32_t t = threadIdx.x;
  const uint8_t* const text = P.text;

  // ---- sizes ---------------------------------------------------------------------------------------------------
  RecView rv;
  rv.beg = rv.end = rv.fast = rv.head_len = rv.seq = rv.seq_len = rv.qual = rv.qual_len = 0;
  bool kept = false;
  if (t < nrec) {
    const uint32_t b = P.off[r0 + t], e = P.off[r0 + t + 1u];
    const uint32_t hl = P.hlen ? __ldcg(P.hlen + r0 + t) : 0u;
    const uint32_t kt = (r0 + t) / P.kmask_recs, kb = (r0 + t) - kt * P.kmask_recs;
    const unsigned long long km = ld_relaxed_u64(P.kmask + kt);  // (complete: warp 0 saw it)
    kept = (km >> kb) & 1ull;
    if (CAREFUL) {
      if (kept && !(P.hlen && parse_probe_g(text, b, e, hl, rv))) rv = parse_record_g(text, b, e);
    } else {
      rv.beg = b;
      rv.end = e;
      rv.fast = 1;
      rv.head_len = hl;  // (the guess, for the check behind the copy)
    }
  }
  const bool verb = kept && rv.fast;  // goes out verbatim
  const uint32_t size = !kept ? 0u : verb ? rv.end - rv.beg : 1u + rv.head_len + 1u + rv.seq_len + 3u + rv.qual_len + 1u;
  uint32_t incl = size;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  // runs: a verbatim record continues its predecessor's run iff that one is verbatim too (then both sides are contiguous)
  const unsigned vm = __ballot_sync(0xffffffffu, verb);
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kCompactThreads / 32; ++w) {
    const uint32_t v = s_warp[w];
    if (w < warp) wbase += v;
    total += v;
  }
  const uint32_t excl = wbase + incl - size;  // this record's offset inside the tile's output
  // ---- ordered compaction across tiles ---------------------------------------------------------------------------
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(P.desc[0] + tile)), "r"(total | kTileReady) : "memory");
      const unsigned long long old = atomicAdd(P.desc[2] + 2 * (size_t)(tile >> 5), (1ull << 48) | (unsigned long long)total);
      const uint32_t in_grp = min(32u, P.ntiles - ((tile >> 5) << 5));
      if ((uint32_t)(old >> 48) + 1u == in_grp)  // whoever completes a group of 32 tiles publishes its aggregate
        desc_raise(P.desc[1] + 2 * (size_t)(tile >> 5), kFlagAgg | ((old & ((1ull << 48) - 1)) + total));
    }
    __syncwarp();
    const ulonglong2 g = lookback<false>(P.desc[0], P.desc[1], tile, total, 0u, lane);
    if (lane == 0) s_base = g.x;
  }
  __syncthreads();
  const unsigned long long gbase = s_base;
  if (threadIdx.x == 0 && tile == P.ntiles - 1) P.result->out_len[P.stream] = gbase + total;  // the batch's last tile
  if (gbase + total > P.cap) {
    if (threadIdx.x == 0) atomicOr(&P.result->error, 1u);
    return;
  }
  if (total == 0) return;
  // ---- foreign framing: bytewise, by the record's thread -------------------------------------------------------
  uint8_t* const tile_out = P.out + gbase;
  if (kept && !verb) emit_record_g(text, rv.beg, rv.head_len, rv.seq, rv.seq_len, rv.qual, rv.qual_len, tile_out + excl);
  // ---- copy: the runs of this warp's 32 records, all lanes together ---------------------------------------------------
  for (unsigned todo = vm; todo;) {
    const int i0 = __ffs(todo) - 1;                                  // the run's first record ...
    const unsigned above = ~(vm >> i0);
    const int i1 = above ? i0 + __ffs(above) - 2 : 31;              // ... and its last
    todo &= i1 >= 31 ? 0u : ~((2u << i1) - 1u);
    const uint32_t src = __shfl_sync(0xffffffffu, rv.beg, i0);
    const uint32_t d0 = __shfl_sync(0xffffffffu, excl, i0);
    const uint32_t n = __shfl_sync(0xffffffffu, excl + size, i1) - d0;
    const uint8_t* const sp = text + src;
    uint8_t* const dp = tile_out + d0;
    const uintptr_t A = reinterpret_cast<uintptr_t>(dp);
    const uintptr_t D0 = (A + 15u) & ~uintptr_t(15), D1 = (A + n) & ~uintptr_t(15);
    const uintptr_t S0 = reinterpret_cast<uintptr_t>(sp) + (D0 - A);  // source of the first whole chunk
    const uint32_t sh = (uint32_t)(S0 & 15u);
    const uint4* const sa = reinterpret_cast<const uint4*>(S0 & ~uintptr_t(15));
    // One turn = kCompactUnroll x 512 destination bytes.  Every source chunk is loaded exactly once: a lane's second
    // chunk is its neighbour's first and comes over by shuffle (lane 31's is lane 0's of the next 512 bytes, which
    // lane 0 loads as one extra chunk per turn).  Loading both chunks per lane doubled the requests to L2, and a fifth
    // of the doubles went all the way to DRAM (ncu: 3.8 GB read for 3.0 GB of kept text).
#pragma unroll 1
    for (uintptr_t Dw = D0; Dw < D1; Dw += 512u * kCompactUnroll) {
      uint4 x[kCompactUnroll + 1];
#pragma unroll
      for (int u = 0; u <= kCompactUnroll; ++u) {
        const uintptr_t Au = Dw + 512u * u + 16u * (uint32_t)lane;
        const bool ld = (u < kCompactUnroll || lane == 0) && Au < D1 + 16u;  // (one chunk past the last: its low bytes end the run)
        x[u] = ld ? __ldg(sa + ((Au - D0) >> 4)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kCompactUnroll; ++u) {
        const uint4 v = lane == 0 ? x[u + 1] : x[u];
        const int from = (lane + 1) & 31;
        uint4 y;
        y.x = __shfl_sync(0xffffffffu, v.x, from);
        y.y = __shfl_sync(0xffffffffu, v.y, from);
        y.z = __shfl_sync(0xffffffffu, v.z, from);
        y.w = __shfl_sync(0xffffffffu, v.w, from);
        const uintptr_t Au = Dw + 512u * u + 16u * (uint32_t)lane;
        if (Au < D1) *reinterpret_cast<uint4*>(Au) = realign16(x[u], y, sh);
      }
    }
    uint32_t pad = (uint32_t)((16u - (A & 15u)) & 15u);
    pad = pad < n ? pad : n;
    const uint32_t tl = (uint32_t)(D1 - A);  // (wraps when the run ends before its first aligned chunk)
    const uint32_t tlo = (int32_t)tl > (int32_t)pad ? tl : pad;
    const uint32_t j = lane < 16 ? (uint32_t)lane : tlo + ((uint32_t)lane - 16u);
    if (lane < 16 ? j < pad : j < n) dp[j] = (uint8_t)__ldg(sp + j);
  }
  if (!CAREFUL && kept) {  // the framing that the sizes assumed (the lines are in L1 / L2 now)
    RecView chk;
    if (!(P.hlen && parse_probe_g(text, rv.beg, rv.end, rv.head_len, chk)) && !parse_record_g(text, rv.beg, rv.end).fast)
      atomicOr(&P.result->error, 4u);
  }
}

}  // namespace bk
