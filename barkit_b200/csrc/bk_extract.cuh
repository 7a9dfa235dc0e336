// The extract kernel: ordered compaction (decoupled look-back) + per-tile pipeline.
// Building blocks (parse / match / emit / PTX wrappers) are in bk_kernels.cuh.
#pragma once
#include "bk_kernels.cuh"

namespace bk {

// ---- ordered compaction: decoupled look-back over tiles (one warp per output stream) -----------------

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kLookWindows = 4;  // 32-descriptor windows fetched per round (independent loads in flight)

// Exclusive prefix of this tile's aggregate over all earlier tiles; publishes aggregate, then prefix.
// Several hundred tiles are in flight chip-wide, so the nearest tile that already holds an inclusive
// prefix is typically a few windows back: fetch kLookWindows windows per round trip instead of one.
__device__ __forceinline__ unsigned long long lookback(unsigned long long* desc, uint32_t tile,
                                                       unsigned long long agg, int lane) {
  if (tile == 0) {
    if (lane == 0) st_relaxed_u64(desc, kFlagPrefix | agg);
    return 0;
  }
  if (lane == 0) st_relaxed_u64(desc + tile, kFlagAgg | agg);
  unsigned long long excl = 0;
  int look = (int)tile - 1;
  bool done = false;
  while (!done) {
    unsigned long long d[kLookWindows];
#pragma unroll
    for (int j = 0; j < kLookWindows; ++j) {
      const int idx = look - 32 * j - lane;
      d[j] = idx >= 0 ? ld_relaxed_u64(desc + idx) : kFlagPrefix;  // virtual tiles before 0: prefix 0
    }
#pragma unroll
    for (int j = 0; j < kLookWindows; ++j) {
      if (done) break;
      const int idx = look - 32 * j - lane;
      while (__any_sync(0xffffffffu, (d[j] >> 62) == 0)) {
        if ((d[j] >> 62) == 0) d[j] = ld_relaxed_u64(desc + idx);
      }
      const unsigned pm = __ballot_sync(0xffffffffu, (d[j] >> 62) == 2);
      const unsigned long long v = d[j] & kValueMask;
      if (pm) {
        const int first = __ffs(pm) - 1;  // nearest predecessor holding an inclusive prefix
        excl += warp_sum_u64(lane <= first ? v : 0ull);
        done = true;
      } else {
        excl += warp_sum_u64(v);
      }
    }
    look -= 32 * kLookWindows;
  }
  if (lane == 0) st_relaxed_u64(desc + tile, kFlagPrefix | (excl + agg));
  return excl;
}

// ---- the kernel ----------------------------------------------------------------------------------------
//
// CTA = one warp per mate (2 warps paired-end, 1 warp single-end); warp m owns mate m's records of the
// tile and output stream m (-o / -O).  The two warps meet once per tile (match exchange: the PE keep
// rule needs both mates' decisions, run.rs:149-160); everything else -- TMA load, parse, match, scan,
// look-back, emit, TMA store -- is private to the warp.  Tiles are claimed from a global counter, in
// order, so every predecessor a look-back waits on belongs to a running CTA that has nothing to wait
// for before publishing its aggregate.

constexpr int kSkipTile = -2147483647 - 1;

template <bool PAIRED>
__global__ void __launch_bounds__(PAIRED ? 64 : 32) extract_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint8_t cls_tab[2][256];
  __shared__ int s_ms[2][2][32];   // [iteration parity][mate][lane]
  __shared__ uint32_t s_next[2];   // [iteration parity]
  const int lane = threadIdx.x & 31;
  const int m = PAIRED ? (int)(threadIdx.x >> 5) : 0;  // this warp's mate = its output stream
  const bk_devprog& pg = P.prog[m];
  const bool trim = !(P.flags & 2u);
  const bool rc = (P.flags & 4u) != 0;
  const uint32_t R = P.tile_recs;

  uint8_t* const in_s = smem + (m ? P.in_cap[0] : 0);
  uint8_t* const out_s = smem + P.in_cap[0] + (PAIRED ? P.in_cap[1] : 0) + (m ? P.out_cap[0] : 0);
  const uint8_t* const text = P.text[m];
  const uint32_t* const off = P.off[m];
  uint8_t* const gdst = P.out[m];
  unsigned long long* const desc = P.desc[m];

  // byte -> "which classes contain it" table, so a class test is one shared-memory lookup
  for (int c = lane; c < 256; c += 32) {
    uint32_t mask = 0;
    if (c < 128)
      for (int k = 0; k < BK_MAX_CLASSES; ++k) mask |= ((pg.cls[k][c >> 5] >> (c & 31)) & 1u) << k;
    cls_tab[m][c] = (uint8_t)mask;
  }
  if (lane == 0) mbar_init(&bar[m], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (threadIdx.x == 0) s_next[1] = atomicAdd(P.tile_counter, 1u);
  __syncthreads();
  uint32_t tile = s_next[1];
  uint32_t parity = 0, it = 0;

  // record offsets of the next tile are fetched one tile ahead (registers only)
  uint32_t o_pre = 0, oe_pre = 0;
  if (tile < P.ntiles) {
    const uint32_t r0 = tile * R, nrec = min(R, P.n - r0);
    o_pre = lane < (int)nrec ? off[r0 + lane] : 0;
    oe_pre = off[r0 + nrec];
  }

  while (tile < P.ntiles) {
    const uint32_t r0 = tile * R;
    const uint32_t nrec = min(R, P.n - r0);
    const bool active = lane < (int)nrec;

    // ---- load: this mate's contiguous text range of the tile ------------------------------------
    const uint32_t o = o_pre, oe = oe_pre;
    uint32_t onext = __shfl_down_sync(0xffffffffu, o, 1);
    if (lane == (int)nrec - 1) onext = oe;
    const uint32_t gbase = __shfl_sync(0xffffffffu, o, 0) & ~15u;
    const uint32_t bytes = ((oe + 15u) & ~15u) - gbase;
    const bool too_big = bytes > P.in_cap[m];  // host sized the staging buffer from max_rec_bytes
    const uint32_t rb = o - gbase, re = onext - gbase;
    RecView rv;
    int ms = -1;
    if (!too_big) {
      if (lane == 0) {
        mbar_arrive_expect_tx(&bar[m], bytes);
        bulk_g2s(in_s, text + gbase, bytes, &bar[m]);
      }
      mbar_wait(&bar[m], parity);
      parity ^= 1;
      // ---- parse + match (lane per record) -----------------------------------------------------
      if (active) {
        rv = parse_record(in_s, rb, re);
        if (pg.present) {
          const uint8_t* sq = in_s + rv.seq;
          ms = match_read<false>(pg, cls_tab[m], sq, rv.seq_len);
          if (ms < 0 && rc) ms = match_read<true>(pg, cls_tab[m], sq, rv.seq_len);  // parse.rs:61-66
        }
        if (P.match[m]) P.match[m][r0 + lane] = ms;
      }
    } else if (lane == 0) {
      atomicOr(&P.result->error, 2u);
    }

    // ---- exchange with the other mate's warp ----------------------------------------------------------
    s_ms[it & 1][m][lane] = too_big ? kSkipTile : ms;
    __syncthreads();
    int ms_other = -1;
    bool skip = too_big;
    if (PAIRED) {
      ms_other = s_ms[it & 1][m ^ 1][lane];
      skip = skip || (s_ms[it & 1][m ^ 1][0] == kSkipTile);
    }
    // run.rs:71-78 (SE: unmatched reads vanish), run.rs:149-160 (PE: keep iff either mate matched)
    const bool keep = !skip && active && (ms >= 0 || ms_other >= 0);

    // ---- size + scan + look-back ------------------------------------------------------------------------
    const uint32_t olen = keep ? out_bytes(pg, rv, ms, trim) : 0;
    uint32_t incl = olen;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    const uint32_t ooff = incl - olen;
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned long long gout = lookback(desc, tile, total, lane);

    {  // statistics
      const unsigned kb = __ballot_sync(0xffffffffu, keep);
      const unsigned mb = __ballot_sync(0xffffffffu, active && ms >= 0);
      if (lane == 0) {
        if (m == 0 && kb) atomicAdd(&P.result->n_out, (unsigned long long)__popc(kb));
        if (mb) atomicAdd(&P.result->n_match[m], (unsigned long long)__popc(mb));
        if (tile == P.ntiles - 1) P.result->out_len[m] = gout + total;
      }
    }

    if (gout + total > P.cap[m]) {
      if (lane == 0) atomicOr(&P.result->error, 1u);
    } else if (total) {
      // ---- emit into the shared output tile (same 16-byte phase as the global destination) -----------
      const uint32_t phase = (uint32_t)(gout & 15ull);
      if (keep) emit_record(pg, in_s, rv, ms, trim, out_s, phase + ooff);
      fence_proxy_async_smem();
      __syncwarp();
      // ---- store: aligned middle by TMA, ragged edges by the lanes -------------------------------------
      const unsigned long long g0 = gout, g1 = gout + total;
      const unsigned long long a0 = (g0 + 15ull) & ~15ull, a1 = g1 & ~15ull;
      if (a0 >= a1) {
        for (unsigned long long x = g0 + lane; x < g1; x += 32) gdst[x] = out_s[phase + (uint32_t)(x - g0)];
      } else {
        if (lane == 0) {
          bulk_s2g(gdst + a0, out_s + phase + (uint32_t)(a0 - g0), (uint32_t)(a1 - a0));
          bulk_commit();
        }
        unsigned long long x = g0 + lane;
        if (x < a0) gdst[x] = out_s[phase + (uint32_t)(x - g0)];
        x = a1 + lane;
        if (x < g1) gdst[x] = out_s[phase + (uint32_t)(x - g0)];
        if (lane == 0) bulk_wait_read0();  // the tile may be overwritten once the bulk engine has read it
      }
      __syncwarp();
    }
    // Claim the next tile only now: a tile claimed early would sit unpublished while this warp
    // works (and waits in its own look-back), and every later tile's look-back would wait on it.
    if (threadIdx.x == 0) s_next[it & 1] = atomicAdd(P.tile_counter, 1u);
    __syncthreads();
    tile = s_next[it & 1];
    ++it;
    if (tile < P.ntiles) {
      const uint32_t q0 = tile * R, qn = min(R, P.n - q0);
      o_pre = lane < (int)qn ? off[q0 + lane] : 0;
      oe_pre = off[q0 + qn];
    }
  }
}

}  // namespace bk
