// The extract kernel: warp-specialised tile pipeline with ordered compaction (decoupled look-back).
// Building blocks (parse / match / emit / PTX wrappers) are in bk_kernels.cuh.
#pragma once
#include "bk_kernels.cuh"

namespace bk {

constexpr int kStages = 3;  // input ring depth per mate (front warp runs at most this far ahead)
constexpr int kBackWarps = kEmitParts;  // back warps per mate: warp j writes part j of every record (emit_part)

// ---- ordered compaction: decoupled look-back over tiles ------------------------------------------------

__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kLookWindows = 4;  // 32-descriptor windows fetched per round (independent loads in flight)

// Exclusive prefix of tile `tile` over all earlier tiles.  The tile's own aggregate has already been
// published (by the front warp); this publishes the inclusive prefix once it is known.
#ifdef BK_PROFILE_PHASES
__device__ unsigned long long g_look_rounds, g_look_spins, g_look_windows;
#endif
__device__ __noinline__ unsigned long long lookback(unsigned long long* desc, uint32_t tile,
                                                       unsigned long long agg, int lane) {
  unsigned long long excl = 0;
  if (tile != 0) {
    int look = (int)tile - 1;
    bool done = false;
    while (!done) {
#ifdef BK_PROFILE_PHASES
      if (lane == 0) atomicAdd(&g_look_rounds, 1ull);
#endif
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
#ifdef BK_PROFILE_PHASES
        if (lane == 0) atomicAdd(&g_look_windows, 1ull);
#endif
        while (__any_sync(0xffffffffu, (d[j] >> 62) == 0)) {
#ifdef BK_PROFILE_PHASES
          if (lane == 0) atomicAdd(&g_look_spins, 1ull);
#endif
          if ((d[j] >> 62) == 0) d[j] = ld_relaxed_u64(desc + idx);
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (d[j] >> 62) == 2);
        // aggregates are one tile's bytes (32 bits is plenty); only the prefix itself is 64-bit
        const unsigned long long v = d[j] & kValueMask;
        const int first = pm ? __ffs(pm) - 1 : 32;  // nearest predecessor holding an inclusive prefix
        excl += warp_sum_u32(lane < first ? (uint32_t)v : 0u);
        if (pm) {
          excl += __shfl_sync(0xffffffffu, v, first);
          done = true;
        }
      }
      look -= 32 * kLookWindows;
    }
  }
  if (lane == 0) st_relaxed_u64(desc + tile, kFlagPrefix | (excl + agg));
  return excl;
}

// ---- the kernel ----------------------------------------------------------------------------------------
//
// Per mate m (one for single-end, two for paired-end) a FRONT warp and a BACK warp share a ring of
// kStages staged input tiles:
//
//   front F_m : wait ring slot free -> claim tile (global counter, in order) -> TMA load -> parse ->
//               match -> exchange match decisions with the other mate's front warp (PE keep rule,
//               run.rs:149-160) -> sizes + warp scan -> PUBLISH the tile aggregate -> hand the slot on
//   scan  S_m : take slot -> look-back over earlier tiles -> publish the inclusive prefix -> pass the
//               tile's global output offset to the back warp
//   back  B_m : (kBackWarps warps) take slot -> each warp emits its word-aligned share of every
//               record into the shared output tile -> release slot -> one of them issues the TMA store
//
// The front warp never waits on another tile's result, so a claimed tile's aggregate is published
// after a short, bounded delay; that is what keeps every other tile's look-back from stalling (a
// warp that both publishes and waits delays its own next aggregate by its wait, and the delay
// compounds down the chain).  The scan warp resolves prefixes as soon as aggregates exist, so
// look-backs stay short, and its spinning costs nobody's critical path.  Lane i of a front/back
// warp owns record i of the tile.

// Optional phase timers (compile with -DBK_PROFILE_PHASES): per-role cycle sums, lane 0 only.
#ifdef BK_PROFILE_PHASES
#define PROF_INIT long long p_t = clock64(); long long p_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PROF(i) { const long long p_n = clock64(); p_acc[i] += p_n - p_t; p_t = p_n; }
#define PROF_FLUSH(base) \
  if (lane == 0 && m == 0) for (int p_i = 0; p_i < 8; ++p_i) atomicAdd(&P.prof[(base) + p_i], (unsigned long long)p_acc[p_i]);
#else
#define PROF_INIT
#define PROF(i)
#define PROF_FLUSH(base)
#endif

// named barrier of one mate's back warps (ids 2 and 3; 0 is __syncthreads, 1 the front warps')
__device__ __forceinline__ void back_sync(int m) {
  if (m == 0)
    asm volatile("bar.sync 2, %0;" ::"n"(kBackWarps * 32) : "memory");
  else
    asm volatile("bar.sync 3, %0;" ::"n"(kBackWarps * 32) : "memory");
}

constexpr int kSkipTile = -2147483647 - 1;
constexpr uint32_t kEndTile = 0xFFFFFFFFu;
constexpr uint32_t kNotKept = 0xFFFFFFFFu;
constexpr uint32_t kNoPrefetch = 0xFFFFFFFEu;

struct TileMeta {  // front -> back hand-over, one per (mate, stage); SoA so lanes do not conflict
  uint32_t beg[32], end[32], head_len[32], seq[32], seq_len[32], qual[32], qual_len_fast[32];
  int32_t ms[32];
  uint32_t ooff[32];  // record's offset inside the tile's output, kNotKept if dropped
  uint32_t olen[32];  // record's output bytes
  uint32_t tile;      // kEndTile terminates the back warp
  uint32_t total;     // output bytes of the tile on this stream
  unsigned long long gout;  // global output offset of the tile on this stream
};

template <bool PAIRED>
__global__ void __launch_bounds__((PAIRED ? 2 : 1) * (2 + kBackWarps) * 32, 2) extract_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NM = PAIRED ? 2 : 1;
  __shared__ __align__(8) uint64_t bar_full[2][kStages], bar_pub[2][kStages], bar_ready[2][kStages],
      bar_empty[2][kStages];
  __shared__ uint8_t cls_tab[2][kTabBytes];  // per mate: class bits per byte, then the complement table
  __shared__ int s_ms[2][2][32];  // [tile parity][mate][lane]
  __shared__ uint32_t s_tile[2];  // [tile parity]
  __shared__ uint32_t s_pref[2];  // [tile parity] tile claimed ahead by front warp 0, or kNoPrefetch
  __shared__ TileMeta meta[NM][kStages];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // Warp roles, lowest warp ids first: back, scan, front.  The scheduler favours higher warp ids, and
  // the front warps' claim -> publish latency is what every other tile's look-back waits on.
  constexpr int NB = NM * kBackWarps;
  const bool is_front = warp >= NB + NM;
  const bool is_scan = !is_front && warp >= NB;
  const int m = warp % NM;  // mate = output stream
  const int bj = warp < NB ? warp / NM : 0;  // back warp index within the mate
  const bk_devprog& pg = P.prog[m];
  const bool trim = !(P.flags & 2u);
  const bool rc = (P.flags & 4u) != 0;
  const uint32_t R = P.tile_recs;

  // dynamic shared memory: [mate0 ring][mate1 ring][out0][out1]
  // (32-bit shared-space addresses; see the note at the top of bk_kernels.cuh)
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t in_ring = smem_base + (m ? kStages * P.in_cap[0] : 0);
  const uint32_t in_stride = P.in_cap[m];
  const uint32_t out_s = smem_base + kStages * (P.in_cap[0] + (PAIRED ? P.in_cap[1] : 0)) + (m ? P.out_cap[0] : 0);
  const uint32_t tab = smem_u32(&cls_tab[m][0]);

  if (is_front) {
    // byte -> "which classes contain it" table, so a class test is one shared-memory lookup
    for (int c = lane; c < 256; c += 32) {
      uint32_t mask = 0;
      if (c < 128)
        for (int k = 0; k < BK_MAX_CLASSES; ++k) mask |= ((pg.cls[k][c >> 5] >> (c & 31)) & 1u) << k;
      cls_tab[m][c] = (uint8_t)mask;
      cls_tab[m][kTabComp + c] = (uint8_t)complement((uint32_t)c);
    }
    if (lane == 0)
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&bar_full[m][s], 1);
        mbar_init(&bar_pub[m][s], 1);
        mbar_init(&bar_ready[m][s], 1);
        mbar_init(&bar_empty[m][s], kBackWarps);
      }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  if (is_front) {
    // =============================== FRONT ======================================================
    const uint8_t* const text = P.text[m];
    const uint32_t* const off = P.off[m];
    // Start the load of `tile` into ring slot `slot`: record offsets, then one 1-D TMA copy of the
    // tile's contiguous text range.  Returns this lane's record bounds inside the staged tile.
    struct Staged { uint32_t rb, re, nrec; bool too_big; };
    auto stage_tile = [&](uint32_t tile, uint32_t slot) -> Staged {
      Staged st;
      const uint32_t r0 = tile * R;
      st.nrec = min(R, P.n - r0);
      const uint32_t o = lane < (int)st.nrec ? off[r0 + lane] : 0;
      const uint32_t oe = off[r0 + st.nrec];
      uint32_t onext = __shfl_down_sync(0xffffffffu, o, 1);
      if (lane == (int)st.nrec - 1) onext = oe;
      const uint32_t gbase = __shfl_sync(0xffffffffu, o, 0) & ~15u;
      const uint32_t bytes = ((oe + 15u) & ~15u) - gbase;
      st.too_big = bytes > in_stride;  // host sized the ring from max_rec_bytes
      st.rb = o - gbase;
      st.re = onext - gbase;
      if (!st.too_big && lane == 0) {
        mbar_arrive_expect_tx(&bar_full[m][slot], bytes);
        bulk_g2s(in_ring + slot * in_stride, text + gbase, bytes, &bar_full[m][slot]);
      }
      return st;
    };
    PROF_INIT
    bool have = false;  // tile of this iteration already claimed and its load in flight
    uint32_t tile = 0;
    Staged cur = {0, 0, 0, false};
    for (uint32_t k = 0;; ++k) {
      const uint32_t s = k % kStages, ph = (k / kStages) & 1u;
      TileMeta& tm = meta[m][s];
      if (!have) {
        // slot free -- on BOTH mates before anyone claims: a tile claimed while the other mate's front
        // warp still waits for its slot would sit unpublished behind that wait
        mbar_wait(&bar_empty[m][s], ph ^ 1u);
        if (PAIRED) mbar_wait(&bar_empty[m ^ 1][s], ph ^ 1u);
        PROF(0)
        // claim with the slot in hand: the aggregate follows within one load + parse
        if (PAIRED) {
          if (m == 0 && lane == 0) s_tile[k & 1] = atomicAdd(P.tile_counter, 1u);
          asm volatile("bar.sync 1, 64;" ::: "memory");
          tile = s_tile[k & 1];
        } else {
          tile = 0;
          if (lane == 0) tile = atomicAdd(P.tile_counter, 1u);
          tile = __shfl_sync(0xffffffffu, tile, 0);
        }
        PROF(1)
        if (tile < P.ntiles) cur = stage_tile(tile, s);
        PROF(2)
      }
      if (tile >= P.ntiles) {
        if (lane == 0) {
          tm.tile = kEndTile;
          mbar_arrive(&bar_pub[m][s]);
        }
        break;
      }
      // ---- prefetch: if the next ring slot is already free (both mates), claim tile k+1 now and
      // put its load in flight behind this tile's parse + match.  Holding a claimed tile through a
      // bounded amount of own work is fine; what must never happen is waiting on OTHER tiles with
      // an unpublished claim in hand, hence "only if the slot is free right now".
      const uint32_t s1 = (k + 1) % kStages, ph1 = ((k + 1) / kStages) & 1u;
      uint32_t nxt = kNoPrefetch;
      if (PAIRED) {
        if (m == 0 && lane == 0) {
          const bool ok = mbar_test(&bar_empty[0][s1], ph1 ^ 1u) && mbar_test(&bar_empty[1][s1], ph1 ^ 1u);
          s_pref[(k + 1) & 1] = ok ? atomicAdd(P.tile_counter, 1u) : kNoPrefetch;
        }
        asm volatile("bar.sync 1, 64;" ::: "memory");
        nxt = s_pref[(k + 1) & 1];
      } else {
        if (lane == 0 && mbar_test(&bar_empty[0][s1], ph1 ^ 1u)) nxt = atomicAdd(P.tile_counter, 1u);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
      }
      Staged nst = {0, 0, 0, false};
      if (nxt != kNoPrefetch && nxt < P.ntiles) {
        mbar_wait(&bar_empty[m][s1], ph1 ^ 1u);  // already complete; acquires the slot for this warp
        nst = stage_tile(nxt, s1);
      }

      const uint32_t r0 = tile * R;
      const uint32_t nrec = cur.nrec;
      const bool active = lane < (int)nrec;
      const bool too_big = cur.too_big;
      const uint32_t in_s = in_ring + s * in_stride;
      RecView rv;
      rv.beg = rv.end = rv.fast = rv.head_len = rv.seq = rv.seq_len = rv.qual = rv.qual_len = 0;
      int ms = -1;
      if (!too_big) {
        mbar_wait(&bar_full[m][s], ph);
        PROF(3)
        // ---- parse + match (lane per record) ---------------------------------------------------
        if (active) {
          rv = parse_record(in_s + cur.rb, in_s + cur.re);
          if (pg.present && rv.qual_len == rv.seq_len) {  // unequal lengths never leave bk_tokenize
            ms = match_read<false>(pg, tab, rv.seq, rv.seq_len);
            if (ms < 0 && rc) ms = match_read<true>(pg, tab, rv.seq, rv.seq_len);  // parse.rs:61-66
          }
          if (P.match[m]) P.match[m][r0 + lane] = ms;
        }
      } else if (lane == 0) {
        atomicOr(&P.result->error, 2u);
      }

      PROF(4)
      // ---- exchange match decisions with the other mate's front warp -----------------------------
      int ms_other = -1;
      bool skip = too_big;
      if (PAIRED) {
        s_ms[k & 1][m][lane] = too_big ? kSkipTile : ms;
        asm volatile("bar.sync 1, 64;" ::: "memory");
        ms_other = s_ms[k & 1][m ^ 1][lane];
        skip = skip || (s_ms[k & 1][m ^ 1][0] == kSkipTile);
      }
      PROF(5)
      // run.rs:71-78 (SE: unmatched reads vanish), run.rs:149-160 (PE: keep iff either mate matched)
      const bool keep = !skip && active && (ms >= 0 || ms_other >= 0);

      // ---- size + scan, publish the aggregate -------------------------------------------------------
      const uint32_t olen = keep ? out_bytes(pg, rv, ms, trim) : 0;
      uint32_t incl = olen;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      if (lane == 0) st_relaxed_u64(P.desc[m] + tile, kFlagAgg | (unsigned long long)total);

      {  // statistics
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        const unsigned mb = __ballot_sync(0xffffffffu, active && ms >= 0);
        if (lane == 0) {
          if (m == 0 && kb) atomicAdd(&P.result->n_out, (unsigned long long)__popc(kb));
          if (mb) atomicAdd(&P.result->n_match[m], (unsigned long long)__popc(mb));
        }
      }

      // ---- hand the slot to the back warp -------------------------------------------------------------
      tm.beg[lane] = rv.beg;
      tm.end[lane] = rv.end;
      tm.head_len[lane] = rv.head_len;
      tm.seq[lane] = rv.seq;
      tm.seq_len[lane] = rv.seq_len;
      tm.qual[lane] = rv.qual;
      tm.qual_len_fast[lane] = (rv.qual_len << 1) | rv.fast;
      tm.ms[lane] = ms;
      tm.ooff[lane] = keep ? incl - olen : kNotKept;
      tm.olen[lane] = olen;
      if (lane == 0) {
        tm.tile = tile;
        tm.total = total;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_pub[m][s]);
      PROF(6)
      have = nxt != kNoPrefetch;
      if (have) {
        tile = nxt;
        cur = nst;
      }
    }
    PROF_FLUSH(0)
  } else if (is_scan) {
    // =============================== SCAN =======================================================
    PROF_INIT
    for (uint32_t k = 0;; ++k) {
      const uint32_t s = k % kStages, ph = (k / kStages) & 1u;
      TileMeta& tm = meta[m][s];
      mbar_wait(&bar_pub[m][s], ph);
      PROF(0)
      const uint32_t tile = tm.tile;
      if (tile != kEndTile) {
        const unsigned long long gout = lookback(P.desc[m], tile, tm.total, lane);
        if (lane == 0) tm.gout = gout;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_ready[m][s]);
      PROF(1)
      if (tile == kEndTile) break;
    }
    PROF_FLUSH(8)
  } else {
    // =============================== BACK =======================================================
    uint8_t* const gdst = P.out[m];
    PROF_INIT
    for (uint32_t k = 0;; ++k) {
      const uint32_t s = k % kStages, ph = (k / kStages) & 1u;
      const TileMeta& tm = meta[m][s];
      mbar_wait(&bar_ready[m][s], ph);
      PROF(0)
      const uint32_t tile = tm.tile;
      if (tile == kEndTile) break;
      const uint32_t total = tm.total;
      const unsigned long long gout = tm.gout;
      if (lane == 0 && tile == P.ntiles - 1) P.result->out_len[m] = gout + total;
      const bool fits = gout + total <= P.cap[m];
      if (!fits && lane == 0) atomicOr(&P.result->error, 1u);
      const uint32_t phase = (uint32_t)(gout & 15ull);
      if (fits && total) {
        // ---- emit into the shared output tile (same 16-byte phase as the global destination) --------
        const uint32_t ooff = tm.ooff[lane];
        if (ooff != kNotKept) {
          RecView rv;
          rv.beg = tm.beg[lane];
          rv.end = tm.end[lane];
          rv.head_len = tm.head_len[lane];
          rv.seq = tm.seq[lane];
          rv.seq_len = tm.seq_len[lane];
          rv.qual = tm.qual[lane];
          const uint32_t qf = tm.qual_len_fast[lane];
          rv.qual_len = qf >> 1;
          rv.fast = qf & 1u;
          emit_part(pg, rv, tm.ms[lane], trim, out_s + phase + ooff, bj);
        }
      }
      __syncwarp();
      PROF(1)
      if (lane == 0) mbar_arrive(&bar_empty[m][s]);  // ring slot (input tile + meta) is free again
      if (fits && total) fence_proxy_async_smem();
      // all shares written -> one warp stores; nobody overwrites the tile until the store has read it
      back_sync(m);
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
          if (lane == 0) bulk_wait_read0();  // the output tile may be overwritten once the engine has read it
        }
        __syncwarp();
      }
      back_sync(m);
      PROF(2)
    }
    if (bj == 0) { PROF_FLUSH(16) }
  }
}

}  // namespace bk
