// The extract kernel for patterns anchored at the read start ("scout" pipeline).
//
// extract_kernel (bk_extract.cuh) keeps a staged input tile in shared memory for the whole of
// claim -> load -> parse -> match -> publish -> look-back -> assemble -> store: ~23 k cycles per ring slot,
// of which only the last 4-5 k need the tile's text on chip, and 227 KB of shared memory hold eight
// such slots per SM.  Here the sizing half of the pipeline runs AHEAD of the data ring, on L2:
//
//   scout  S_m : wait META slot free -> claim tile -> record offsets -> each lane copies the first
//                144 bytes of its record (cp.async, 16-byte chunks) into a private window -> parse +
//                match there (`^` patterns touch nothing else; the four framing probes behind the
//                window are plain global byte loads) -> exchange with the other mate's scout -> sizes,
//                warp scan -> PUBLISH the tile aggregate -> meta slot (0.4 KB per mate and tile)
//   scan       : look-back over earlier tiles -> the tile's global output offsets -> meta slot
//   loader     : wait DATA slot free + tile sized -> one 1-D TMA load per mate (an L2 hit: the scout's
//                claim prefetched the text)
//   back   B_m : wait data landed + offsets known -> assemble (emit_part / bulk16, unchanged) -> store
//
// A data slot is now held for load + assemble only; the meta ring (kMeta tiles per CTA) absorbs the
// latency of sizing and of the look-back, which no longer costs shared memory worth mentioning.
// Records the window cannot serve (header line longer than the window, CR line ends, text after
// '+', truncated last record) take scout_slow: the same parse and match bytewise from global memory.
//
// Reference semantics are those of bk_extract.cuh (run.rs:63-80, :149-195; parse.rs:58-171;
// fastq.rs:259-263); only the schedule differs.
#pragma once
#include "bk_extract.cuh"

namespace bk {
namespace v5 {

#ifndef BK5_DATA
#define BK5_DATA 2
#endif
#ifndef BK5_PAIRS
#define BK5_PAIRS 3
#endif
constexpr int kData = BK5_DATA;   // staged input tiles per mate (load + assemble in flight)
constexpr int kMeta = 8;   // tiles sized ahead per CTA
constexpr int kWinChunks = 9;
constexpr uint32_t kWinStride = 16u * kWinChunks;  // 144 B: lanes' 16-byte chunks fall into distinct bank groups
constexpr uint32_t kMaxPatLen = 64;                // longer `^` patterns stay on extract_kernel (host check)

struct Meta {  // one per (mate, meta slot)
  uint32_t goff[32];  // record start, offset into the mate's text
  uint32_t hl[32];    // head_len | fast << 30 | matched << 31
  uint32_t ooff[32];  // offset inside the tile's output, kNotKept if dropped
  uint32_t tile, nrec, too_big, total;
  uint32_t gbase, bytes, gend, pad;
  unsigned long long gout;
};

__device__ __forceinline__ uint32_t ldg8(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ uint32_t lds_volatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// one copy of the polling loop (the kernel's hot code has to fit the instruction cache)
__device__ __noinline__ void mbar_wait_f(uint32_t bar_saddr, uint32_t parity) {
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar_saddr), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(128);
  }
}
__device__ __forceinline__ void mbar_wait5(uint64_t* bar, uint32_t parity) { mbar_wait_f(smem_u32(bar), parity); }

struct Scouted {
  uint32_t head_len, seq_len, qual_len, fast;
  int ms;
};

// parse_record + match_read<false> for one record, bytewise from global memory (rare records only).
// `tab` = the mate's program block in shared memory.  Patterns here are anchored at the read start.
__device__ __noinline__ Scouted scout_slow(const uint8_t* T, uint32_t b, uint32_t e, uint32_t tab, bool present) {
  Scouted r;
  auto find = [&](uint32_t from) {
    uint32_t p = from;
    while (p < e && T[p] != '\n') ++p;
    return p;
  };
  auto strip = [&](uint32_t lb, uint32_t le) { return (le > lb && T[le - 1] == '\r') ? le - 1 : le; };
  const uint32_t nl1 = find(b);
  const uint32_t sb = nl1 + 1 < e ? nl1 + 1 : e;
  const uint32_t rem = e - sb, L = (rem - 4u) >> 1;
  bool done = false;
  if (rem >= 4u && !(rem & 1u) && nl1 > b) {  // the writer's own framing (parse_record's closed form)
    const bool framed = T[sb + L] == '\n' && T[sb + L + 1] == '+' && T[sb + L + 2] == '\n' && T[e - 1] == '\n';
    const bool cr = T[sb + L - 1] == '\r' || T[e - 2] == '\r';
    if (framed && (L == 0 || !cr)) {
      const bool hcr = nl1 > b + 1u && T[nl1 - 1] == '\r';
      r.head_len = nl1 - (b + 1u) - (hcr ? 1u : 0u);
      r.seq_len = r.qual_len = L;
      r.fast = !hcr;
      done = true;
    }
  }
  if (!done) {
    r.fast = 0;
    r.head_len = strip(b + 1, nl1 > b ? nl1 : b + 1) - (b + 1);
    const uint32_t nl2 = find(sb);
    r.seq_len = strip(sb, nl2) - sb;
    const uint32_t pb = nl2 + 1 < e ? nl2 + 1 : e;
    const uint32_t nl3 = find(pb);
    const uint32_t qb = nl3 + 1 < e ? nl3 + 1 : e;
    const uint32_t nl4 = find(qb);
    r.qual_len = strip(qb, nl4) - qb;
  }
  r.ms = -1;
  if (present && r.qual_len == r.seq_len) {
    const uint4 h = lds128(tab + kTabHdr);  // nops, total_len, anchor_start (set), anchor_end
    bool ok = r.seq_len >= h.y && !(h.w && r.seq_len != h.y);
    for (uint32_t o = 0; ok && o < h.x; ++o) {
      const MOp op = tab_op(tab, o);
      const uint8_t* p = T + sb + op.off;
      if (op.kind == BK_OP_LIT) {
        uint32_t mism = 0, bad = 0;
        for (uint32_t i = 0; i < op.len; ++i) {
          const uint32_t c = p[i];
          const uint32_t ne = (c != lds8(tab + kTabLit + op.idx + i));
          mism += ne;
          bad |= ne & (uint32_t)((c >= 0x80u) | (c == '\n'));  // a wildcard slot is the regex `.` (pattern.rs:10,71)
        }
        ok = (mism <= op.k) && !bad;
      } else {
        uint32_t acc = 0xFFu;
        for (uint32_t i = 0; i < op.len; ++i) acc &= lds8(tab + p[i]);
        ok = (acc >> op.idx) & 1u;
      }
    }
    if (ok) r.ms = 0;
  }
  return r;
}

constexpr int kFrontPairs = BK5_PAIRS, kScanWarps = 2, kBack = kEmitParts;
constexpr int cta_warps(bool paired) { return (paired ? 2 : 1) * (kFrontPairs + kBack) + kScanWarps + 1; }
constexpr uint32_t kEnd = 0xFFFFFFFFu;

template <bool PAIRED>
__global__ void __launch_bounds__(cta_warps(PAIRED) * 32, 2) extract_kernel5(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int NM = PAIRED ? 2 : 1;
  __shared__ __align__(8) uint64_t bar_full[2][kData], bar_empty[2][kData];
  __shared__ __align__(8) uint64_t meta_empty[2][kMeta], bar_pub[kMeta], bar_ready[kMeta];
  __shared__ __align__(16) uint8_t cls_tab[2][kTabBytes];
  __shared__ int s_ms[kFrontPairs][2][2][32];  // [front pair][round parity][mate][lane]
  __shared__ uint32_t s_claim[kFrontPairs];
  __shared__ uint32_t s_claimed;  // tiles claimed by this CTA so far: claims are made in ring order
  __shared__ Meta meta[NM][kMeta];
  __shared__ uint4 run_desc[NM][32];
  __shared__ __align__(16) uint8_t win[NM * kFrontPairs][32 * kWinStride + 16];

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int NB = NM * kBack;
  const bool is_back = warp < NB;
  const bool is_scan = !is_back && warp < NB + kScanWarps;
  const bool is_front = !is_back && !is_scan && warp < NB + kScanWarps + NM * kFrontPairs;
  const int fw = warp - NB - kScanWarps;
  const uint32_t nb0 = !PAIRED ? kBack
                       : (P.prog[0].present && !P.prog[1].present) ? kBack + 1
                       : (!P.prog[0].present && P.prog[1].present) ? kBack - 1 : kBack;
  const int m = is_back ? (warp < (int)nb0 ? 0 : 1) : is_front ? fw % NM : 0;
  const int fp = is_front ? fw / NM : 0;
  const int bj = is_back ? (m ? warp - (int)nb0 : warp) : 0;
  const uint32_t nparts = m ? NB - nb0 : nb0;
  const bk_devprog& pg = P.prog[m];
  const bool trim = !(P.flags & 2u);
  const uint32_t R = P.tile_recs;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t out_s = smem_base + kData * (P.in_cap[0] + (PAIRED ? P.in_cap[1] : 0)) + (m ? P.out_cap[0] : 0);
  const uint32_t tab = smem_u32(&cls_tab[m][0]);

  if (is_front && fp == 0) {
    for (int c = lane; c < 256; c += 32) {
      uint32_t mask = 0;
      if (c < 128)
        for (int k = 0; k < BK_MAX_CLASSES; ++k) mask |= ((pg.cls[k][c >> 5] >> (c & 31)) & 1u) << k;
      cls_tab[m][c] = (uint8_t)mask;
      cls_tab[m][kTabComp + c] = (uint8_t)complement((uint32_t)c);
    }
    uint32_t* const tw = reinterpret_cast<uint32_t*>(&cls_tab[m][0]);
    if (lane == 0) {
      tw[kTabHdr / 4 + 0] = pg.nops;
      tw[kTabHdr / 4 + 1] = pg.total_len;
      tw[kTabHdr / 4 + 2] = pg.anchor_start;
      tw[kTabHdr / 4 + 3] = pg.anchor_end;
    }
    for (int i = lane; i < 2 * BK_MAX_OPS; i += 32) tw[kTabOps / 4 + i] = reinterpret_cast<const uint32_t*>(pg.ops)[i];
    for (int i = lane; i < BK_MAX_LIT / 4; i += 32) tw[kTabLit / 4 + i] = reinterpret_cast<const uint32_t*>(pg.lit)[i];
    if (lane == 0) {
      for (int s = 0; s < kData; ++s) {
        mbar_init(&bar_full[m][s], 1);
        mbar_init(&bar_empty[m][s], nparts);
      }
      for (int j = 0; j < kMeta; ++j) {
        mbar_init(&meta_empty[m][j], nparts);
        if (m == 0) {
          mbar_init(&bar_pub[j], NM);
          mbar_init(&bar_ready[j], 1);
        }
      }
      if (m == 0) s_claimed = 0;
    }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  if (!is_back && !is_scan && !is_front) {
    // =============================== LOADER =====================================================
    for (uint32_t k = 0;; ++k) {
      const uint32_t s = k % kData, j = k % kMeta;
      mbar_wait5(&bar_pub[j], (k / kMeta) & 1u);  // tile claimed and sized
#pragma unroll
      for (int mm = 0; mm < NM; ++mm) mbar_wait5(&bar_empty[mm][s], ((k / kData) & 1u) ^ 1u);
      const uint32_t tile = meta[0][j].tile;
      if (lane == 0) {
#pragma unroll
        for (int mm = 0; mm < NM; ++mm) {
          const Meta& tm = meta[mm][j];
          if (tile == kEnd || tm.too_big) {
            mbar_arrive(&bar_full[mm][s]);
          } else {
            const uint32_t in_s = smem_base + (mm ? kData * P.in_cap[0] : 0) + s * P.in_cap[mm];
            mbar_arrive_expect_tx(&bar_full[mm][s], tm.bytes);
            bulk_g2s(in_s, P.text[mm] + tm.gbase, tm.bytes, &bar_full[mm][s]);
          }
        }
      }
      if (tile != kEnd) { EVT(tile, 15) }
      if (tile == kEnd) break;
    }
  } else if (is_front) {
    // =============================== SCOUT ======================================================
    uint32_t n_kept = 0, n_matched = 0;
    const uint32_t pair_bar = fp == 0 ? 1u : 3u + (uint32_t)fp;
    const uint32_t win_s = smem_u32(&win[fw][0]) + (uint32_t)lane * kWinStride;
    const uint8_t* const T = P.text[m];
    const unsigned long long tl16 = ((unsigned long long)P.text_len[m] + 15ull) & ~15ull;
    for (uint32_t k = fp;; k += kFrontPairs) {
      const uint32_t j = k % kMeta, phj = (k / kMeta) & 1u;
      const uint32_t xb = (k / kFrontPairs) & 1u;
      Meta& tm = meta[m][j];
      uint32_t tile = 0;
      if (m == 0) {
        // the meta slot is free on every mate; claims of one CTA are made in ring order, so that its
        // tiles carry ascending numbers and the two end markers sit in consecutive slots
#pragma unroll
        for (int mm = 0; mm < NM; ++mm) mbar_wait5(&meta_empty[mm][j], phj ^ 1u);
        if (lane == 0) {
          while (lds_volatile(&s_claimed) != k) __nanosleep(20);
#ifdef BK_PROFILE_PHASES
          const long long t_free = clock64();
#endif
          tile = atomicAdd(P.tile_counter, 1u);
          sts_volatile(&s_claimed, k + 1u);
          if (PAIRED) s_claim[fp] = tile;
#ifdef BK_PROFILE_PHASES
          if (tile < P.ntiles) P.prof[(size_t)tile * 16 + 0] = (unsigned long long)t_free;
#endif
        }
        tile = __shfl_sync(0xffffffffu, tile, 0);
      }
      if (PAIRED) {
        nbar_sync(pair_bar, 64);
        if (m == 1) tile = s_claim[fp];
      }
      if (tile >= P.ntiles) {
        if (lane == 0) {
          tm.tile = kEnd;
          mbar_arrive(&bar_pub[j]);
        }
        break;
      }
      if (m == 0) { EVT(tile, 1) }
      const uint32_t r0 = tile * R;
      const uint32_t nrec = min(R, P.n - r0);
      const bool active = lane < (int)nrec;
      const uint32_t o = active ? P.off[m][r0 + lane] : 0u;
      const uint32_t oe = P.off[m][r0 + nrec];
      uint32_t onext = __shfl_down_sync(0xffffffffu, o, 1);
      if (lane == (int)nrec - 1) onext = oe;
      const uint32_t o0 = __shfl_sync(0xffffffffu, o, 0);
      const uint32_t gbase = o0 & ~15u;
      const uint32_t bytes = ((oe + 15u) & ~15u) - gbase;
      const bool too_big = bytes > P.in_cap[m];
      if (m == 0) { EVT(tile, 2) }
      // ---- the record's first bytes -> window; the two probes at its end ------------------------------
      const uint32_t b = o, e = onext;
      const bool live = active && !too_big;
      uint32_t c3 = 0, c5 = 0;
      if (live) {
        const uint32_t gb = b & ~15u;
        const unsigned long long left = (tl16 - gb) >> 4;  // 16-byte chunks of text from gb on
        const uint32_t nch = left < (unsigned long long)kWinChunks ? (uint32_t)left : (uint32_t)kWinChunks;
#pragma unroll 1
        for (uint32_t c = 0; c < nch; ++c) cp_async16(win_s + 16u * c, T + gb + 16u * c);
        if (e - b >= 2u) {
          c3 = ldg8(T + e - 1);
          c5 = ldg8(T + e - 2);
        }
      }
      if (lane == 0) {
        // Warm L2 for the tile claimed one grid further on (see extract_kernel's loader): its offsets and,
        // by extrapolation, its text -- which the scout's windows and, later, the TMA load both hit.
        const uint32_t nt = tile + gridDim.x;
        if (nt < P.ntiles) {
          bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(P.off[m]) + (((size_t)nt * R * 4u) & ~size_t(15)), ((4u * R + 15u) & ~15u) + 16u);
          const unsigned long long est = (unsigned long long)gbase + (unsigned long long)gridDim.x * (oe - o0);
          if (est + bytes + 256u <= P.text_len[m]) bulk_prefetch_l2(T + (est & ~15ull), bytes + 256u);
        }
      }
      cp_async_wait_all();
      __syncwarp();
      if (m == 0) { EVT(tile, 3) }
      // ---- parse + match inside the window -----------------------------------------------------------
      Scouted sc;
      sc.head_len = sc.seq_len = sc.qual_len = sc.fast = 0;
      sc.ms = -1;
      bool slow = false;
      if (live) {
        const uint32_t ws = win_s + (b & 15u), wlim = win_s + kWinStride;
        const uint32_t reclen = e - b;
        const uint32_t wend = ws + (reclen < wlim - ws ? reclen : wlim - ws);
        const uint32_t nl1 = find_nl_head(ws, wend);
        const uint32_t sbw = nl1 + 1u;
        const uint32_t sb = b + (sbw - ws);
        const uint32_t rem = e - sb, L = (rem - 4u) >> 1;
        const bool shape = nl1 < wend && nl1 > ws && rem >= 4u && !(rem & 1u) && sb <= e;
        slow = !shape;
        if (shape) {
          const uint32_t c0 = ldg8(T + sb + L), c1 = ldg8(T + sb + L + 1u), c2 = ldg8(T + sb + L + 2u);
          const uint32_t c4 = ldg8(T + sb + L - 1u), c6 = lds8(nl1 - 1u);
          // (the probes are in flight while the pattern is matched)
          if (pg.present) {
            if (sbw + lds32(tab + kTabHdr + 4u) <= wlim) sc.ms = match_read<false>(tab, sbw, L);
            else slow = true;
          }
          const bool framed = (c0 == '\n') & (c1 == '+') & (c2 == '\n') & (c3 == '\n');
          const bool cr = (c4 == '\r') | (c5 == '\r');
          if (framed && (L == 0 || !cr)) {
            const bool hcr = nl1 > ws + 1u && c6 == '\r';
            sc.head_len = nl1 - (ws + 1u) - (hcr ? 1u : 0u);
            sc.seq_len = sc.qual_len = L;
            sc.fast = !hcr;
          } else {
            slow = true;
          }
        }
      }
      if (__any_sync(0xffffffffu, slow)) {
        if (slow) sc = scout_slow(T, b, e, tab, pg.present != 0);
      }
      __syncwarp();
      if (too_big && lane == 0) atomicOr(&P.result->error, 2u);
      const int ms = sc.ms;
      if (m == 0) { EVT(tile, 5) }
      // ---- exchange match decisions with the other mate's scout ----------------------------------------
      int ms_other = -1;
      bool skip = too_big;
      if (PAIRED) {
        s_ms[fp][xb][m][lane] = too_big ? kSkipTile : ms;
        nbar_sync(pair_bar, 64);
        ms_other = s_ms[fp][xb][m ^ 1][lane];
        skip = skip || (s_ms[fp][xb][m ^ 1][0] == kSkipTile);
      }
      if (live && P.match[m]) P.match[m][r0 + lane] = ms;
      const bool keep = !skip && active && (ms >= 0 || ms_other >= 0);  // run.rs:71-78, :149-160

      // ---- size + scan, publish the aggregate -------------------------------------------------------
      RecView rv;
      rv.head_len = sc.head_len;
      rv.seq_len = sc.seq_len;
      rv.qual_len = sc.qual_len;
      const uint32_t olen = keep ? out_bytes(pg, rv, ms, trim) : 0;
      uint32_t incl = olen;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long blk_old = 0;
      if (lane == 0) {
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(P.desc[0] + tile) + m),
                     "r"(total | kTileReady)
                     : "memory");
        blk_old = atomicAdd(P.desc[2] + 2 * (size_t)(tile >> 5) + m, (1ull << 48) | (unsigned long long)total);
      }
      EVT(tile, m == 0 ? 6 : 14)
      n_kept += keep;
      n_matched += active && ms >= 0;

      // ---- meta slot: all the loader, the scan and the back warps need ----------------------------------
      tm.goff[lane] = o;
      tm.hl[lane] = sc.head_len | (sc.fast << 30) | (ms >= 0 ? 0x80000000u : 0u);
      tm.ooff[lane] = keep ? incl - olen : kNotKept;
      if (lane == 0) {
        tm.tile = tile;
        tm.nrec = nrec;
        tm.too_big = too_big;
        tm.total = total;
        tm.gbase = gbase;
        tm.bytes = bytes;
        tm.gend = oe;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_pub[j]);
      if (m == 0) { EVT(tile, 7) }
      if (lane == 0) {
        const uint32_t blk = tile >> 5;
        const uint32_t in_blk = min(32u, P.ntiles - (blk << 5));
        if ((uint32_t)(blk_old >> 48) + 1u == in_blk)
          desc_raise(P.desc[1] + 2 * (size_t)blk + m, kFlagAgg | ((blk_old & ((1ull << 48) - 1)) + total));
      }
    }
    {
      const uint32_t nk = __reduce_add_sync(0xffffffffu, n_kept), nm = __reduce_add_sync(0xffffffffu, n_matched);
      if (lane == 0) {
        if (m == 0 && nk) atomicAdd(&P.result->n_out, (unsigned long long)nk);
        if (nm) atomicAdd(&P.result->n_match[m], (unsigned long long)nm);
      }
    }
  } else if (is_scan) {
    // =============================== SCAN =======================================================
    const int sj = warp - NB;
    for (uint32_t k = sj;; k += kScanWarps) {
      const uint32_t j = k % kMeta;
      mbar_wait5(&bar_pub[j], (k / kMeta) & 1u);
      const uint32_t tile = meta[0][j].tile;
      if (tile == kEnd) break;
      EVT(tile, 8)
      const ulonglong2 g = lookback<PAIRED>(P.desc[0], P.desc[1], tile, meta[0][j].total, PAIRED ? meta[NM - 1][j].total : 0u, lane);
      if (lane == 0) {
        meta[0][j].gout = g.x;
        if (PAIRED) meta[NM - 1][j].gout = g.y;
        mbar_arrive(&bar_ready[j]);
      }
      EVT(tile, 9)
    }
  } else {
    // =============================== BACK =======================================================
    uint8_t* const gdst = P.out[m];
    for (uint32_t k = 0;; ++k) {
      const uint32_t s = k % kData, j = k % kMeta;
      const Meta& tm = meta[m][j];
      if (bj == 0) {  // one warp per mate polls; the others block in hardware on the named barrier
        mbar_wait5(&bar_full[m][s], (k / kData) & 1u);
        if (tm.tile != kEnd) mbar_wait5(&bar_ready[j], (k / kMeta) & 1u);
      }
      back_sync(m, nparts);
      const uint32_t tile = tm.tile;
      if (tile == kEnd) break;
      if (m == 0 && bj == 0) { EVT(tile, 10) }
      const uint32_t total = tm.total;
      const unsigned long long gout = tm.gout;
      if (lane == 0 && bj == 0 && tile == P.ntiles - 1) P.result->out_len[m] = gout + total;
      const bool fits = gout + total <= P.cap[m];
      if (!fits && lane == 0) atomicOr(&P.result->error, 1u);
      const uint32_t phase = (uint32_t)(gout & 15ull);
      if (fits && total) {
        const uint32_t in_s = smem_base + (m ? kData * P.in_cap[0] : 0) + s * P.in_cap[m];
        const uint32_t ooff = tm.ooff[lane];
        const bool kept = ooff != kNotKept;
        const uint32_t go = tm.goff[lane], hl = tm.hl[lane];
        uint32_t gn = __shfl_down_sync(0xffffffffu, go, 1);
        if (lane == (int)tm.nrec - 1) gn = tm.gend;
        RecView rv;
        rv.beg = in_s + (go - tm.gbase);
        rv.end = in_s + (gn - tm.gbase);
        rv.head_len = hl & 0x3FFFFFFFu;
        rv.fast = (hl >> 30) & 1u;
        const int ms = (hl >> 31) ? 0 : -1;
        rv.seq = rv.beg + 2u + rv.head_len;
        rv.seq_len = rv.qual_len = (rv.end - rv.seq - 4u) >> 1;
        rv.qual = rv.seq + rv.seq_len + 3u;
        if (kept && !rv.fast) rv = parse_record(rv.beg, rv.end);  // foreign framing: the general parser, on chip
        const uint32_t dst = out_s + phase + ooff;
        const bool unc = kept && rv.fast && ms < 0;
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, ooff, 1);
        const bool even = kept && lane < 31 && nxt != kNotKept && ((nxt - ooff) & 31u) == 0;
        const uint32_t rot = __popc(__ballot_sync(0xffffffffu, even)) >= 8 ? ((uint32_t)lane & 7u) + 1u : 0u;
        if (kept && !unc) emit_part(pg, rv, ms, trim, dst, (uint32_t)bj, nparts, rot);
        const unsigned um = __ballot_sync(0xffffffffu, unc);
        if (um) {
          const unsigned below = ~um & ((1u << lane) - 1u);
          const int i0 = unc ? (below ? 32 - __clz(below) : 0) : lane;
          const unsigned above = ~(um >> lane);
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
              OutStream os;
              os_begin(os, head ? run_d : run_d + tlo);
              os_run(os, head ? run_s : run_s + tlo, head ? pad : run_n - tlo);
              os_end(os);
            }
          }
          __syncwarp();
          const uint32_t nruns = __popc(starts);
          uint32_t t0 = (uint32_t)bj;
          for (uint32_t r = 0; r < nruns; ++r) {
            bulk16(run_desc[m][r], lane, t0, nparts);
            t0 = t0 == 0 ? nparts - 1u : t0 - 1u;
          }
        }
      }
      __syncwarp();
      if (m == 0 && bj == 0) { EVT(tile, 11) }
      if (lane == 0) {
        mbar_arrive(&bar_empty[m][s]);   // data slot free again
        mbar_arrive(&meta_empty[m][j]);  // ... and the tile's meta slot
      }
      back_sync(m, nparts);  // all parts written
      if (m == 0 && bj == 0) { EVT(tile, 12) }
      if (fits && total) {
        // ---- store: every back warp of the mate moves its share of the aligned middle, 16 bytes per lane
        // (512 consecutive bytes per instruction); the ragged edges go bytewise.  A TMA bulk store here
        // held its issuing warp for ~2 k cycles per tile (tools/phases.py), and the output tile could not
        // be reused before the engine had read it.
        // offsets relative to the 16-byte line the tile's output starts in
        uint8_t* const gl = gdst + (gout & ~15ull);
        const uint32_t sl = out_s;                     // shared address of gl[0] (the tile was laid out with this phase)
        const uint32_t x0 = phase, x1 = phase + total;  // the tile's bytes: gl[x0, x1)
        const uint32_t a0 = (x0 + 15u) & ~15u, a1 = x1 & ~15u;
        for (uint32_t x = a0 + 16u * (32u * (uint32_t)bj + (uint32_t)lane); x < a1; x += 512u * nparts) {
          const uint4 v = lds128(sl + x);
          asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gl + x), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        if (bj == 0) {
          if (a0 >= a1) {
            for (uint32_t x = x0 + lane; x < x1; x += 32) gl[x] = (uint8_t)lds8(sl + x);
          } else {
            uint32_t x = x0 + lane;
            if (x < a0) gl[x] = (uint8_t)lds8(sl + x);
            x = a1 + lane;
            if (x < x1) gl[x] = (uint8_t)lds8(sl + x);
          }
        }
        __syncwarp();
      }
      if (m == 0 && bj == 0) { EVT(tile, 13) }
    }
  }
}

}  // namespace v5
}  // namespace bk
