// match_kernel: the match of one mate's reads as a pass of its own (pattern.rs:225-230, parse.rs:58-68).
//
// A pattern that needs a search -- not anchored, a reverse-complement retry, variable repetition -- costs ten to a
// hundred times an anchored test in code as well as in time, and inside extract_kernel that code competes with the
// emit path for the SM's instruction cache (point 3: the paired general kernel ran at half the
// anchored kernel's rate with an 87 % instruction-cache hit rate).  Such a mate is therefore matched HERE, ahead of
// the extract launches on the same stream, and they take the result from the table this kernel leaves:
//
//   match[i] = start | variant << 24 of read i's leftmost-first match, or -1
//
// No ordering, no look-back, no output text: a warp owns a tile of up to 32 consecutive records, one lane issues a
// single 1-D TMA bulk load of the tile's contiguous text range into the warp's own shared-memory buffer, the lanes
// parse and match one read each with the extract kernel's own building blocks (parse_record, match_unanchored,
// match_read, match_variants: same code, same semantics), and the 32 results go out as one coalesced store.  Warps
// are independent (grid-stride over tiles), so every SM keeps 16-24 loads in flight without any pipeline logic.
#pragma once
#include "bk_extract.cuh"

namespace bk {

constexpr int kMatchWarps = 8;

struct MatchParams {
  bk_devprog prog;
  const uint8_t* text;
  const uint32_t* off;
  int32_t* match;          // out
  uint4* spans;            // out, programs without a fixed shape (bk_free_shape): match length and capture spans of every match (KParams::xtab)
  DevResult* result;       // error bit 1 << 1: a tile did not fit the warp's buffer (as in extract_kernel)
  const bk_devprog* vprog; // the pattern's variants when nvar > 1
  uint32_t nvar;
  uint32_t n;
  uint32_t ntiles;
  uint32_t tile_recs;      // <= 32
  uint32_t cap;            // bytes of one warp's buffer (a multiple of 16)
  uint32_t rc;             // reverse-complement retry (parse.rs:61-66)
  const uint2* vmcode;     // the general form (prog.vm_len > 0): its instructions (bk_vmop), match_kernel<true>
};

// VM = true: the general form only (match_vm keeps a backtracking stack in local memory, which the other matchers'
// launches should not have to carry)
template <bool VM>
__global__ void __launch_bounds__(kMatchWarps * 32) match_kernel(const __grid_constant__ MatchParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(16) uint8_t tabs[kTabBytes];
  __shared__ __align__(8) uint64_t bar[kMatchWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t W = blockDim.x >> 5;  // (fewer than kMatchWarps when records are very long)
  const bk_devprog& pg = P.prog;
  if (warp == 0) {
    fill_tab(tabs, pg, lane, true);
    if (lane < kMatchWarps) mbar_init(&bar[lane], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const uint32_t tab = smem_u32(tabs);
  const uint32_t buf = smem_u32(smem) + (uint32_t)warp * P.cap;
  const uint32_t R = P.tile_recs;
  const bool unanch = !pg.anchor_start && !pg.anchor_end;
  const bool var = P.nvar > 1u;
  // variants that share their first literal can be searched with the bit-parallel scan (match_variants_scan)
  __shared__ uint8_t s_voff[BK_MAX_VARIANTS];
  __shared__ uint32_t s_scan;  // 0: no; else 1 + the smallest literal offset
  if (warp == 0 && var) {
    bool ok = unanch && lane < (int)P.nvar;
    uint32_t off = 0;
    if (lane < (int)P.nvar) {
      const bk_devprog& pv = P.vprog[lane];
      const bk_op a = pv.ops[0], z = P.vprog[0].ops[0];
      ok = ok && pv.nops && !pv.anchor_start && !pv.anchor_end && bs_eligible(a.kind, a.len, a.k) && a.len == z.len && a.k == z.k;
      for (uint32_t i = 0; ok && i < a.len; ++i) ok = pv.lit[a.idx + i] == P.vprog[0].lit[z.idx + i];
      off = a.off;
      s_voff[lane] = (uint8_t)off;
    }
    const bool all_ok = __all_sync(0xffffffffu, ok || lane >= (int)P.nvar) && unanch;
    const uint32_t omin = __reduce_min_sync(0xffffffffu, lane < (int)P.nvar ? off : 0xFFFFFFFFu);
    const uint32_t omax = __reduce_max_sync(0xffffffffu, lane < (int)P.nvar ? off : 0u);
    if (lane == 0) s_scan = all_ok && omax - omin <= 24u && omax < 256u ? 1u + omin : 0u;
  }
  __shared__ __align__(8) uint2 s_vm[VM ? BK_VM_MAX : 1];
  if (VM)
    for (uint32_t i = threadIdx.x; i < pg.vm_len; i += blockDim.x) s_vm[i] = __ldg(P.vmcode + i);
  __syncthreads();
  uint32_t ph = 0;
  for (uint32_t tile = blockIdx.x * W + warp; tile < P.ntiles; tile += gridDim.x * W) {
    const uint32_t r0 = tile * R, nrec = min(R, P.n - r0);
    const bool active = lane < (int)nrec;
    const uint32_t o = active ? __ldg(P.off + r0 + lane) : 0u;
    const uint32_t oe = __ldg(P.off + r0 + nrec);
    uint32_t onext = __shfl_down_sync(0xffffffffu, o, 1);
    if (lane == (int)nrec - 1) onext = oe;
    const uint32_t gbase = __shfl_sync(0xffffffffu, o, 0) & ~15u;
    const uint32_t bytes = ((oe + 15u) & ~15u) - gbase;
    int ms = -1;
    if (bytes > P.cap) {  // (the host sized the buffer from max_rec_bytes)
      if (lane == 0) atomicOr(&P.result->error, 2u);
    } else {
      if (lane == 0) {
        mbar_arrive_expect_tx(&bar[warp], bytes);
        bulk_g2s(buf, P.text + gbase, bytes, &bar[warp]);
      }
      mbar_wait(&bar[warp], ph);
      ph ^= 1u;
      RecView rv;
      rv.beg = rv.end = rv.fast = rv.head_len = rv.seq = rv.seq_len = rv.qual = rv.qual_len = 0;
      if (active) rv = parse_record(buf + (o - gbase), buf + (onext - gbase));
      __syncwarp();
      const bool can = active && rv.qual_len == rv.seq_len;  // unequal lengths never leave bk_tokenize
      if (VM) {  // the general form: a backtracking program
        uint4 sp = make_uint4(0, 0, 0, 0);
        uint32_t gave_up = 0;
        if (can) {
          ms = match_vm<false>(tab, smem_u32(s_vm), pg.vm_len, pg.ncaps, rv.seq, rv.seq_len, &sp, &gave_up);
          if (ms == -1 && P.rc) {
            ms = match_vm<true>(tab, smem_u32(s_vm), pg.vm_len, pg.ncaps, rv.seq, rv.seq_len, &sp, &gave_up);
            // parse.rs:166 on the forward bytes under the offsets (F7)
            const uint32_t o3[3] = {sp.x >> 16, sp.y >> 16, sp.z >> 16}, l3[3] = {sp.y & 0xFFFFu, sp.z & 0xFFFFu, sp.w & 0xFFFFu};
            for (uint32_t c = 0; ms >= 0 && c < pg.ncaps; ++c)
              if (!span_utf8_ok(rv.seq + (uint32_t)ms + o3[c], rv.seq + (uint32_t)ms + o3[c] + l3[c])) ms = -1;
          }
        }
        if (ms < 0) ms = -1;
        if (gave_up) atomicOr(&P.result->error, 8u);
        if (active) P.spans[r0 + lane] = sp;
      } else if (pg.nins) {  // unbounded repetition: backtracking over the insertion points
        unsigned long long ex = 0;
        bool was_rc = false;
        if (can) {
          ms = match_par<false>(tab, rv.seq, rv.seq_len, &ex);
          if (ms < 0 && P.rc) {
            ms = match_par<true>(tab, rv.seq, rv.seq_len, &ex);
            was_rc = true;
          }
        }
        uint32_t off[3] = {0, 0, 0}, clen[3] = {0, 0, 0}, tlen = pg.total_len;
        if (ms >= 0) {
          for (uint32_t i = 0; i < pg.nins; ++i) tlen += (uint32_t)(ex >> (16u * i)) & 0xFFFFu;
          for (uint32_t c = 0; c < pg.ncaps; ++c) {
            par_cap(pg, c, ex, off[c], clen[c]);
            // parse.rs:166 on the forward bytes under the offsets (F7)
            if (was_rc && !span_utf8_ok(rv.seq + (uint32_t)ms + off[c], rv.seq + (uint32_t)ms + off[c] + clen[c])) ms = -1;
          }
        }
        if (active) P.spans[r0 + lane] = pack_spans(tlen, off, clen);
      } else if (var && s_scan) {
        ms = match_variants_scan(P.vprog, P.nvar, tab, smem_u32(s_voff), s_scan - 1u, rv.seq, rv.seq_len, can);
        if (can && ms < 0 && P.rc) {
          ms = match_variants<true>(P.vprog, P.nvar, tab, rv.seq, rv.seq_len);
          if (ms >= 0 && !caps_utf8_ok(P.vprog[ms >> 24], rv.seq, (uint32_t)ms & 0xFFFFFFu)) ms = -1;
        }
      } else if (var) {
        if (can) {
          ms = match_variants<false>(P.vprog, P.nvar, tab, rv.seq, rv.seq_len);
          if (ms < 0 && P.rc) {
            ms = match_variants<true>(P.vprog, P.nvar, tab, rv.seq, rv.seq_len);
            if (ms >= 0 && !caps_utf8_ok(P.vprog[ms >> 24], rv.seq, (uint32_t)ms & 0xFFFFFFu)) ms = -1;
          }
        }
      } else {
        if (unanch) ms = match_unanchored_rd(tab, rv.seq, rv.seq_len, can, lane);
        else if (can) ms = match_read_shared(tab, rv.seq, rv.seq_len);
        if (can && ms < 0 && P.rc) {
          ms = match_read<true>(tab, rv.seq, rv.seq_len);
          if (ms >= 0 && !caps_utf8_ok(pg, rv.seq, (uint32_t)ms)) ms = -1;
        }
      }
      __syncwarp();  // every lane has finished with the buffer before the next tile's load overwrites it
    }
    if (active) P.match[r0 + lane] = ms;
  }
}

// the matchers' error bits into the batch's result (enqueue_tabled)
__global__ void merge_error_kernel(DevResult* result, const DevResult* merr) {
  if (merr->error) atomicOr(&result->error, merr->error);
}

}  // namespace bk
