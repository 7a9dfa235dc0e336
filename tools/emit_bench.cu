// Dev microbenchmark: cycles per tile of the per-lane emit (bk_kernels.cuh emit_record) in isolation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/emit_bench tools/emit_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../barkit_b200/csrc/bk_kernels.cuh"

using namespace bk;

struct BP {
  bk_devprog pg;
  int iters;
  int matched;   // 1: R1-style matched+trim record, 0: unchanged record (one run)
  int rec_bytes; // record length in the input tile
  int warps;
  int parts, part;
};

__global__ void emit_bench(const __grid_constant__ BP P, const uint8_t* tile, uint32_t tile_bytes,
                           unsigned long long* cycles, uint32_t* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* in_s = smem + warp * 2 * 16384;
  uint8_t* out_s = in_s + 16384;
  for (uint32_t i = lane; i < tile_bytes / 4; i += 32)
    reinterpret_cast<uint32_t*>(in_s)[i] = reinterpret_cast<const uint32_t*>(tile)[i];
  __syncwarp();
  const uint32_t in_a = smem_u32(in_s), out_a = smem_u32(out_s);
  RecView rv = parse_record(in_a + lane * P.rec_bytes, in_a + (lane + 1) * P.rec_bytes);
  const int ms = P.matched ? 0 : -1;
  const uint32_t olen = out_bytes(P.pg, rv, ms, true);
  uint32_t incl = olen;
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const uint32_t ooff = incl - olen;
  __syncwarp();
  long long t0 = clock64();
  for (int it = 0; it < P.iters; ++it) {
    {
      const uint32_t d = out_a + (it & 15) + ooff;
      if (P.parts > 1) emit_part(P.pg, rv, ms, true, d, P.part);
      else for (int q = 0; q < 4; ++q) emit_part(P.pg, rv, ms, true, d, q);
    }
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
  if (out_s[lane] == 0xFF) sink[0] = 1;
}

int main(int argc, char** argv) {
  int ctas_per_sm = argc > 1 ? atoi(argv[1]) : 1;
  int warps = argc > 2 ? atoi(argv[2]) : 1;
  int matched = argc > 3 ? atoi(argv[3]) : 1;
  const int L = 150, H = argc > 4 ? atoi(argv[4]) : 23, rl = H + 2 * L + 6;
  std::vector<uint8_t> tile(32 * rl + 64, 'A');
  for (int r = 0; r < 32; ++r) {
    uint8_t* p = tile.data() + r * rl;
    p[0] = '@';
    memset(p + 1, 'h', H);
    p[1 + H] = '\n';
    memset(p + 2 + H, 'C', L);
    p[2 + H + L] = '\n';
    p[3 + H + L] = '+';
    p[4 + H + L] = '\n';
    memset(p + 5 + H + L, 'I', L);
    p[5 + H + 2 * L] = '\n';
  }
  BP P;
  memset(&P, 0, sizeof P);
  P.pg.present = 1;
  P.pg.anchor_start = 1;
  P.pg.total_len = 35;
  P.pg.ncaps = 2;
  P.pg.caps[0] = {0, 16, 2, {'C', 'B', 0}};
  P.pg.caps[1] = {23, 12, 3, {'U', 'M', 'I'}};
  P.pg.hdr_growth = 67;
  P.iters = 200;
  P.matched = matched;
  P.rec_bytes = rl;
  P.warps = warps;
  P.parts = argc > 5 ? atoi(argv[5]) : 1;
  P.part = argc > 6 ? atoi(argv[6]) : 0;
  uint8_t* d_tile;
  unsigned long long* d_cyc;
  uint32_t* d_sink;
  cudaMalloc(&d_tile, tile.size());
  cudaMalloc(&d_cyc, 8);
  cudaMalloc(&d_sink, 4);
  cudaMemcpy(d_tile, tile.data(), tile.size(), cudaMemcpyHostToDevice);
  cudaMemset(d_cyc, 0, 8);
  size_t smem = (size_t)warps * 2 * 16384;
  cudaFuncSetAttribute(emit_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = 148 * ctas_per_sm;
  uint32_t tb = (32 * rl + 63) & ~63u;
  emit_bench<<<grid, 32 * warps, smem>>>(P, d_tile, tb, d_cyc, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  unsigned long long cyc;
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("ctas/sm %d warps/cta %d matched %d part %d/%d: %.0f cycles per tile-emit per warp\n", ctas_per_sm, warps, matched, P.part, P.parts,
         (double)cyc / ((double)grid * warps * P.iters));
  return 0;
}
