// Dev microbenchmark: cycles per tile of the front warp's per-lane parse + match in isolation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/front_bench tools/front_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../barkit_b200/csrc/bk_kernels.cuh"

using namespace bk;

struct FP {
  bk_devprog pg;
  int iters;
  int rec_bytes;
  int what;  // 0 parse+match, 1 parse only, 2 match only
};

__global__ void front_bench(const __grid_constant__ FP P, const uint8_t* tile, uint32_t tile_bytes,
                            unsigned long long* cycles, int* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint8_t tabs[kTabBytes];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* in_s = smem + warp * 16384;
  for (uint32_t i = lane; i < tile_bytes / 4; i += 32)
    reinterpret_cast<uint32_t*>(in_s)[i] = reinterpret_cast<const uint32_t*>(tile)[i];
  for (int c = threadIdx.x; c < 256; c += blockDim.x) {
    uint32_t mask = 0;
    if (c < 128)
      for (int k = 0; k < BK_MAX_CLASSES; ++k) mask |= ((P.pg.cls[k][c >> 5] >> (c & 31)) & 1u) << k;
    tabs[c] = (uint8_t)mask;
    tabs[kTabComp + c] = (uint8_t)complement((uint32_t)c);
  }
  __syncthreads();
  const uint32_t in_a = smem_u32(in_s), tab = smem_u32(tabs);
  RecView rv = parse_record(in_a + lane * P.rec_bytes, in_a + (lane + 1) * P.rec_bytes);
  int acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < P.iters; ++it) {
    if (P.what != 2) rv = parse_record(in_a + lane * P.rec_bytes + (it & 0), in_a + (lane + 1) * P.rec_bytes);
    if (P.what != 1) acc += match_read<false>(P.pg, tab, rv.seq, rv.seq_len);
    __syncwarp();
  }
  long long t1 = clock64();
  if (lane == 0) atomicAdd(cycles, (unsigned long long)(t1 - t0));
  if (acc == 12345 || rv.seq_len == 77777) sink[0] = acc;
}

int main(int argc, char** argv) {
  int ctas_per_sm = argc > 1 ? atoi(argv[1]) : 1;
  int warps = argc > 2 ? atoi(argv[2]) : 1;
  int what = argc > 3 ? atoi(argv[3]) : 0;
  const int L = 150, H = 23, rl = H + 2 * L + 6;
  std::vector<uint8_t> tile(32 * rl + 64, 'A');
  for (int r = 0; r < 32; ++r) {
    uint8_t* p = tile.data() + r * rl;
    p[0] = '@';
    memset(p + 1, 'h', H);
    p[1 + H] = '\n';
    for (int i = 0; i < L; ++i) p[2 + H + i] = "ACGT"[(i * 7 + r) & 3];
    memcpy(p + 2 + H + 16, "ATGCCAT", 7);
    p[2 + H + L] = '\n';
    p[3 + H + L] = '+';
    p[4 + H + L] = '\n';
    memset(p + 5 + H + L, 'I', L);
    p[5 + H + 2 * L] = '\n';
  }
  FP P;
  memset(&P, 0, sizeof P);
  // ^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12}), k = 1, as bk_pattern.cc lowers it
  P.pg.present = 1;
  P.pg.anchor_start = 1;
  P.pg.total_len = 35;
  P.pg.nops = 3;
  P.pg.ops[0] = {16, 7, BK_OP_LIT, 1, 0};
  P.pg.ops[1] = {0, 16, BK_OP_CLASS, 0, 0};
  P.pg.ops[2] = {23, 12, BK_OP_CLASS, 0, 0};
  memcpy(P.pg.lit, "ATGCCAT", 7);
  for (const char* c = "ATGCN"; *c; ++c) P.pg.cls[0][*c >> 5] |= 1u << (*c & 31);
  P.pg.ncaps = 2;
  P.pg.caps[0] = {0, 16, 2, {'C', 'B', 0}};
  P.pg.caps[1] = {23, 12, 3, {'U', 'M', 'I'}};
  P.pg.hdr_growth = 67;
  P.iters = 200;
  P.rec_bytes = rl;
  P.what = what;
  uint8_t* d_tile;
  unsigned long long* d_cyc;
  int* d_sink;
  cudaMalloc(&d_tile, tile.size());
  cudaMalloc(&d_cyc, 8);
  cudaMalloc(&d_sink, 4);
  cudaMemcpy(d_tile, tile.data(), tile.size(), cudaMemcpyHostToDevice);
  cudaMemset(d_cyc, 0, 8);
  size_t smem = (size_t)warps * 16384;
  cudaFuncSetAttribute(front_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = 148 * ctas_per_sm;
  uint32_t tb = (32 * rl + 63) & ~63u;
  front_bench<<<grid, 32 * warps, smem>>>(P, d_tile, tb, d_cyc, d_sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  unsigned long long cyc;
  cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("ctas/sm %d warps/cta %d what %d: %.0f cycles per tile per warp\n", ctas_per_sm, warps, what,
         (double)cyc / ((double)grid * warps * P.iters));
  return 0;
}
