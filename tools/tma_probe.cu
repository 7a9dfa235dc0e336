// Probe: which byte-granular tensor-map layouts does the TMA engine of this GPU accept for loads that start at an
// arbitrary byte?  (The bypass form of the extract kernel moves a mate's kept records with such loads.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tools/tma_probe.cu && ./tma_probe A|B|C
//   A: 1-D tensor of bytes, box 256            B: 2-D {256, rows}, stride 256 (plain rows), box {256,1}
//   C: 2-D {512, rows}, stride 256 (overlapping rows), box {256,1}, any column
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int row, uint8_t* out) {
  __shared__ __align__(1024) uint8_t buf[256];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 256;" ::"r"(smem_u32(&bar)) : "memory");
    if (RANK == 1)
      asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(smem_u32(buf)),
                   "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(smem_u32(&bar))
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(buf)),
                   "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(row), "r"(smem_u32(&bar))
                   : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
  out[threadIdx.x] = buf[threadIdx.x];
}

int main(int argc, char** argv) {
  const char mode = argc > 1 ? argv[1][0] : 'A';
  const size_t N = 1 << 20;
  uint8_t* h = new uint8_t[N];
  for (size_t i = 0; i < N; ++i) h[i] = (uint8_t)(i * 131u + (i >> 8));
  uint8_t *d, *d_out;
  cudaMalloc(&d, N + 16);
  cudaMalloc(&d_out, 256);
  cudaMemcpy(d, h, N, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("%c: no entry point\n", mode); return 1; }
  EncodeTiledFn fn = (EncodeTiledFn)p;
  CUtensorMap map;
  CUresult r;
  const cuuint32_t estr[2] = {1, 1};
  const size_t X = 12345 * 7 + 3;  // an odd byte offset
  int x, row;
  const cuuint64_t one_stride[1] = {16};
  if (mode == 'A') {            // 1-D, strides pointer null
    const cuuint64_t dims[1] = {N};
    const cuuint32_t box[1] = {256};
    r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d, dims, nullptr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    x = (int)X; row = 0;
  } else if (mode == 'H') {     // 1-D, strides pointer non-null
    const cuuint64_t dims[1] = {N};
    const cuuint32_t box[1] = {256};
    r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d, dims, one_stride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    x = (int)X; row = 0;
  } else if (mode == 'G') {     // the whole text as ONE row of a 2-D tensor: {N, 1}, stride N
    const cuuint64_t dims[2] = {N, 1};
    const cuuint64_t strides[1] = {N};
    const cuuint32_t box[2] = {256, 1};
    r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    x = (int)X; row = 0;
  } else {
    // B: {256, rows} stride 256, column 0        C: {512, rows} stride 256 (overlap), odd column
    // D: {512, rows} stride 512, odd column      E: {512, rows} stride 256 (overlap), column multiple of 16
    const bool overlap = mode == 'C' || mode == 'E';
    const cuuint64_t w = mode == 'B' ? 256 : 512, st = overlap || mode == 'B' ? 256 : 512;
    const cuuint64_t dims[2] = {w, (N - 512) / st};
    const cuuint64_t strides[1] = {st};
    const cuuint32_t box[2] = {256, 1};
    r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (mode == 'B') { x = 0; row = (int)(X >> 8); }
    else if (mode == 'D') { x = (int)(X & 255); row = (int)(X >> 9); }
    else if (mode == 'E') { x = (int)(X & 240); row = (int)(X >> 8); }
    else { x = (int)(X & 255); row = (int)(X >> 8); }
  }
  if (r != CUDA_SUCCESS) { printf("%c: encode failed (%d)\n", mode, (int)r); return 1; }
  if (mode == 'A' || mode == 'H') probe<1><<<1, 256>>>(map, x, row, d_out); else probe<2><<<1, 256>>>(map, x, row, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%c: kernel failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
  uint8_t got[256];
  cudaMemcpy(got, d_out, 256, cudaMemcpyDeviceToHost);
  const size_t want0 = mode == 'B' ? (X >> 8) * 256 : mode == 'D' ? (X >> 9) * 512 + (X & 255) : mode == 'E' ? (X >> 8) * 256 + (X & 240) : X;
  int bad = 0;
  for (int i = 0; i < 256; ++i) bad += got[i] != h[want0 + i];
  printf("%c: ok, %d of 256 bytes differ (box at byte %zu)\n", mode, bad, want0);
  return bad != 0;
}
