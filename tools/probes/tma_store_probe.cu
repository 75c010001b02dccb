// Micro-benchmark: HBM write bandwidth of TMA tensor stores as a function of the box shape (what a GEMM epilogue that
// stages 128-byte-row slabs can reach), next to plain 16-byte stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda -o tma_store_probe tma_store_probe.cu
// Output matrix: bf16 [M = 802816, N = 768] row-major (1.23 GB), written once per launch by 148 persistent CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// each CTA walks tiles of [rows x cols] elements; per tile it issues (rows / box_r) * (cols / box_c) tensor stores from one
// shared-memory slab (contents irrelevant) and keeps at most `depth` bulk groups in flight
__global__ void __launch_bounds__(128, 1)
tma_store(const __grid_constant__ CUtensorMap map, int M, int N, int tile_r, int tile_c, int box_r, int box_c, int slab_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < slab_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int tiles_c = N / tile_c, tiles = (M / tile_r) * tiles_c;
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int r0 = (t / tiles_c) * tile_r, c0 = (t % tiles_c) * tile_c;
    uint32_t off = 0;
    for (int r = 0; r < tile_r; r += box_r)
      for (int c = 0; c < tile_c; c += box_c) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map), "r"(s + off),
                     "r"(c0 + c), "r"(r0 + r)
                     : "memory");
        off += box_r * box_c * 2;
        if (off + box_r * box_c * 2 > (uint32_t)slab_bytes) off = 0;
      }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void stg(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4((unsigned)i, 1, 2, 3);
}

int main() {
  const int M = 802816, N = 768;
  const size_t bytes = (size_t)M * N * 2;
  void* out;
  cudaMalloc(&out, bytes);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](auto f) { f(); float best = 1e9; for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; } return best; };
  cudaFuncSetAttribute(tma_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int box_r, box_c, swz, tile_r, tile_c; const char* name; };
  Cfg cfgs[] = {
      {32, 64, 1, 128, 256, "box 32 rows x 128 B (swizzle 128B)  [GEMM epilogue today]"},
      {128, 64, 1, 128, 256, "box 128 rows x 128 B (swizzle 128B)"},
      {256, 64, 1, 256, 256, "box 256 rows x 128 B (swizzle 128B)"},
      {32, 128, 0, 128, 256, "box 32 rows x 256 B (no swizzle)"},
      {32, 256, 0, 128, 256, "box 32 rows x 512 B (no swizzle)"},
      {128, 256, 0, 128, 256, "box 128 rows x 512 B (no swizzle)"},
      {64, 256, 0, 128, 768, "box 64 rows x 512 B (no swizzle), tile = full 768-column rows"},
      {32, 32, 2, 128, 256, "box 32 rows x 64 B (swizzle 64B)"},
  };
  for (auto& c : cfgs) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {(cuuint32_t)c.box_c, (cuuint32_t)c.box_r};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = c.swz == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : c.swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-70s encode failed (%d)\n", c.name, (int)r); continue; }
    const int slab = 128 * 1024;
    float ms = time([&] { tma_store<<<148, 128, slab + 1024>>>(map, M, N, c.tile_r, c.tile_c, c.box_r, c.box_c, slab); });
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-70s %7.1f GB/s  %s\n", c.name, bytes / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  float ms = time([&] { stg<<<148 * 16, 512>>>((uint4*)out, bytes / 16); });
  printf("%-70s %7.1f GB/s\n", "plain 16-byte stores (grid 2368 x 512)", bytes / ms / 1e6);
  return 0;
}
