// Micro-benchmark: issue cost / throughput of small tcgen05.mma (kind::f16, bf16, M = 128) on one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../vision-transformers-pytorch_b200/csrc -o mma_probe mma_probe.cu
// For each shape: R back-to-back MMAs by one elected thread, then one commit; cycles from first issue to the commit's
// mbarrier completing (t_done) and to the last instruction having issued (t_issue), divided by R.
//   chains = number of distinct accumulators the sequence rotates over (1 = every MMA depends on the previous one)
#include "common.cuh"
#include <vector>
void vtb_set_error(const char*, ...) {}

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

template <int N, int CHAINS, int MODE>   // MODE 0: SS K-major both, 1: SS B MN-major, 2: TS (A in TMEM) B MN-major, 3: SS A MN-major, B MN-major
__global__ void __launch_bounds__(128, 1) probe(unsigned int* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if ((threadIdx.x & 31) == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(slot, 512);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  constexpr int R = 32;
  if (warp == 1) {
    constexpr bool a_mn = MODE == 3, b_mn = MODE >= 1, a_tmem = MODE == 2;
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, a_mn, b_mn);
    const uint64_t da = a_mn ? umma_desc_sw128(smem_u32(smem), 16384, 1024) : umma_desc_sw128(smem_u32(smem), 0, 1024);
    const uint64_t db = umma_desc_sw128(smem_u32(smem) + 32768, b_mn ? 16384 : 0, 1024);
    constexpr int stride = (N * CHAINS <= 384) ? N : 0;
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 3; ++rep) {   // last repetition is reported
      t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint32_t d = tm + 128 + (r % CHAINS) * stride;
          constexpr int dummy = 0; (void)dummy;
          const int ks = r & 3;
          if (a_tmem) umma_ts(d, tm + ks * 8, db + (b_mn ? ks * 128 : ks * 2), idesc, r >= CHAINS ? 1u : 0u);
          else umma_bf16(d, da + (a_mn ? ks * 128 : ks * 2), db + (b_mn ? ks * 128 : ks * 2), idesc, r >= CHAINS ? 1u : 0u);
        }
        umma_commit(bar);
      }
      __syncwarp();
      t1 = clock64();
      mbar_wait(bar, rep & 1);
      t2 = clock64();
    }
    if ((threadIdx.x & 31) == 0) { out[0] = (unsigned int)(t1 - t0); out[1] = (unsigned int)(t2 - t0); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

unsigned int* d_out;
template <int N, int CHAINS, int MODE>
void run() {
  const char* names[4] = {"SS A:K B:K", "SS A:K B:MN", "TS A:tmem B:MN", "SS A:MN B:MN"};
  cudaFuncSetAttribute(probe<N, CHAINS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe<N, CHAINS, MODE><<<1, 128, 100 * 1024>>>(d_out);
  unsigned int h[2];
  cudaError_t e = cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s at N=%d\n", cudaGetErrorString(e), N); exit(1); }
  printf("%4d %6d %22s %12.1f %12.1f\n", N, CHAINS, names[MODE], h[0] / 32.0, h[1] / 32.0);
}
template <int N> void run_n() {
  run<N, 1, 0>(); run<N, 2, 0>(); run<N, 1, 1>(); run<N, 1, 2>(); run<N, 2, 2>(); run<N, 1, 3>();
}
int main() {
  cudaMalloc(&d_out, 8);
  printf("%4s %6s %22s %12s %12s   (cycles per MMA, 32 unrolled MMAs, M = 128, K = 16)\n", "N", "chains", "mode", "issue", "done");
  run_n<16>(); run_n<32>(); run_n<48>(); run_n<64>(); run_n<96>(); run_n<128>(); run_n<208>(); run_n<256>();
  return 0;
}
