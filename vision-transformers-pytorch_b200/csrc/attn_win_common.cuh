// Helpers shared by the tcgen05 window (attention_win_tc.cu) and halo (attention_halo_tc.cu) attention kernels:
// non-blocking barrier probes, cp.async pieces, TMEM stores, A-from-TMEM UMMA, register re-balancing.
#pragma once
#include "common.cuh"

namespace {

constexpr int WT_THREADS = 512;
constexpr float WT_L2E = 1.4426950408889634f;
constexpr float WT_LN2 = 0.6931471805599453f;

// a / b for 0 <= a < 2^23 with inv = 1 / b (one multiply and a fix-up instead of an integer division)
__device__ __forceinline__ int wt_div(int a, int b, float inv) {
  int q = __float2int_rz((float)a * inv);
  const int r = a - q * b;
  if (r < 0) --q;
  else if (r >= b) ++q;
  return q;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// arrive on `bar` (without touching its pending count) once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void* gp) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(gp) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* gp) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gp) : "memory");
}
// a polling role that makes no progress for ~2 s traps (-> launch error) instead of hanging the GPU
struct WtWatchdog {
  long long t0; uint32_t spins;
  __device__ __forceinline__ void reset() { spins = 0; }
  __device__ __forceinline__ void idle() {
    if (spins == 0) t0 = clock64();
    if ((++spins & 4095u) == 0 && clock64() - t0 > 4000000000LL) __trap();
  }
};
__device__ __forceinline__ unsigned long long wt_lds_u64(uint32_t saddr) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint32_t sw128(int row, int chunk) {
  return (uint32_t)row * 128u + ((uint32_t)(chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ float wt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void wt_tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void wt_tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void wt_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 wt_lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void wt_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int N> __device__ __forceinline__ void wt_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void wt_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void wt_proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 32 fp32 -> bf16 -> 64 contiguous bytes of global memory (no scaling)
__device__ __forceinline__ void wt_store_row32_raw(bf16* dst, const uint32_t (&a)[32]) {
#pragma unroll
  for (int e = 0; e < 32; e += 8)
    *reinterpret_cast<uint4*>(dst + e) =
        make_uint4(pack_bf16(__uint_as_float(a[e]), __uint_as_float(a[e + 1])),
                   pack_bf16(__uint_as_float(a[e + 2]), __uint_as_float(a[e + 3])),
                   pack_bf16(__uint_as_float(a[e + 4]), __uint_as_float(a[e + 5])),
                   pack_bf16(__uint_as_float(a[e + 6]), __uint_as_float(a[e + 7])));
}
// release a barrier once per warp after every lane is done
__device__ __forceinline__ void wt_warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
// 32 fp32 -> bf16 (x scale) -> 64 contiguous bytes of global memory
__device__ __forceinline__ void wt_store_row32(bf16* dst, const uint32_t (&a)[32], float s) {
#pragma unroll
  for (int e = 0; e < 32; e += 8)
    *reinterpret_cast<uint4*>(dst + e) =
        make_uint4(pack_bf16(__uint_as_float(a[e]) * s, __uint_as_float(a[e + 1]) * s),
                   pack_bf16(__uint_as_float(a[e + 2]) * s, __uint_as_float(a[e + 3]) * s),
                   pack_bf16(__uint_as_float(a[e + 4]) * s, __uint_as_float(a[e + 5]) * s),
                   pack_bf16(__uint_as_float(a[e + 6]) * s, __uint_as_float(a[e + 7]) * s));
}

}  // namespace
