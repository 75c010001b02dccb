// Micro-benchmark: HBM bandwidth of write-only / read-only / copy streams on one B200 (what bounds a write-heavy kernel:
// the copy figure in MEASURED_PEAKS.json counts read + write bytes of a 1:1 stream).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_probe hbm_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill(uint4* p, size_t n, uint4 v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void sum(const uint4* p, size_t n, unsigned* out) {
  unsigned a = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(p + i);
    a += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (a == 0x12345678u) *out = a;
}
__global__ void copy(const uint4* s, uint4* d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = __ldg(s + i);
}
// write-heavy mix of a narrow-K GEMM epilogue: read 1 byte for every `ratio` bytes written
__global__ void mix(const uint4* s, uint4* d, size_t n, int ratio) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = make_uint4((unsigned)i, 1, 2, 3);
    if (i % ratio == 0) v = __ldg(s + i / ratio);
    d[i] = v;
  }
}
int main() {
  const size_t bytes = (size_t)2 << 30, n = bytes / 16;
  uint4 *a, *b; unsigned* o;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](auto f) { f(); f(); float best = 1e9; for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; } return best; };
  for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
    float tf = time([&] { fill<<<blocks, 512>>>(a, n, make_uint4(1, 2, 3, 4)); });
    float ts = time([&] { sum<<<blocks, 512>>>(a, n, o); });
    float tc = time([&] { copy<<<blocks, 512>>>(a, b, n); });
    float tm = time([&] { mix<<<blocks, 512>>>(a, b, n, 8); });
    printf("grid %5d x 512: write-only %7.1f GB/s   read-only %7.1f GB/s   copy (r+w) %7.1f GB/s   write + 1/8 read %7.1f GB/s\n", blocks,
           bytes / tf / 1e6, bytes / ts / 1e6, 2.0 * bytes / tc / 1e6, bytes * 1.125 / tm / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
