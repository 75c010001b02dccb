"""TEST INFRASTRUCTURE ONLY — never imported by the product.

Compiles the SOURCE of a simple (no tensor-core / TMA / warp-collective) CUDA kernel file for the host with g++, so that
its index arithmetic, byte unpacking and table handling can be checked against the oracle in the CPU suite, where no GPU
exists.  The kernel file is copied unmodified into a scratch tree next to a shim `common.cuh` that maps the handful of CUDA
built-ins it uses onto plain C++ (IEEE float ops, std::log / std::sin, a loop over blocks); the
`<<<grid, block, smem, stream>>>` launch is rewritten to a call of that loop.  The threads of a block run as real host
threads meeting at a std::barrier for `__syncthreads()` (blocks run one after the other; `__shared__` = function-static).

This proves nothing about performance or about the GPU's libm; the `-m gpu` tests remain the parity tests proper.
"""
import ctypes as C
import os
import re
import shutil
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vision-transformers-pytorch_b200", "csrc")

SHIM = r"""
#pragma once
#include <algorithm>
#include <barrier>
#include <cfenv>
#include <cmath>
#include <memory>
#include <thread>
#include <vector>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
struct float3 { float x, y, z; };
static inline float3 make_float3(float a, float b, float c) { return {a, b, c}; }
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
struct dim3_ { unsigned x, y, z; };
static thread_local dim3_ threadIdx, blockIdx;
static dim3_ gridDim, blockDim;
struct uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return {a, b, c, d}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
typedef void* cudaStream_t;
using std::min; using std::max;
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline float __uint_as_float(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t v; std::memcpy(&v, &f, 4); return v; }
// round-toward-zero add: the exact sum of two floats fits a double; truncate it to float precision by hand
static inline float __fadd_rz(float a, float b) {
  const double e = (double)a + (double)b;
  float r = (float)e;  // nearest
  if ((double)r != e && std::fabs((double)r) > std::fabs(e)) r = std::nextafterf(r, 0.0f);
  return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
// volatile keeps the compiler from contracting / re-associating the explicitly rounded steps
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static std::barrier<>* g_block_barrier = nullptr;
static inline void __syncthreads() { g_block_barrier->arrive_and_wait(); }
static char g_err[512];
static inline void vtb_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, 512, fmt, ap); va_end(ap); }
#define VTB_CHECK(cond, code, ...) do { if (!(cond)) { vtb_set_error(__VA_ARGS__); return (code); } } while (0)
#define VTB_LAUNCH_CHECK() do {} while (0)
static inline int vtb_num_sms() { return 3; }  // small on purpose: exercises the grid-stride loop
template <class K, class... A> static void cpu_launch(K kernel, int grid, int block, A... args) {
  gridDim = {(unsigned)grid, 1, 1}; blockDim = {(unsigned)block, 1, 1};
  for (int b = 0; b < grid; ++b) {
    std::barrier<> bar(block);
    g_block_barrier = &bar;
    std::vector<std::thread> ts;
    for (int t = 0; t < block; ++t)
      ts.emplace_back([=] { blockIdx = {(unsigned)b, 0, 0}; threadIdx = {(unsigned)t, 0, 0}; kernel(args...); });
    for (auto& th : ts) th.join();
  }
}
"""

LAUNCH = re.compile(r"(\w+(?:<[^<>]*>)?)<<<\s*([^,]+),\s*([^,]+),\s*[^,]+,\s*[^>]+>>>\(")


def build(cu_name):
    """-> ctypes handle of the host build of csrc/<cu_name> (cached per process)."""
    if cu_name in _cache:
        return _cache[cu_name]
    tmp = tempfile.mkdtemp(prefix="vtb_emul_")
    src_dir = os.path.join(tmp, "pkg", "csrc")
    os.makedirs(src_dir)
    os.makedirs(os.path.join(tmp, "include"))
    shutil.copy(os.path.join(ROOT, "include", "vtb200.h"), os.path.join(tmp, "include", "vtb200.h"))
    with open(os.path.join(src_dir, "common.cuh"), "w") as f:
        f.write(SHIM)
    text = open(os.path.join(CSRC, cu_name)).read()
    text, n = LAUNCH.subn(r"cpu_launch(\1, \2, \3, ", text)
    assert n >= 1, "no kernel launch found"
    cpp = os.path.join(src_dir, cu_name.replace(".cu", ".cpp"))
    with open(cpp, "w") as f:
        f.write(text)
    so = os.path.join(tmp, "emul.so")
    subprocess.run(["g++", "-O1", "-std=c++20", "-pthread", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
    _cache[cu_name] = C.CDLL(so)
    shutil.rmtree(tmp, ignore_errors=True)  # the mapping stays valid after the file is unlinked
    return _cache[cu_name]


_cache = {}
