// Memory-bound helpers around the GEMM / attention kernels: casts, bias-gradient column sums,
// patch gather/scatter (conv k=s as GEMM operand), cls-row fill, row pooling.
#include "common.cuh"
#include "../../include/vtb200.h"
#include <stdarg.h>
#include <string.h>

// ------------------------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";
void vtb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* vtb_last_error(void) { return g_err; }
extern "C" int vtb_version(void) { return 100; }
int vtb_gemm_init();
extern "C" int vtb_init(void) { return vtb_gemm_init(); }

namespace {

inline int grid_for(long n, int threads, int per_thread = 1) {
  long b = (n + (long)threads * per_thread - 1) / ((long)threads * per_thread);
  const long cap = (long)(vtb_num_sms() > 0 ? vtb_num_sms() : 148) * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__global__ void cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
  const long n4 = n >> 2;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

__global__ void cast2d_kernel(const float* __restrict__ src, long lds, bf16* __restrict__ dst, long ldd,
                              long rows, int cols) {
  const long total = rows * cols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / cols;
    const int c = (int)(i - r * cols);
    dst[r * ldd + c] = __float2bfloat16(src[r * lds + c]);
  }
}

__global__ void scale_cast_kernel(const float* __restrict__ src, const float* __restrict__ row_scale,
                                  int rows_per_scale, long rows, int cols, bf16* __restrict__ dst) {
  const int c4 = cols >> 2;
  const long total = rows * c4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / c4;
    const float sc = row_scale ? __ldg(row_scale + r / rows_per_scale) : 1.f;
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(v.x * sc, v.y * sc), pack_bf16(v.z * sc, v.w * sc));
  }
}

// g = bf16(src * scale[row]) and per-column sums of the ROUNDED values; thread = 4 consecutive columns,
// blockDim.y row lanes, rows partitioned over blockIdx.y; CTA-level reduction in smem then one atomic per column.
__global__ void __launch_bounds__(256)
scale_cast_colsum_kernel(const float* __restrict__ src, const float* __restrict__ row_scale, int rows_per_scale,
                         long rows, int cols, bf16* __restrict__ dst, float* __restrict__ colsum, long rows_per_block) {
  __shared__ float part[8][32][4];
  const int c4 = blockIdx.x * 32 + threadIdx.x;  // group of 4 columns
  const int rl = threadIdx.y;
  const long r0 = (long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (4 * c4 < cols) {
    for (long r = r0 + rl; r < r1; r += 8) {
      const float sc = row_scale ? __ldg(row_scale + r / rows_per_scale) : 1.f;
      const float4 v = *reinterpret_cast<const float4*>(src + r * cols + 4 * c4);
      const uint32_t lo = pack_bf16(v.x * sc, v.y * sc), hi = pack_bf16(v.z * sc, v.w * sc);
      *reinterpret_cast<uint2*>(dst + r * cols + 4 * c4) = make_uint2(lo, hi);
      const float2 f0 = unpack_bf16(lo), f1 = unpack_bf16(hi);
      a0 += f0.x; a1 += f0.y; a2 += f1.x; a3 += f1.y;
    }
  }
  part[rl][threadIdx.x][0] = a0; part[rl][threadIdx.x][1] = a1;
  part[rl][threadIdx.x][2] = a2; part[rl][threadIdx.x][3] = a3;
  __syncthreads();
  if (rl == 0 && 4 * c4 < cols) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) v += part[i][threadIdx.x][j];
      atomicAdd(colsum + 4 * c4 + j, v);
    }
  }
}

__global__ void silu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = silu_f(x[i]);
}
__global__ void silu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                float* __restrict__ dx, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dx[i] = dy[i] * silu_grad_f(x[i]);
}

// out[n] += sum_m X[m, n]; each thread owns 8 consecutive columns (one 16-byte load per row), 8 row lanes
// per CTA, rows partitioned over grid.y; 4 independent row loads in flight per thread.
constexpr int CS_ROWS = 8;
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ X, long M, int N, int ld, float* __restrict__ out,
              long rows_per_block) {
  __shared__ float part[CS_ROWS][32][8];
  const int c8 = blockIdx.x * 32 + (threadIdx.x & 31);  // group of 8 columns
  const int rl = threadIdx.x >> 5;
  const long r0 = (long)blockIdx.y * rows_per_block;
  const long r1 = min(M, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const bool vec = (8 * c8 + 8 <= N);
  if (vec) {
    long r = r0 + rl;
    for (; r + 3 * CS_ROWS < r1; r += 4 * CS_ROWS) {
      uint4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = *reinterpret_cast<const uint4*>(X + (r + u * CS_ROWS) * ld + 8 * c8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t w[4] = {t[u].x, t[u].y, t[u].z, t[u].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_bf16(w[q]);
          acc[2 * q] += f.x; acc[2 * q + 1] += f.y;
        }
      }
    }
    for (; r < r1; r += CS_ROWS) {
      const uint4 t = *reinterpret_cast<const uint4*>(X + r * ld + 8 * c8);
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16(w[q]);
        acc[2 * q] += f.x; acc[2 * q + 1] += f.y;
      }
    }
  } else if (8 * c8 < N) {
    for (long r = r0 + rl; r < r1; r += CS_ROWS)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (8 * c8 + j < N) acc[j] += __bfloat162float(X[r * ld + 8 * c8 + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[rl][threadIdx.x & 31][j] = acc[j];
  __syncthreads();
  if (rl == 0 && 8 * c8 < N) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < CS_ROWS; ++i) v += part[i][threadIdx.x][j];
      if (8 * c8 + j < N) atomicAdd(out + 8 * c8 + j, v);
    }
  }
}

struct PatchGeom { int B, C, H, W, p, Ho, Wo, F; };

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }

// One thread per PAIR of output features (consecutive in the fast source dimension for every mode:
// c_major+NCHW -> px pairs; pos-major+NHWC -> c pairs), so both reads and writes are coalesced.
template <typename T>
__global__ void patch_gather_kernel(const T* __restrict__ src, int src_nchw, int c_major, PatchGeom g,
                                    bf16* __restrict__ dst) {
  const long total = (long)g.B * g.Ho * g.Wo * (g.F / 2);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int f = (int)(i % (g.F / 2)) * 2;
    const long row = i / (g.F / 2);
    const int bx = (int)(row % g.Wo);
    const long t = row / g.Wo;
    const int by = (int)(t % g.Ho);
    const int b = (int)(t / g.Ho);
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ff = f + e;
      int c, py, px;
      if (c_major) { c = ff / (g.p * g.p); const int r = ff - c * g.p * g.p; py = r / g.p; px = r - py * g.p; }
      else { const int s = ff / g.C; c = ff - s * g.C; py = s / g.p; px = s - py * g.p; }
      const int y = by * g.p + py, x = bx * g.p + px;
      const long off = src_nchw ? (((long)b * g.C + c) * g.H + y) * g.W + x
                                : (((long)b * g.H + y) * g.W + x) * g.C + c;
      v[e] = ldf<T>(src + off);
    }
    *reinterpret_cast<uint32_t*>(dst + row * g.F + f) = pack_bf16(v[0], v[1]);
  }
}

// Vectorised pos-major gather / scatter for NHWC tensors (feature = (py*p + px)*C + c: the `patchify` order, and — with the
// conv weight permuted to [C_out, p, p, C] — also the order used for the k = s convolutions over NHWC token maps).  The
// per-vector source offset inside a patch does not depend on the patch, so it is tabulated once per CTA; per patch only
// two 32-bit divisions remain (per thread per ROW, not per element).  16-byte loads, 8/16-byte stores, fully coalesced.
template <typename T, int V>
__global__ void __launch_bounds__(256)
patch_gather_nhwc_vec_kernel(const T* __restrict__ src, PatchGeom g, bf16* __restrict__ dst, int rows_total,
                             int rows_per_cta) {
  extern __shared__ int s_delta[];
  const int cv = g.C / V, vpr = g.p * g.p * cv;
  for (int v = threadIdx.x; v < vpr; v += blockDim.x) {
    const int sidx = v / cv, vc = v - sidx * cv;
    const int py = sidx / g.p, px = sidx - py * g.p;
    s_delta[v] = (py * g.W + px) * g.C + vc * V;
  }
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows_total, r0 + rows_per_cta);
  for (int row = r0; row < r1; ++row) {
    const int q = row / g.Wo, bx = row - q * g.Wo;
    const int b = q / g.Ho, by = q - b * g.Ho;
    const T* sp = src + (((long)b * g.H + (long)by * g.p) * g.W + (long)bx * g.p) * g.C;
    bf16* dp = dst + (long)row * g.F;
    for (int v = threadIdx.x; v < vpr; v += blockDim.x) {
      if (V == 8) {  // bf16 source: a 16-byte copy
        *reinterpret_cast<uint4*>(dp + v * 8) = __ldg(reinterpret_cast<const uint4*>(sp + s_delta[v]));
      } else {       // f32 source: 4 floats -> 4 bf16
        const float4 x = __ldg(reinterpret_cast<const float4*>(sp + s_delta[v]));
        *reinterpret_cast<uint2*>(dp + v * 4) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
      }
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
patch_scatter_nhwc_vec_kernel(const T* __restrict__ dA, PatchGeom g, float* __restrict__ dx, int accumulate,
                              int rows_total, int rows_per_cta) {
  extern __shared__ int s_delta[];
  const int cv = g.C / V, vpr = g.p * g.p * cv;
  for (int v = threadIdx.x; v < vpr; v += blockDim.x) {
    const int sidx = v / cv, vc = v - sidx * cv;
    const int py = sidx / g.p, px = sidx - py * g.p;
    s_delta[v] = (py * g.W + px) * g.C + vc * V;
  }
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows_total, r0 + rows_per_cta);
  for (int row = r0; row < r1; ++row) {
    const int q = row / g.Wo, bx = row - q * g.Wo;
    const int b = q / g.Ho, by = q - b * g.Ho;
    float* op = dx + (((long)b * g.H + (long)by * g.p) * g.W + (long)bx * g.p) * g.C;
    const T* ip = dA + (long)row * g.F;
    for (int v = threadIdx.x; v < vpr; v += blockDim.x) {
      float x[V];
      if (V == 8) {
        const uint4 r = __ldg(reinterpret_cast<const uint4*>(ip + v * 8));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 f = unpack_bf16(w[k]); x[2 * k] = f.x; x[2 * k + 1] = f.y; }
      } else {
        const float4 r = __ldg(reinterpret_cast<const float4*>(ip + v * 4));
        x[0] = r.x; x[1] = r.y; x[2] = r.z; x[3] = r.w;
      }
      float4* o = reinterpret_cast<float4*>(op + s_delta[v]);
#pragma unroll
      for (int k = 0; k < V / 4; ++k) {
        float4 t = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
        if (accumulate) { const float4 old = o[k]; t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w; }
        o[k] = t;  // patches do not overlap: no atomics needed
      }
    }
  }
}

// rows per CTA so that the grid is a few CTAs per SM
inline int patch_rows_per_cta(long rows) {
  const long ctas = (long)(vtb_num_sms() > 0 ? vtb_num_sms() : 148) * 8;
  long r = (rows + ctas - 1) / ctas;
  return (int)(r < 1 ? 1 : r);
}

template <typename T>
__global__ void patch_scatter_kernel(const T* __restrict__ dA, int c_major, PatchGeom g,
                                     float* __restrict__ dx, int accumulate, int dst_nchw) {
  const long total = (long)g.B * g.Ho * g.Wo * g.F;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ff = (int)(i % g.F);
    const long row = i / g.F;
    const int bx = (int)(row % g.Wo);
    const long t = row / g.Wo;
    const int by = (int)(t % g.Ho);
    const int b = (int)(t / g.Ho);
    int c, py, px;
    if (c_major) { c = ff / (g.p * g.p); const int r = ff - c * g.p * g.p; py = r / g.p; px = r - py * g.p; }
    else { const int s = ff / g.C; c = ff - s * g.C; py = s / g.p; px = s - py * g.p; }
    const int y = by * g.p + py, x = bx * g.p + px;
    const long off = dst_nchw ? (((long)b * g.C + c) * g.H + y) * g.W + x
                              : (((long)b * g.H + y) * g.W + x) * g.C + c;
    const float v = ldf<T>(dA + i);
    dx[off] = accumulate ? dx[off] + v : v;  // patches do not overlap: no atomics needed
  }
}

__global__ void vit_assemble_kernel(const float* __restrict__ tok, const float* __restrict__ cls,
                                    const float* __restrict__ pos, int B, int n, int D, float* __restrict__ x) {
  const int D4 = D >> 2;
  const long total = (long)B * (n + 1) * D4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % D4);
    const long r = i / D4;
    const int t = (int)(r % (n + 1));
    const long b = r / (n + 1);
    const float4 p = reinterpret_cast<const float4*>(pos)[(long)t * D4 + c];
    const float4 s = (t == 0) ? reinterpret_cast<const float4*>(cls)[c]
                              : reinterpret_cast<const float4*>(tok)[(b * n + (t - 1)) * D4 + c];
    reinterpret_cast<float4*>(x)[i] = make_float4(s.x + p.x, s.y + p.y, s.z + p.z, s.w + p.w);
  }
}

// dst[b, w, h, :] = src[b, h, w, :] in 8-byte units (4 bf16 or 2 f32)
__global__ void transpose_hw_kernel(const uint2* __restrict__ src, uint2* __restrict__ dst, int B, int H, int W,
                                    int C8) {
  const long total = (long)B * H * W * C8;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long t = i / C8;
    const int h = (int)(t % H); t /= H;
    const int w = (int)(t % W);
    const long b = t / W;
    dst[i] = src[((b * H + h) * W + w) * C8 + c];
  }
}

// y[b,y,x,c] = x[b,y,x,c] + sum_{ky,kx} w[c,ky,kx] x[b,y+ky-1,x+kx-1,c]; one thread per float4 of channels
__global__ void dwconv3x3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int B, int H,
                                     int W, int C, float* __restrict__ y) {
  const int C4 = C >> 2;
  const long total = (long)B * H * W * C4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long t = i / C4;
    const int px = (int)(t % W); t /= W;
    const int py = (int)(t % H);
    const long b = t / H;
    float4 acc = reinterpret_cast<const float4*>(x)[i];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = py + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = px + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float4 v = reinterpret_cast<const float4*>(x)[((b * H + yy) * W + xx) * C4 + c4];
        const int c = c4 * 4;
        acc.x += __ldg(w + (c + 0) * 9 + ky * 3 + kx) * v.x;
        acc.y += __ldg(w + (c + 1) * 9 + ky * 3 + kx) * v.y;
        acc.z += __ldg(w + (c + 2) * 9 + ky * 3 + kx) * v.z;
        acc.w += __ldg(w + (c + 3) * 9 + ky * 3 + kx) * v.w;
      }
    }
    reinterpret_cast<float4*>(y)[i] = acc;
  }
}

// dx = dy + sum_k w[c,k] dy[shifted by -k]; dw[c,k] += sum dy[b,y,x,c] x[b,y+ky-1,x+kx-1,c].
// blockDim = (C4 lanes, PIX pixels): each thread keeps 9x4 dw partials over its pixels, CTA-reduces in smem.
constexpr int DW_PIX = 8;
__global__ void dwconv3x3_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ dy, int B, int H, int W, int C,
                                     float* __restrict__ dx, float* __restrict__ dw, long pix_per_block) {
  extern __shared__ float s_dw[];  // [C][9]
  const int C4 = C >> 2;
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < C * 9; i += blockDim.x * blockDim.y) s_dw[i] = 0.f;
  __syncthreads();
  const long npix = (long)B * H * W;
  const long p0 = (long)blockIdx.x * pix_per_block, p1 = min(npix, p0 + pix_per_block);
  for (int c4 = threadIdx.x; c4 < C4; c4 += blockDim.x) {
    float part[9][4];
#pragma unroll
    for (int k = 0; k < 9; ++k) part[k][0] = part[k][1] = part[k][2] = part[k][3] = 0.f;
    const int c = c4 * 4;
    for (long p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
      const int px = (int)(p % W);
      const long t = p / W;
      const int py = (int)(t % H);
      const long b = t / H;
      const float4 g = reinterpret_cast<const float4*>(dy)[p * C4 + c4];
      float4 acc = g;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int k = ky * 3 + kx;
          // forward tap: y[p] += w[k] x[p + (ky-1, kx-1)]  => dw[k] += dy[p] x[p + off]; dx[p] += w[k] dy[p - off]
          const int yy = py + ky - 1, xx = px + kx - 1;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const float4 v = reinterpret_cast<const float4*>(x)[((b * H + yy) * W + xx) * C4 + c4];
            part[k][0] += g.x * v.x; part[k][1] += g.y * v.y; part[k][2] += g.z * v.z; part[k][3] += g.w * v.w;
          }
          const int y2 = py - (ky - 1), x2 = px - (kx - 1);
          if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
            const float4 d = reinterpret_cast<const float4*>(dy)[((b * H + y2) * W + x2) * C4 + c4];
            acc.x += __ldg(w + (c + 0) * 9 + k) * d.x;
            acc.y += __ldg(w + (c + 1) * 9 + k) * d.y;
            acc.z += __ldg(w + (c + 2) * 9 + k) * d.z;
            acc.w += __ldg(w + (c + 3) * 9 + k) * d.w;
          }
        }
      }
      reinterpret_cast<float4*>(dx)[p * C4 + c4] = acc;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(&s_dw[(c + e) * 9 + k], part[k][e]);
  }
  __syncthreads();
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < C * 9; i += blockDim.x * blockDim.y)
    if (s_dw[i] != 0.f) atomicAdd(dw + i, s_dw[i]);
}

__global__ void fill_rows_kernel(float* __restrict__ x, long stride, int groups, int cols,
                                 const float* __restrict__ a, const float* __restrict__ b) {
  const long total = (long)groups * cols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const long gidx = i / cols;
    x[gidx * stride + c] = a[c] + (b ? b[c] : 0.f);
  }
}

// out[r*cols + c] += sum_g x[g*group_stride + r*cols + c]
__global__ void rowgroup_sum_kernel(const float* __restrict__ x, long group_stride, int groups, int rows,
                                    int cols, float* __restrict__ out) {
  const long total = (long)rows * cols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int gi = 0; gi < groups; ++gi) acc += x[(long)gi * group_stride + i];
    out[i] += acc;
  }
}

__global__ void mean_rows_fwd_kernel(const float* __restrict__ x, int groups, int n, int cols,
                                     float* __restrict__ out) {
  const long total = (long)groups * cols;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const long gi = i / cols;
    float acc = 0.f;
    for (int r = 0; r < n; ++r) acc += x[(gi * n + r) * cols + c];
    out[i] = acc / n;
  }
}
__global__ void mean_rows_bwd_kernel(const float* __restrict__ dy, int groups, int n, int cols,
                                     float* __restrict__ dx) {
  const long total = (long)groups * n * cols;
  const float inv = 1.f / n;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const long gi = i / ((long)n * cols);
    dx[i] = dy[gi * cols + c] * inv;
  }
}

int patch_geom(const char* who, int B, int C, int H, int W, int p, PatchGeom* g) {
  VTB_CHECK(B > 0 && C > 0 && H > 0 && W > 0 && p > 0 && H % p == 0 && W % p == 0, -1,
            "%s: bad geometry B=%d C=%d H=%d W=%d p=%d", who, B, C, H, W, p);
  g->B = B; g->C = C; g->H = H; g->W = W; g->p = p; g->Ho = H / p; g->Wo = W / p; g->F = p * p * C;
  return 0;
}

}  // namespace

extern "C" int vtb_cast_f32_bf16(const float* src, void* dst, int64_t n, vtb_stream_t s) {
  VTB_CHECK(src && dst && n >= 0, -1, "vtb_cast_f32_bf16: bad args");
  if (n == 0) return 0;
  VTB_CHECK(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, -1, "vtb_cast_f32_bf16: alignment");
  cast_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)s>>>(src, (bf16*)dst, n);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_cast_f32_bf16_2d(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows,
                                    int32_t cols, vtb_stream_t s) {
  VTB_CHECK(src && dst && rows > 0 && cols > 0, -1, "vtb_cast_f32_bf16_2d: bad args");
  cast2d_kernel<<<grid_for(rows * cols, 256), 256, 0, (cudaStream_t)s>>>(src, lds, (bf16*)dst, ldd, rows, cols);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_scale_cast_bf16(const float* src, const float* row_scale, int32_t rows_per_scale,
                                   int64_t rows, int32_t cols, void* dst, vtb_stream_t s) {
  VTB_CHECK(src && dst && rows > 0 && cols > 0 && cols % 4 == 0, -1, "vtb_scale_cast_bf16: bad args");
  VTB_CHECK(!row_scale || rows_per_scale > 0, -1, "vtb_scale_cast_bf16: rows_per_scale");
  VTB_CHECK(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, -1, "vtb_scale_cast_bf16: alignment");
  scale_cast_kernel<<<grid_for(rows * (cols / 4), 256, 2), 256, 0, (cudaStream_t)s>>>(
      src, row_scale, rows_per_scale, rows, cols, (bf16*)dst);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_scale_cast_colsum_bf16(const float* src, const float* row_scale, int32_t rows_per_scale,
                                          int64_t rows, int32_t cols, void* dst, float* colsum, vtb_stream_t s) {
  VTB_CHECK(src && dst && colsum && rows > 0 && cols > 0 && cols % 4 == 0, -1, "vtb_scale_cast_colsum_bf16: bad args");
  VTB_CHECK(!row_scale || rows_per_scale > 0, -1, "vtb_scale_cast_colsum_bf16: rows_per_scale");
  VTB_CHECK(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, -1, "vtb_scale_cast_colsum_bf16: alignment");
  const int gx = (cols / 4 + 31) / 32;
  int gy = (6 * (vtb_num_sms() > 0 ? vtb_num_sms() : 148) + gx - 1) / gx;
  long rpb = (rows + gy - 1) / gy;
  if (rpb < 64) rpb = 64;
  gy = (int)((rows + rpb - 1) / rpb);
  scale_cast_colsum_kernel<<<dim3(gx, gy), dim3(32, 8), 0, (cudaStream_t)s>>>(src, row_scale, rows_per_scale, rows, cols,
                                                                            (bf16*)dst, colsum, rpb);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_silu_fwd(const float* x, float* y, int64_t n, vtb_stream_t s) {
  VTB_CHECK(x && y && n > 0, -1, "vtb_silu_fwd: bad args");
  silu_fwd_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)s>>>(x, y, n);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_silu_bwd(const float* x, const float* dy, float* dx, int64_t n, vtb_stream_t s) {
  VTB_CHECK(x && dy && dx && n > 0, -1, "vtb_silu_bwd: bad args");
  silu_bwd_kernel<<<grid_for(n, 256, 4), 256, 0, (cudaStream_t)s>>>(x, dy, dx, n);
  VTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------- element dropout
// out = resid + row_scale[row] * (keep ? x * scale : 0): nn.Dropout applied to an activation (layer.py:194, vit.py:57-61,
// pvt.py:141) with the keep mask drawn by torch's own generator (vtb200.blocks.make_dropout_keep), its adjoint (same
// call on the gradient), and — with resid / row_scale — the DropPath + residual step of a ViT branch whose output
// Linear can no longer fuse them once a per-element mask sits in between.  No reference config uses p > 0.
namespace {
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, const uint8_t* __restrict__ keep, float scale,
                               const float* __restrict__ resid, const float* __restrict__ row_scale,
                               long elems_per_scale, long n, T* __restrict__ out) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float v = keep[i] ? (float)x[i] * scale : 0.f;
    if (row_scale) v *= __ldg(row_scale + i / elems_per_scale);
    if (resid) v += resid[i];
    out[i] = (T)v;
  }
}
}  // namespace

extern "C" int vtb_dropout(const void* x, const uint8_t* keep, float scale, int64_t n, int32_t is_f32,
                           const float* resid, const float* row_scale, int64_t elems_per_scale, void* out,
                           vtb_stream_t s) {
  VTB_CHECK(x && keep && out && n > 0, -1, "vtb_dropout: bad args");
  VTB_CHECK(!row_scale || elems_per_scale > 0, -1, "vtb_dropout: elems_per_scale");
  VTB_CHECK(!resid || is_f32, -1, "vtb_dropout: the residual form is fp32 only");
  if (is_f32)
    dropout_kernel<float><<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)s>>>((const float*)x, keep, scale, resid, row_scale,
                                                                            elems_per_scale, n, (float*)out);
  else
    dropout_kernel<bf16><<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)s>>>((const bf16*)x, keep, scale, resid, row_scale,
                                                                           elems_per_scale, n, (bf16*)out);
  VTB_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------- validation mode
// fp32 -> three bf16 column blocks of one GEMM operand row, so that ONE tcgen05 GEMM with K' = 3K accumulates
//   a_hi b_hi + a_lo b_hi + a_hi b_lo   (hi = bf16(x), lo = bf16(x - hi): ~16 mantissa bits per operand, fp32 accumulation)
// A side: [hi | lo | hi], B side: [hi | hi | lo].
namespace {
__global__ void split3_kernel(const float* __restrict__ src, long lds, long rows, int cols, int b_side,
                              bf16* __restrict__ dst) {
  const long total = rows * cols;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long r = e / cols;
    const int c = (int)(e - r * cols);
    const float x = src[r * lds + c];
    const bf16 hi = __float2bfloat16_rn(x);
    const bf16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    bf16* d = dst + r * 3 * cols + c;
    d[0] = hi;
    d[cols] = b_side ? hi : lo;
    d[2 * cols] = b_side ? lo : hi;
  }
}
__global__ void silu_exact_kernel(const float* __restrict__ x, float* __restrict__ y, long n) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const float v = x[e];
    y[e] = v / (1.f + expf(-v));
  }
}
template <typename T>
__global__ void patch_gather_f32_kernel(const T* __restrict__ src, int src_nchw, int c_major, PatchGeom g,
                                        float* __restrict__ dst) {
  const long total = (long)g.B * g.Ho * g.Wo * g.F;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ff = (int)(i % g.F);
    const long row = i / g.F;
    const int bx = (int)(row % g.Wo);
    const long t = row / g.Wo;
    const int by = (int)(t % g.Ho);
    const int b = (int)(t / g.Ho);
    int c, py, px;
    if (c_major) { c = ff / (g.p * g.p); const int r = ff - c * g.p * g.p; py = r / g.p; px = r - py * g.p; }
    else { const int s = ff / g.C; c = ff - s * g.C; py = s / g.p; px = s - py * g.p; }
    const int y = by * g.p + py, x = bx * g.p + px;
    const long off = src_nchw ? (((long)b * g.C + c) * g.H + y) * g.W + x : (((long)b * g.H + y) * g.W + x) * g.C + c;
    dst[i] = src[off];
  }
}
}  // namespace

extern "C" int vtb_split3_bf16(const float* src, int64_t lds, int64_t rows, int32_t cols, int32_t b_side, void* dst,
                               vtb_stream_t s) {
  VTB_CHECK(src && dst && rows > 0 && cols > 0 && lds >= cols, -1, "vtb_split3_bf16: bad args");
  split3_kernel<<<grid_for(rows * cols, 256), 256, 0, (cudaStream_t)s>>>(src, lds, rows, cols, b_side, (bf16*)dst);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_silu_fwd_exact(const float* x, float* y, int64_t n, vtb_stream_t s) {
  VTB_CHECK(x && y && n > 0, -1, "vtb_silu_fwd_exact: bad args");
  silu_exact_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)s>>>(x, y, n);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_patch_gather_f32(const float* src, int32_t src_nchw, int32_t c_major, int32_t B, int32_t C, int32_t H,
                                    int32_t W, int32_t p, float* dst, vtb_stream_t s) {
  PatchGeom g;
  int rc = patch_geom("vtb_patch_gather_f32", B, C, H, W, p, &g);
  if (rc) return rc;
  VTB_CHECK(src && dst, -1, "vtb_patch_gather_f32: null pointer");
  const long total = (long)B * g.Ho * g.Wo * g.F;
  patch_gather_f32_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>(src, src_nchw, c_major, g, dst);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_colsum_bf16(const void* X, int64_t M, int32_t N, int32_t ld, float* out, vtb_stream_t s) {
  VTB_CHECK(X && out && M > 0 && N > 0, -1, "vtb_colsum_bf16: bad args");
  VTB_CHECK(ld % 8 == 0 && ((uintptr_t)X & 15) == 0, -1, "vtb_colsum_bf16: rows must be 16-byte aligned");
  const int gx = ((N + 7) / 8 + 31) / 32;
  int gy = (8 * (vtb_num_sms() > 0 ? vtb_num_sms() : 148) + gx - 1) / gx;
  long rpb = (M + gy - 1) / gy;
  if (rpb < 128) rpb = 128;
  gy = (int)((M + rpb - 1) / rpb);
  colsum_kernel<<<dim3(gx, gy), 256, 0, (cudaStream_t)s>>>((const bf16*)X, M, N, ld, out, rpb);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_patch_gather(const void* src, int32_t src_bf16, int32_t src_nchw, int32_t c_major,
                                int32_t B, int32_t C, int32_t H, int32_t W, int32_t p, void* dst,
                                vtb_stream_t s) {
  PatchGeom g;
  int rc = patch_geom("vtb_patch_gather", B, C, H, W, p, &g);
  if (rc) return rc;
  VTB_CHECK(src && dst, -1, "vtb_patch_gather: null pointer");
  VTB_CHECK(g.F % 2 == 0, -1, "vtb_patch_gather: feature count must be even");
  {  // vectorised path: pos-major features of an NHWC source whose channel rows split into 16-byte vectors
    const int V = src_bf16 ? 8 : 4;
    const long rows = (long)B * g.Ho * g.Wo;
    const size_t tab = (size_t)p * p * (C / V) * sizeof(int);
    if (!c_major && !src_nchw && C % V == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0 &&
        rows < (1L << 31) && tab <= 48 * 1024) {
      const int rpc = patch_rows_per_cta(rows);
      const int grid = (int)((rows + rpc - 1) / rpc);
      if (src_bf16)
        patch_gather_nhwc_vec_kernel<bf16, 8><<<grid, 256, tab, (cudaStream_t)s>>>((const bf16*)src, g, (bf16*)dst, (int)rows, rpc);
      else
        patch_gather_nhwc_vec_kernel<float, 4><<<grid, 256, tab, (cudaStream_t)s>>>((const float*)src, g, (bf16*)dst, (int)rows, rpc);
      VTB_LAUNCH_CHECK();
      return 0;
    }
  }
  const long total = (long)B * g.Ho * g.Wo * (g.F / 2);
  if (src_bf16)
    patch_gather_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>((const bf16*)src, src_nchw, c_major, g, (bf16*)dst);
  else
    patch_gather_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>((const float*)src, src_nchw, c_major, g, (bf16*)dst);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_patch_scatter(const void* dA, int32_t dA_f32, int32_t c_major, int32_t B, int32_t C,
                                 int32_t H, int32_t W, int32_t p, float* dx, int32_t accumulate,
                                 int32_t dst_nchw, vtb_stream_t s) {
  PatchGeom g;
  int rc = patch_geom("vtb_patch_scatter", B, C, H, W, p, &g);
  if (rc) return rc;
  VTB_CHECK(dA && dx, -1, "vtb_patch_scatter: null pointer");
  {  // vectorised path (see vtb_patch_gather)
    const int V = dA_f32 ? 4 : 8;
    const long rows = (long)B * g.Ho * g.Wo;
    const size_t tab = (size_t)p * p * (C / V) * sizeof(int);
    if (!c_major && !dst_nchw && C % V == 0 && ((uintptr_t)dA & 15) == 0 && ((uintptr_t)dx & 15) == 0 &&
        rows < (1L << 31) && tab <= 48 * 1024) {
      const int rpc = patch_rows_per_cta(rows);
      const int grid = (int)((rows + rpc - 1) / rpc);
      if (dA_f32)
        patch_scatter_nhwc_vec_kernel<float, 4><<<grid, 256, tab, (cudaStream_t)s>>>((const float*)dA, g, dx, accumulate, (int)rows, rpc);
      else
        patch_scatter_nhwc_vec_kernel<bf16, 8><<<grid, 256, tab, (cudaStream_t)s>>>((const bf16*)dA, g, dx, accumulate, (int)rows, rpc);
      VTB_LAUNCH_CHECK();
      return 0;
    }
  }
  const long total = (long)B * g.Ho * g.Wo * g.F;
  if (dA_f32)
    patch_scatter_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>((const float*)dA, c_major, g, dx, accumulate, dst_nchw);
  else
    patch_scatter_kernel<bf16><<<grid_for(total, 256), 256, 0, (cudaStream_t)s>>>((const bf16*)dA, c_major, g, dx, accumulate, dst_nchw);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_vit_assemble_tokens(const float* tok, const float* cls, const float* pos, int32_t B,
                                       int32_t n, int32_t D, float* x, vtb_stream_t s) {
  VTB_CHECK(tok && cls && pos && x && B > 0 && n > 0 && D > 0 && D % 4 == 0, -1,
            "vtb_vit_assemble_tokens: bad args");
  vit_assemble_kernel<<<grid_for((long)B * (n + 1) * (D / 4), 256, 2), 256, 0, (cudaStream_t)s>>>(tok, cls, pos, B, n, D, x);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_transpose_hw(const void* src, void* dst, int32_t is_f32, int32_t B, int32_t H, int32_t W,
                                int32_t C, vtb_stream_t s) {
  VTB_CHECK(src && dst && B > 0 && H > 0 && W > 0 && C > 0, -1, "vtb_transpose_hw: bad args");
  const int per8 = is_f32 ? 2 : 4;
  VTB_CHECK(C % per8 == 0 && ((uintptr_t)src & 7) == 0 && ((uintptr_t)dst & 7) == 0, -1,
            "vtb_transpose_hw: rows must be multiples of 8 bytes");
  const int C8 = C / per8;
  transpose_hw_kernel<<<grid_for((long)B * H * W * C8, 256, 2), 256, 0, (cudaStream_t)s>>>(
      (const uint2*)src, (uint2*)dst, B, H, W, C8);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_dwconv3x3_fwd(const float* x, const float* w, int32_t B, int32_t H, int32_t W, int32_t C,
                                 float* y, vtb_stream_t s) {
  VTB_CHECK(x && w && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, -1, "vtb_dwconv3x3_fwd: bad args");
  dwconv3x3_fwd_kernel<<<grid_for((long)B * H * W * (C / 4), 256, 2), 256, 0, (cudaStream_t)s>>>(x, w, B, H, W, C, y);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_dwconv3x3_bwd(const float* x, const float* w, const float* dy, int32_t B, int32_t H,
                                 int32_t W, int32_t C, float* dx, float* dw, vtb_stream_t s) {
  VTB_CHECK(x && w && dy && dx && dw && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, -1,
            "vtb_dwconv3x3_bwd: bad args");
  const long npix = (long)B * H * W;
  int bx = C / 4 < 32 ? C / 4 : 32;
  dim3 block(bx, DW_PIX);
  long blocks = (long)(vtb_num_sms() > 0 ? vtb_num_sms() : 148) * 8;
  long ppb = (npix + blocks - 1) / blocks;
  if (ppb < DW_PIX) ppb = DW_PIX;
  blocks = (npix + ppb - 1) / ppb;
  dwconv3x3_bwd_kernel<<<(unsigned)blocks, block, (size_t)C * 9 * sizeof(float), (cudaStream_t)s>>>(
      x, w, dy, B, H, W, C, dx, dw, ppb);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_fill_rows(float* x, int64_t row_stride_groups, int32_t groups, int32_t cols,
                             const float* a, const float* b, vtb_stream_t s) {
  VTB_CHECK(x && a && groups > 0 && cols > 0, -1, "vtb_fill_rows: bad args");
  fill_rows_kernel<<<grid_for((long)groups * cols, 256), 256, 0, (cudaStream_t)s>>>(x, row_stride_groups, groups, cols, a, b);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_rowgroup_sum(const float* x, int64_t group_stride, int32_t groups, int32_t rows,
                                int32_t cols, float* out, vtb_stream_t s) {
  VTB_CHECK(x && out && groups > 0 && rows > 0 && cols > 0, -1, "vtb_rowgroup_sum: bad args");
  rowgroup_sum_kernel<<<grid_for((long)rows * cols, 128), 128, 0, (cudaStream_t)s>>>(x, group_stride, groups, rows, cols, out);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_mean_rows_fwd(const float* x, int32_t groups, int32_t n, int32_t cols, float* out,
                                 vtb_stream_t s) {
  VTB_CHECK(x && out && groups > 0 && n > 0 && cols > 0, -1, "vtb_mean_rows_fwd: bad args");
  mean_rows_fwd_kernel<<<grid_for((long)groups * cols, 128), 128, 0, (cudaStream_t)s>>>(x, groups, n, cols, out);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_mean_rows_bwd(const float* dy, int32_t groups, int32_t n, int32_t cols, float* dx,
                                 vtb_stream_t s) {
  VTB_CHECK(dy && dx && groups > 0 && n > 0 && cols > 0, -1, "vtb_mean_rows_bwd: bad args");
  mean_rows_bwd_kernel<<<grid_for((long)groups * n * cols, 256), 256, 0, (cudaStream_t)s>>>(dy, groups, n, cols, dx);
  VTB_LAUNCH_CHECK();
  return 0;
}
