// Fused DINO loss, forward + gradient in one launch (loss.py:119-142), and the row kernels of the
// DINO projection head (vit.py:206-262): GELU, L2 normalisation, weight-norm reparametrisation.
//
//   q_iq  = softmax((teacher[iq] - center) / t_teacher)            iq in {0, 1} (the two global crops)
//   loss  = mean over the pairs (iq, v != iq) and the images of   - sum_k q_iq[k] log_softmax(student[v] / t_student)[k]
//
// The reference evaluates 18 log-softmax / multiply / sum chains over [B, 65536] chunks (and autograd walks them back):
// ~70 passes over the logits.  Here one CTA owns one image: pass 1 streams its 2 teacher rows and n_crops student rows
// once for the softmax statistics (online max / sum in the log2 domain), pass 2 streams them again and, per column,
// forms q_0, q_1, every student probability, the dot products and the gradient row in registers:
//   loss_b        = sum_v [ cnt_v lse_v - (Qsum_v . s_v) ]      Qsum_v = sum_{iq != v} q_iq,  cnt_v = #{iq != v}
//   d loss / d student[v, b, k] = (cnt_v softmax(s_v)[k] - Qsum_v[k]) / (n_pairs B t_student)
// HBM-bound: (2 reads + 1 gradient write) of the logits.
#include "common.cuh"
#include "../../include/vtb200.h"
#include <math.h>

namespace {

constexpr int DL_THREADS = 512;
constexpr int DL_MAX_ROWS = 18;  // 2 teacher + up to 16 student crops
constexpr float DL_L2E = 1.4426950408889634f;
constexpr float DL_LN2 = 0.6931471805599453f;

__device__ __forceinline__ float dl_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (m, s) <- combine with (m2, s2): running max and sum of 2^(y - max)
__device__ __forceinline__ void dl_combine(float& m, float& s, float m2, float s2) {
  const float M = fmaxf(m, m2);
  const float e1 = (m == M) ? 1.f : dl_ex2(m - M);   // (-inf, 0) pairs of idle lanes must not produce inf - inf
  const float e2 = (m2 == M) ? 1.f : dl_ex2(m2 - M);
  s = s * e1 + s2 * e2;
  m = M;
}

__global__ void __launch_bounds__(DL_THREADS)
dino_loss_kernel(const float* __restrict__ student, const float* __restrict__ teacher,
                 const float* __restrict__ center, int n_crops, int batch, int dim, float inv_ts_l2, float inv_tt_l2,
                 float inv_ts, float coef, float loss_scale, float* __restrict__ loss,
                 float* __restrict__ dstudent) {
  __shared__ float s_m[DL_MAX_ROWS], s_z[DL_MAX_ROWS];
  __shared__ float s_red[2][DL_THREADS / 32];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = dim >> 2;
  const int rows = n_crops + 2;
  const float4* c4 = reinterpret_cast<const float4*>(center);

  // ---------------------------------------------------------------- pass 1: softmax statistics of every row
  for (int r = 0; r < rows; ++r) {
    const bool is_t = r < 2;
    const float4* src = is_t ? reinterpret_cast<const float4*>(teacher + ((long)r * batch + b) * dim)
                             : reinterpret_cast<const float4*>(student + ((long)(r - 2) * batch + b) * dim);
    const float sc = is_t ? inv_tt_l2 : inv_ts_l2;
    float m = -INFINITY, s = 0.f;
    for (int v = tid; v < nv; v += DL_THREADS) {
      float4 x = __ldg(src + v);
      if (is_t) {
        const float4 c = __ldg(c4 + v);
        x.x -= c.x; x.y -= c.y; x.z -= c.z; x.w -= c.w;
      }
      const float y0 = x.x * sc, y1 = x.y * sc, y2 = x.z * sc, y3 = x.w * sc;
      const float M = fmaxf(fmaxf(m, fmaxf(y0, y1)), fmaxf(y2, y3));
      s = s * dl_ex2(m - M) + (dl_ex2(y0 - M) + dl_ex2(y1 - M)) + (dl_ex2(y2 - M) + dl_ex2(y3 - M));
      m = M;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      dl_combine(m, s, m2, s2);
    }
    if (lane == 0) { s_red[0][warp] = m; s_red[1][warp] = s; }
    __syncthreads();
    if (warp == 0) {
      m = (lane < DL_THREADS / 32) ? s_red[0][lane] : -INFINITY;
      s = (lane < DL_THREADS / 32) ? s_red[1][lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        dl_combine(m, s, m2, s2);
      }
      if (lane == 0) { s_m[r] = m; s_z[r] = s; }
    }
    __syncthreads();
  }

  // ---------------------------------------------------------------- pass 2: probabilities, dot products, gradient
  const float mt0 = s_m[0], mt1 = s_m[1];
  const float izt0 = 1.f / s_z[0], izt1 = 1.f / s_z[1];
  const float4* t0 = reinterpret_cast<const float4*>(teacher + (long)b * dim);
  const float4* t1 = reinterpret_cast<const float4*>(teacher + ((long)batch + b) * dim);
  float dot = 0.f;  // sum_v Qsum_v . s_v (natural-log units of s)
  for (int v = tid; v < nv; v += DL_THREADS) {
    const float4 c = __ldg(c4 + v);
    const float4 a0 = __ldg(t0 + v), a1 = __ldg(t1 + v);
    float q0[4], q1[4];
    q0[0] = dl_ex2((a0.x - c.x) * inv_tt_l2 - mt0) * izt0; q0[1] = dl_ex2((a0.y - c.y) * inv_tt_l2 - mt0) * izt0;
    q0[2] = dl_ex2((a0.z - c.z) * inv_tt_l2 - mt0) * izt0; q0[3] = dl_ex2((a0.w - c.w) * inv_tt_l2 - mt0) * izt0;
    q1[0] = dl_ex2((a1.x - c.x) * inv_tt_l2 - mt1) * izt1; q1[1] = dl_ex2((a1.y - c.y) * inv_tt_l2 - mt1) * izt1;
    q1[2] = dl_ex2((a1.z - c.z) * inv_tt_l2 - mt1) * izt1; q1[3] = dl_ex2((a1.w - c.w) * inv_tt_l2 - mt1) * izt1;
    for (int cr = 0; cr < n_crops; ++cr) {
      const long row = ((long)cr * batch + b) * dim;
      const float4 sv = __ldg(reinterpret_cast<const float4*>(student + row) + v);
      const float ms = s_m[2 + cr], izs = 1.f / s_z[2 + cr];
      const float cnt = (cr < 2) ? 1.f : 2.f;
      const float x[4] = {sv.x, sv.y, sv.z, sv.w};
      float g[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float qs = (cr == 0) ? q1[e] : ((cr == 1) ? q0[e] : q0[e] + q1[e]);
        const float p = dl_ex2(x[e] * inv_ts_l2 - ms) * izs;
        dot += qs * (x[e] * inv_ts);
        g[e] = coef * (cnt * p - qs);
      }
      if (dstudent) reinterpret_cast<float4*>(dstudent + row)[v] = make_float4(g[0], g[1], g[2], g[3]);
    }
  }
  dot = warp_sum(dot);
  if (lane == 0) s_red[0][warp] = dot;
  __syncthreads();
  if (warp == 0) {
    dot = (lane < DL_THREADS / 32) ? s_red[0][lane] : 0.f;
    dot = warp_sum(dot);
    if (lane == 0) {
      float lb = -dot;
      for (int cr = 0; cr < n_crops; ++cr)
        lb += ((cr < 2) ? 1.f : 2.f) * (s_m[2 + cr] + log2f(s_z[2 + cr])) * DL_LN2;
      atomicAdd(loss, lb * loss_scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------ DINO head (vit.py:206-262)
// Row kernels, one warp per row.  l2norm: F.normalize(x, dim=-1) (vit.py:259) emitted as the bf16 GEMM operand of the
// weight-normed last layer; weight_norm: w = v * g / ||v|| per output row (nn.utils.weight_norm, vit.py:244-248) emitted
// as bf16 as well, so neither the normalised activations nor the 65 536 x 256 effective weight ever exist in fp32.
constexpr int RN_WARPS = 8;

template <bool WN>  // WN: scale by g[row] (weight norm); else plain L2 normalisation with eps clamp
__global__ void __launch_bounds__(RN_WARPS * 32)
rownorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g, long rows, int cols, float eps,
                   bf16* __restrict__ y, float* __restrict__ inv_out) {
  const long r = (long)blockIdx.x * RN_WARPS + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + r * cols;
  float ss = 0.f;
  for (int c = lane; c < cols; c += 32) ss = fmaf(xr[c], xr[c], ss);
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  if (lane == 0) inv_out[r] = inv;
  const float sc = WN ? inv * __ldg(g + r) : inv;
  for (int c = lane; c < cols; c += 32) y[r * cols + c] = __float2bfloat16(xr[c] * sc);
}

// l2norm bwd:      dx = inv (dy - y (y . dy)),  y = x inv
// weight_norm bwd: dg = (dW . v) inv;  dv = g inv (dW - v (dW . v) inv^2)
template <bool WN>
__global__ void __launch_bounds__(RN_WARPS * 32)
rownorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ g,
                   const float* __restrict__ inv_in, long rows, int cols, float* __restrict__ dx,
                   float* __restrict__ dg) {
  const long r = (long)blockIdx.x * RN_WARPS + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + r * cols;
  const float* dr = dy + r * cols;
  float dot = 0.f;
  for (int c = lane; c < cols; c += 32) dot = fmaf(xr[c], dr[c], dot);
  dot = warp_sum(dot);
  const float inv = inv_in[r];
  const float sc = WN ? inv * __ldg(g + r) : inv;
  const float k = dot * inv * inv;
  if (WN && dg && lane == 0) dg[r] = dot * inv;
  for (int c = lane; c < cols; c += 32) dx[r * cols + c] = sc * (dr[c] - xr[c] * k);
}

__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, bf16* __restrict__ yb, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float o = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));  // nn.GELU() (exact, vit.py:228)
    if (y) y[i] = o;
    if (yb) yb[i] = __float2bfloat16(o);
  }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
    dx[i] = dy[i] * (cdf + v * pdf);
  }
}

}  // namespace

extern "C" int vtb_l2norm_fwd(const float* x, int64_t rows, int32_t cols, float eps, void* y_bf16, float* inv,
                              vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(x && y_bf16 && inv && rows >= 0 && cols > 0, -1, "vtb_l2norm_fwd: bad arguments");
  if (rows == 0) return 0;
  rownorm_fwd_kernel<false><<<(unsigned)((rows + RN_WARPS - 1) / RN_WARPS), RN_WARPS * 32, 0, stream>>>(
      x, nullptr, rows, cols, eps, static_cast<bf16*>(y_bf16), inv);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_l2norm_bwd(const float* dy, const float* x, const float* inv, int64_t rows, int32_t cols, float* dx,
                              vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(dy && x && inv && dx && rows >= 0 && cols > 0, -1, "vtb_l2norm_bwd: bad arguments");
  if (rows == 0) return 0;
  rownorm_bwd_kernel<false><<<(unsigned)((rows + RN_WARPS - 1) / RN_WARPS), RN_WARPS * 32, 0, stream>>>(
      dy, x, nullptr, inv, rows, cols, dx, nullptr);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_weight_norm_fwd(const float* v, const float* g, int64_t rows, int32_t cols, void* w_bf16, float* inv,
                                   vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(v && g && w_bf16 && inv && rows >= 0 && cols > 0, -1, "vtb_weight_norm_fwd: bad arguments");
  if (rows == 0) return 0;
  rownorm_fwd_kernel<true><<<(unsigned)((rows + RN_WARPS - 1) / RN_WARPS), RN_WARPS * 32, 0, stream>>>(
      v, g, rows, cols, 0.f, static_cast<bf16*>(w_bf16), inv);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* inv, int64_t rows,
                                   int32_t cols, float* dv, float* dg, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(dw && v && g && inv && dv && rows >= 0 && cols > 0, -1, "vtb_weight_norm_bwd: bad arguments");
  if (rows == 0) return 0;
  rownorm_bwd_kernel<true><<<(unsigned)((rows + RN_WARPS - 1) / RN_WARPS), RN_WARPS * 32, 0, stream>>>(
      dw, v, g, inv, rows, cols, dv, dg);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_gelu_fwd(const float* x, float* y, void* y_bf16, int64_t n, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(x && (y || y_bf16) && n >= 0, -1, "vtb_gelu_fwd: bad arguments");
  if (n == 0) return 0;
  const long blocks = (n + 255) / 256;
  gelu_fwd_kernel<<<(unsigned)(blocks < 4736 ? blocks : 4736), 256, 0, stream>>>(x, y, static_cast<bf16*>(y_bf16), n);
  VTB_LAUNCH_CHECK();
  return 0;
}
extern "C" int vtb_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(x && dy && dx && n >= 0, -1, "vtb_gelu_bwd: bad arguments");
  if (n == 0) return 0;
  const long blocks = (n + 255) / 256;
  gelu_bwd_kernel<<<(unsigned)(blocks < 4736 ? blocks : 4736), 256, 0, stream>>>(x, dy, dx, n);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_dino_loss(const float* student, const float* teacher, const float* center, int32_t n_crops,
                             int32_t batch, int32_t dim, float t_student, float t_teacher, float* loss,
                             float* dstudent, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(student && teacher && center && loss, -1, "vtb_dino_loss: null pointer");
  VTB_CHECK(n_crops >= 2 && n_crops + 2 <= DL_MAX_ROWS, -1, "vtb_dino_loss: n_crops=%d must be in [2, %d]", n_crops,
            DL_MAX_ROWS - 2);
  VTB_CHECK(batch > 0 && dim > 0 && dim % 4 == 0, -1, "vtb_dino_loss: bad shape batch=%d dim=%d (dim %% 4)", batch, dim);
  VTB_CHECK(t_student > 0.f && t_teacher > 0.f, -1, "vtb_dino_loss: temperatures must be positive");
  VTB_CHECK((((uintptr_t)student | (uintptr_t)teacher | (uintptr_t)center | (uintptr_t)dstudent) & 15) == 0, -1,
            "vtb_dino_loss: 16-byte aligned rows expected");
  const int n_pairs = 2 * n_crops - 2;
  const float coef = 1.f / ((float)n_pairs * (float)batch * t_student);
  dino_loss_kernel<<<batch, DL_THREADS, 0, stream>>>(student, teacher, center, n_crops, batch, dim,
                                                     DL_L2E / t_student, DL_L2E / t_teacher, 1.f / t_student, coef,
                                                     1.f / ((float)n_pairs * (float)batch), loss, dstudent);
  VTB_LAUNCH_CHECK();
  return 0;
}
