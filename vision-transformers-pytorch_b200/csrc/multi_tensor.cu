// Multi-tensor step-side kernels: the per-parameter loops of the reference's training step —
// autocast weight casts (train.py:273), EMA `accumulate` (train_util.py:70-84, train_dino.py:257-261), clip_grad_norm_
// (train.py:294), adaptive_grad_clip (optimizer.py:12-26), AdamW (config/*.conf `type: adamw`) — each as ONE launch over
// a tensor list, and the fused MixLoss + accuracy row kernel (loss.py:53-86, train_util.py:53-67).
//
// All HBM-bound streams.  A list travels in the kernel parameters (pointer table + chunk prefix, <= 13 KB): no
// device-side table, nothing allocated.  One CTA streams one VTB_MT_CHUNK-element chunk of one tensor; the tensor of a
// CTA is found by a binary search over the chunk prefix (8 steps, constant bank).  This file is compiled WITHOUT
// --use_fast_math: the optimizer arithmetic keeps IEEE division / sqrt and denormals, like the ATen kernels it replaces.
#include "common.cuh"
#include "../../include/vtb200.h"
#include <math.h>

namespace {

constexpr int MT_MAX = VTB_MT_MAX_TENSORS;
constexpr int MT_CHUNK = VTB_MT_CHUNK;
constexpr int MT_THREADS = 256;

template <int D>
struct MtPack {
  void* ptr[D][MT_MAX];
  long numel[MT_MAX];
  int start[MT_MAX + 1];  // first CTA of tensor t (chunks, or AGC units); start[n] = grid size
  int n;
};

template <int D>
__device__ __forceinline__ int mt_find(const MtPack<D>& pk, int blk) {
  int lo = 0, hi = pk.n;  // start[lo] <= blk < start[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pk.start[mid] <= blk) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();  // red may still be read from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < nw; ++w) s += red[w];  // fixed order: deterministic
  return s;
}

// ------------------------------------------------------------------------------------------------ cast / ema / scale
// Elementwise bodies over one chunk: F(i4) handles one aligned group of 4 elements, S(i) a single element.
template <typename F4, typename F1>
__device__ __forceinline__ void mt_stream(long cnt, bool vec, F4 f4, F1 f1) {
  if (vec) {
    const int n4 = (int)(cnt >> 2);
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += MT_THREADS) f4(i);
    for (long i = ((long)n4 << 2) + threadIdx.x; i < cnt; i += MT_THREADS) f1(i);
  } else {
    for (long i = threadIdx.x; i < cnt; i += MT_THREADS) f1(i);
  }
}

__device__ __forceinline__ bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
__device__ __forceinline__ bool aligned8(const void* p) { return ((uintptr_t)p & 7) == 0; }

__global__ void __launch_bounds__(MT_THREADS) mt_cast_kernel(const __grid_constant__ MtPack<2> pk) {
  const int t = mt_find(pk, blockIdx.x);
  const long base = (long)(blockIdx.x - pk.start[t]) * MT_CHUNK;
  const long cnt = min((long)MT_CHUNK, pk.numel[t] - base);
  const float* src = static_cast<const float*>(pk.ptr[0][t]) + base;
  bf16* dst = static_cast<bf16*>(pk.ptr[1][t]) + base;
  mt_stream(cnt, aligned16(src) && aligned8(dst),
            [&](int i) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
              reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            },
            [&](long i) { dst[i] = __float2bfloat16(src[i]); });
}

__global__ void __launch_bounds__(MT_THREADS) mt_ema_kernel(const __grid_constant__ MtPack<2> pk, float decay,
                                                            float alpha) {
  const int t = mt_find(pk, blockIdx.x);
  const long base = (long)(blockIdx.x - pk.start[t]) * MT_CHUNK;
  const long cnt = min((long)MT_CHUNK, pk.numel[t] - base);
  float* dst = static_cast<float*>(pk.ptr[0][t]) + base;
  const float* src = static_cast<const float*>(pk.ptr[1][t]) + base;
  // mul_(decay) then add_(src, alpha = 1 - decay): the product is rounded before the (fused) multiply-add
  auto ema = [&](float d, float s) { return fmaf(s, alpha, __fmul_rn(d, decay)); };
  mt_stream(cnt, aligned16(src) && aligned16(dst),
            [&](int i) {
              const float4 s = __ldg(reinterpret_cast<const float4*>(src) + i);
              float4 d = reinterpret_cast<float4*>(dst)[i];
              d.x = ema(d.x, s.x); d.y = ema(d.y, s.y); d.z = ema(d.z, s.z); d.w = ema(d.w, s.w);
              reinterpret_cast<float4*>(dst)[i] = d;
            },
            [&](long i) { dst[i] = ema(dst[i], src[i]); });
}

__global__ void __launch_bounds__(MT_THREADS) mt_scale_kernel(const __grid_constant__ MtPack<1> pk,
                                                              const float* __restrict__ scale) {
  const float sc = __ldg(scale);
  if (sc == 1.f) return;
  const int t = mt_find(pk, blockIdx.x);
  const long base = (long)(blockIdx.x - pk.start[t]) * MT_CHUNK;
  const long cnt = min((long)MT_CHUNK, pk.numel[t] - base);
  float* x = static_cast<float*>(pk.ptr[0][t]) + base;
  mt_stream(cnt, aligned16(x),
            [&](int i) {
              float4 v = reinterpret_cast<float4*>(x)[i];
              v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
              reinterpret_cast<float4*>(x)[i] = v;
            },
            [&](long i) { x[i] *= sc; });
}

// ------------------------------------------------------------------------------------------------ gradient norm
__global__ void __launch_bounds__(MT_THREADS) mt_sumsq_kernel(const __grid_constant__ MtPack<1> pk,
                                                              float* __restrict__ partials) {
  __shared__ float red[MT_THREADS / 32];
  const int t = mt_find(pk, blockIdx.x);
  const long base = (long)(blockIdx.x - pk.start[t]) * MT_CHUNK;
  const long cnt = min((long)MT_CHUNK, pk.numel[t] - base);
  const float* x = static_cast<const float*>(pk.ptr[0][t]) + base;
  float acc = 0.f;
  mt_stream(cnt, aligned16(x),
            [&](int i) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
              acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
            },
            [&](long i) { acc = fmaf(x[i], x[i], acc); });
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(1024) mt_norm_finish_kernel(const float* __restrict__ partials, long n, float max_norm,
                                                              float* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (long i = threadIdx.x; i < n; i += 1024) acc += (double)partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 32; ++w) s += red[w];
    const float norm = (float)sqrt(s);
    out[0] = norm;
    const float coef = max_norm / (norm + 1e-6f);  // clip_grad_norm_: clamp(max_norm / (total_norm + 1e-6), max=1)
    out[1] = (coef > 1.f) ? 1.f : coef;            // a NaN norm stays NaN, as torch.clamp leaves it
  }
}

// ------------------------------------------------------------------------------------------------ AGC
// One CTA per unit (optimizer.py:4-9 `unitwise_norm`): two norms, then (only if clipped) one rescaling pass.
constexpr int AGC_THREADS = 128;
__global__ void __launch_bounds__(AGC_THREADS) mt_agc_kernel(const __grid_constant__ MtPack<3> pk, float clipping,
                                                             float eps) {
  __shared__ float red[AGC_THREADS / 32];
  const int t = mt_find(pk, blockIdx.x);
  const long units = (long)(size_t)pk.ptr[2][t];  // unit count rides in the third pointer slot
  const long cols = pk.numel[t] / units;
  const long off = (long)(blockIdx.x - pk.start[t]) * cols;
  const float* p = static_cast<const float*>(pk.ptr[0][t]) + off;
  float* g = static_cast<float*>(pk.ptr[1][t]) + off;
  const bool vec = aligned16(p) && aligned16(g) && (cols & 3) == 0;
  float pp = 0.f, gg = 0.f;
  if (vec) {
    for (long i = threadIdx.x; i < (cols >> 2); i += AGC_THREADS) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p) + i);
      const float4 b = reinterpret_cast<const float4*>(g)[i];
      pp += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      gg += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    }
  } else {
    for (long i = threadIdx.x; i < cols; i += AGC_THREADS) { pp = fmaf(p[i], p[i], pp); gg = fmaf(g[i], g[i], gg); }
  }
  pp = block_sum(pp, red);
  gg = block_sum(gg, red);
  const float max_norm = fmaxf(sqrtf(pp), eps) * clipping;
  const float g_norm = sqrtf(gg);
  if (g_norm < max_norm) return;  // torch.where(g_norm < max_norm, grad, clipped): untouched
  const float sc = max_norm / fmaxf(g_norm, 1e-6f);
  if (vec) {
    for (long i = threadIdx.x; i < (cols >> 2); i += AGC_THREADS) {
      float4 b = reinterpret_cast<float4*>(g)[i];
      b.x *= sc; b.y *= sc; b.z *= sc; b.w *= sc;
      reinterpret_cast<float4*>(g)[i] = b;
    }
  } else {
    for (long i = threadIdx.x; i < cols; i += AGC_THREADS) g[i] *= sc;
  }
}

// ------------------------------------------------------------------------------------------------ AdamW
struct AdamArgs {
  float decay_mul;   // 1 - lr * weight_decay
  float w1;          // 1 - beta1
  float beta2, w2;   // beta2, 1 - beta2
  float step_size;   // lr / (1 - beta1^step)
  float bc2_sqrt;    // sqrt(1 - beta2^step)
  float eps;
};

__global__ void __launch_bounds__(MT_THREADS) mt_adamw_kernel(const __grid_constant__ MtPack<5> pk, AdamArgs a,
                                                              const float* __restrict__ grad_scale) {
  const int t = mt_find(pk, blockIdx.x);
  const long base = (long)(blockIdx.x - pk.start[t]) * MT_CHUNK;
  const long cnt = min((long)MT_CHUNK, pk.numel[t] - base);
  float* p = static_cast<float*>(pk.ptr[0][t]) + base;
  const float* g = static_cast<const float*>(pk.ptr[1][t]) + base;
  float* m = static_cast<float*>(pk.ptr[2][t]) + base;
  float* v = static_cast<float*>(pk.ptr[3][t]) + base;
  bf16* pb = pk.ptr[4][t] ? static_cast<bf16*>(pk.ptr[4][t]) + base : nullptr;
  const float gs = grad_scale ? __ldg(grad_scale) : 1.f;
  // torch/optim/adamw.py single-tensor order: decay, lerp, mul+addcmul, sqrt/bc2+eps, addcdiv
  auto upd = [&](float& pe, float ge, float& me, float& ve) {
    ge *= gs;
    pe *= a.decay_mul;
    me = me + a.w1 * (ge - me);
    ve = ve * a.beta2 + a.w2 * ge * ge;
    const float denom = sqrtf(ve) / a.bc2_sqrt + a.eps;
    pe = pe - a.step_size * (me / denom);
  };
  const bool vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && (!pb || aligned8(pb));
  mt_stream(cnt, vec,
            [&](int i) {
              float4 pe = reinterpret_cast<float4*>(p)[i];
              const float4 ge = __ldg(reinterpret_cast<const float4*>(g) + i);
              float4 me = reinterpret_cast<float4*>(m)[i];
              float4 ve = reinterpret_cast<float4*>(v)[i];
              upd(pe.x, ge.x, me.x, ve.x); upd(pe.y, ge.y, me.y, ve.y);
              upd(pe.z, ge.z, me.z, ve.z); upd(pe.w, ge.w, me.w, ve.w);
              reinterpret_cast<float4*>(p)[i] = pe;
              reinterpret_cast<float4*>(m)[i] = me;
              reinterpret_cast<float4*>(v)[i] = ve;
              if (pb) reinterpret_cast<uint2*>(pb)[i] = make_uint2(pack_bf16(pe.x, pe.y), pack_bf16(pe.z, pe.w));
            },
            [&](long i) {
              float pe = p[i], me = m[i], ve = v[i];
              upd(pe, g[i], me, ve);
              p[i] = pe; m[i] = me; v[i] = ve;
              if (pb) pb[i] = __float2bfloat16(pe);
            });
}

// ------------------------------------------------------------------------------------------------ MixLoss + accuracy
constexpr int ML_THREADS = 128;
__global__ void __launch_bounds__(ML_THREADS)
mix_loss_kernel(const float* __restrict__ logits, long ld, const long* __restrict__ target1,
                const long* __restrict__ target2, const float* __restrict__ inter, int n_class, float u, float hi,
                float loss_scale, float* __restrict__ loss, float* __restrict__ row_loss, float* __restrict__ dlogits,
                int* __restrict__ correct, int topk) {
  __shared__ float red[ML_THREADS / 32];
  const int r = blockIdx.x;
  const float* x = logits + (long)r * ld;
  const long t1l = target1[r], t2l = target2 ? target2[r] : t1l;
  // A label outside [0, n_class) (e.g. nn.CrossEntropyLoss's ignore_index = -100, which MixLoss does not define) must
  // not become an out-of-bounds read: such a row contributes no loss, a zero gradient and no hit.
  if (t1l < 0 || t1l >= n_class || t2l < 0 || t2l >= n_class) {
    if (dlogits)
      for (int k = threadIdx.x; k < n_class; k += ML_THREADS) dlogits[(long)r * n_class + k] = 0.f;
    if (row_loss && threadIdx.x == 0) row_loss[r] = 0.f;
    return;
  }
  const int t1 = (int)t1l, t2 = (int)t2l;
  const float w = inter ? inter[r] : 1.f;
  const float xt = x[t1];
  // pass 1: row maximum, rank of the target1 logit
  float mx = -INFINITY, above = 0.f;
  for (int k = threadIdx.x; k < n_class; k += ML_THREADS) {
    const float v = x[k];
    mx = fmaxf(mx, v);
    above += (v > xt) ? 1.f : 0.f;
  }
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  above = block_sum(above, red);
  if (correct && threadIdx.x == 0 && xt == xt) {  // a NaN target logit (diverged run) is a miss, not rank 0
    if (above < 1.f) atomicAdd(correct, 1);
    if (above < (float)topk) atomicAdd(correct + 1, 1);
  }
  if (!loss && !row_loss && !dlogits) return;  // accuracy only
  // pass 2: log-sum-exp
  float se = 0.f;
  for (int k = threadIdx.x; k < n_class; k += ML_THREADS) se += expf(x[k] - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  // pass 3: KL terms and gradient
  const float w1 = 1.f - w;
  float kl = 0.f;
  for (int k = threadIdx.x; k < n_class; k += ML_THREADS) {
    const float a = (k == t1) ? hi : u, b = (k == t2) ? hi : u;
    const float t = __fadd_rn(__fmul_rn(w, a), __fmul_rn(w1, b));  // inter * true1 + (1 - inter) * true2
    const float logp = x[k] - lse;
    kl += ((t > 0.f) ? t * logf(t) : 0.f) - t * logp;  // kl_div pointwise: xlogy(t, t) - t * input
    if (dlogits) dlogits[(long)r * n_class + k] = (expf(logp) - t) * loss_scale;
  }
  kl = block_sum(kl, red);
  if (threadIdx.x == 0) {
    if (row_loss) row_loss[r] = kl;
    if (loss) atomicAdd(loss, kl * loss_scale);
  }
}

// ------------------------------------------------------------------------------------------------ host: list packing
inline long chunks_of(long numel) { return (numel + MT_CHUNK - 1) / MT_CHUNK; }

// Calls launch(pack, first_cta_offset) for consecutive slices of the list that fit one parameter pack.
// count(i) = CTAs tensor i needs.  Empty tensors are skipped.
template <int D, typename Count, typename Launch>
int mt_for_each_pack(const void* const* const* lists, const int64_t* numel, int32_t n, Count count, Launch launch) {
  MtPack<D> pk;
  int k = 0;
  long blocks = 0, offset = 0;
  auto flush = [&]() -> int {
    if (k == 0) return 0;
    pk.n = k;
    pk.start[k] = (int)blocks;
    const int rc = launch(pk, offset);
    offset += blocks;
    k = 0;
    blocks = 0;
    return rc;
  };
  for (int i = 0; i < n; ++i) {
    if (numel[i] <= 0) continue;
    const long c = count(i);
    if (k == MT_MAX || blocks + c > 0x7fffffffL) {
      const int rc = flush();
      if (rc) return rc;
    }
    for (int d = 0; d < D; ++d) pk.ptr[d][k] = lists[d] ? const_cast<void*>(lists[d][i]) : nullptr;
    pk.numel[k] = numel[i];
    pk.start[k] = (int)blocks;
    blocks += c;
    ++k;
  }
  return flush();
}

int check_list(const char* who, const void* const* a, const int64_t* numel, int32_t n) {
  VTB_CHECK(n >= 0 && (n == 0 || (a && numel)), -1, "%s: null list", who);
  for (int i = 0; i < n; ++i) {
    VTB_CHECK(numel[i] >= 0, -1, "%s: numel[%d] < 0", who, i);
    VTB_CHECK(numel[i] == 0 || a[i], -1, "%s: tensor %d is null", who, i);
    VTB_CHECK(((uintptr_t)a[i] & 3) == 0, -1, "%s: tensor %d is not 4-byte aligned", who, i);
  }
  return 0;
}

}  // namespace

extern "C" int64_t vtb_mt_num_chunks(const int64_t* numel, int32_t n) {
  int64_t c = 0;
  for (int i = 0; i < n; ++i)
    if (numel[i] > 0) c += chunks_of(numel[i]);
  return c;
}

extern "C" int vtb_mt_cast_f32_bf16(const void* const* src, void* const* dst, const int64_t* numel, int32_t n,
                                    vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_list("vtb_mt_cast_f32_bf16(src)", src, numel, n)) return rc;
  VTB_CHECK(n == 0 || dst, -1, "vtb_mt_cast_f32_bf16: null dst list");
  for (int i = 0; i < n; ++i)
    VTB_CHECK(numel[i] == 0 || (dst[i] && ((uintptr_t)dst[i] & 1) == 0), -1, "vtb_mt_cast_f32_bf16: bad dst %d", i);
  const void* const* lists[2] = {src, const_cast<const void* const*>(dst)};
  return mt_for_each_pack<2>(lists, numel, n, [&](int i) { return chunks_of(numel[i]); },
                             [&](const MtPack<2>& pk, long) {
                               mt_cast_kernel<<<pk.start[pk.n], MT_THREADS, 0, stream>>>(pk);
                               VTB_LAUNCH_CHECK();
                               return 0;
                             });
}

extern "C" int vtb_mt_ema(void* const* dst, const void* const* src, const int64_t* numel, int32_t n, double decay,
                          vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_list("vtb_mt_ema(dst)", const_cast<const void* const*>(dst), numel, n)) return rc;
  if (int rc = check_list("vtb_mt_ema(src)", src, numel, n)) return rc;
  const void* const* lists[2] = {const_cast<const void* const*>(dst), src};
  const float d = (float)decay, alpha = (float)(1.0 - decay);
  return mt_for_each_pack<2>(lists, numel, n, [&](int i) { return chunks_of(numel[i]); },
                             [&](const MtPack<2>& pk, long) {
                               mt_ema_kernel<<<pk.start[pk.n], MT_THREADS, 0, stream>>>(pk, d, alpha);
                               VTB_LAUNCH_CHECK();
                               return 0;
                             });
}

extern "C" int vtb_mt_grad_norm(const void* const* grad, const int64_t* numel, int32_t n, float max_norm,
                                float* partials, float* out, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_list("vtb_mt_grad_norm", grad, numel, n)) return rc;
  VTB_CHECK(out && (partials || vtb_mt_num_chunks(numel, n) == 0), -1, "vtb_mt_grad_norm: null workspace / output");
  const void* const* lists[1] = {grad};
  const int rc = mt_for_each_pack<1>(lists, numel, n, [&](int i) { return chunks_of(numel[i]); },
                                     [&](const MtPack<1>& pk, long offset) {
                                       mt_sumsq_kernel<<<pk.start[pk.n], MT_THREADS, 0, stream>>>(pk, partials + offset);
                                       VTB_LAUNCH_CHECK();
                                       return 0;
                                     });
  if (rc) return rc;
  mt_norm_finish_kernel<<<1, 1024, 0, stream>>>(partials, (long)vtb_mt_num_chunks(numel, n), max_norm, out);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_mt_scale(void* const* x, const int64_t* numel, int32_t n, const float* scale, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_list("vtb_mt_scale", const_cast<const void* const*>(x), numel, n)) return rc;
  VTB_CHECK(scale, -1, "vtb_mt_scale: null scale");
  const void* const* lists[1] = {const_cast<const void* const*>(x)};
  return mt_for_each_pack<1>(lists, numel, n, [&](int i) { return chunks_of(numel[i]); },
                             [&](const MtPack<1>& pk, long) {
                               mt_scale_kernel<<<pk.start[pk.n], MT_THREADS, 0, stream>>>(pk, scale);
                               VTB_LAUNCH_CHECK();
                               return 0;
                             });
}

extern "C" int vtb_mt_agc(const void* const* param, void* const* grad, const int64_t* numel, const int64_t* units,
                          int32_t n, float clipping, float eps, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_list("vtb_mt_agc(param)", param, numel, n)) return rc;
  if (int rc = check_list("vtb_mt_agc(grad)", const_cast<const void* const*>(grad), numel, n)) return rc;
  VTB_CHECK(n == 0 || units, -1, "vtb_mt_agc: null units");
  const void* unit_slots[MT_MAX];
  // the unit counts ride in the third pointer slot of the pack; slices of the list are packed MT_MAX at a time
  for (int first = 0; first < n; first += MT_MAX) {
    const int m = (n - first < MT_MAX) ? n - first : MT_MAX;
    for (int i = 0; i < m; ++i) {
      const int64_t u = units[first + i];
      VTB_CHECK(numel[first + i] == 0 || (u > 0 && numel[first + i] % u == 0), -1,
                "vtb_mt_agc: units[%d]=%lld does not divide numel %lld", first + i, (long long)u,
                (long long)numel[first + i]);
      unit_slots[i] = reinterpret_cast<const void*>((size_t)u);
    }
    const void* const* lists[3] = {param + first, const_cast<const void* const*>(grad) + first, unit_slots};
    const int rc = mt_for_each_pack<3>(lists, numel + first, m, [&](int i) { return (long)units[first + i]; },
                                       [&](const MtPack<3>& pk, long) {
                                         mt_agc_kernel<<<pk.start[pk.n], AGC_THREADS, 0, stream>>>(pk, clipping, eps);
                                         VTB_LAUNCH_CHECK();
                                         return 0;
                                       });
    if (rc) return rc;
  }
  return 0;
}

extern "C" int vtb_mt_adamw(void* const* param, const void* const* grad, void* const* exp_avg, void* const* exp_avg_sq,
                            void* const* p_bf16, const int64_t* numel, int32_t n, double lr, double beta1, double beta2,
                            double eps, double weight_decay, int64_t step, const float* grad_scale,
                            vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_list("vtb_mt_adamw(param)", const_cast<const void* const*>(param), numel, n)) return rc;
  if (int rc = check_list("vtb_mt_adamw(grad)", grad, numel, n)) return rc;
  if (int rc = check_list("vtb_mt_adamw(exp_avg)", const_cast<const void* const*>(exp_avg), numel, n)) return rc;
  if (int rc = check_list("vtb_mt_adamw(exp_avg_sq)", const_cast<const void* const*>(exp_avg_sq), numel, n)) return rc;
  VTB_CHECK(step >= 1, -1, "vtb_mt_adamw: step must be >= 1 (got %lld)", (long long)step);
  VTB_CHECK(lr >= 0 && beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && eps >= 0 && weight_decay >= 0, -1,
            "vtb_mt_adamw: bad hyper-parameters");
  AdamArgs a;
  a.decay_mul = (float)(1.0 - lr * weight_decay);
  a.w1 = (float)(1.0 - beta1);
  a.beta2 = (float)beta2;
  a.w2 = (float)(1.0 - beta2);
  a.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.eps = (float)eps;
  const void* const* lists[5] = {const_cast<const void* const*>(param), grad, const_cast<const void* const*>(exp_avg),
                                 const_cast<const void* const*>(exp_avg_sq), const_cast<const void* const*>(p_bf16)};
  return mt_for_each_pack<5>(lists, numel, n, [&](int i) { return chunks_of(numel[i]); },
                             [&](const MtPack<5>& pk, long) {
                               mt_adamw_kernel<<<pk.start[pk.n], MT_THREADS, 0, stream>>>(pk, a, grad_scale);
                               VTB_LAUNCH_CHECK();
                               return 0;
                             });
}

extern "C" int vtb_mix_loss(const float* logits, int64_t ld, const int64_t* target1, const int64_t* target2,
                            const float* inter, int32_t rows, int32_t n_class, double eps, float loss_scale, float* loss,
                            float* row_loss, float* dlogits, int32_t* correct, int32_t topk, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(logits && target1, -1, "vtb_mix_loss: null pointer");
  VTB_CHECK(rows >= 0 && n_class > 0 && ld >= n_class, -1, "vtb_mix_loss: bad shape rows=%d n_class=%d ld=%lld", rows,
            n_class, (long long)ld);
  VTB_CHECK(eps >= 0 && eps <= 1, -1, "vtb_mix_loss: eps out of [0, 1]");
  VTB_CHECK(!correct || (topk >= 1 && topk <= n_class), -1, "vtb_mix_loss: topk=%d out of [1, n_class]", topk);
  if (rows == 0) return 0;
  static_assert(sizeof(long) == sizeof(int64_t), "LP64 expected");
  const float u = (float)(eps / n_class), hi = (float)(1.0 - eps + eps / n_class);
  mix_loss_kernel<<<rows, ML_THREADS, 0, stream>>>(logits, ld, reinterpret_cast<const long*>(target1),
                                                   reinterpret_cast<const long*>(target2), inter, n_class, u, hi,
                                                   loss_scale, loss, row_loss, dlogits, correct, topk);
  VTB_LAUNCH_CHECK();
  return 0;
}
