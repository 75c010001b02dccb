// LayerNorm forward / backward over the fp32 residual stream (HBM-bound; warp-shuffle reductions).
// One warp per row; the row is kept in registers between the statistics pass and the normalise pass
// so x is read exactly once.  Optional patchify gather/scatter (PatchMerge LN prologue, swin:221-227).
#include "common.cuh"
#include "../../include/vtb200.h"

namespace {

constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;
constexpr int MAX_COLS = 1536;  // register path: <= 12 float4 per lane (template MAXV = 4 / 8 / 12)

struct RowMap {
  int patch_s, Hin, Win, C;  // patchify geometry; patch_s <= 1 => dense rows
  int cols;
};

// pointer to the float4 #v of logical row r
__device__ __forceinline__ long row_v4_offset(const RowMap& g, long r, int v) {
  if (g.patch_s <= 1) return r * (long)g.cols + (long)v * 4;
  const int s = g.patch_s;
  const int Ho = g.Hin / s, Wo = g.Win / s;
  const int bx = (int)(r % Wo);
  const long t = r / Wo;
  const int by = (int)(t % Ho);
  const long b = t / Ho;
  const int f = v * 4;
  const int seg = f / g.C;  // sy*s + sx
  const int c = f - seg * g.C;
  const int sy = seg / s, sx = seg - sy * s;
  return ((b * g.Hin + (long)by * s + sy) * g.Win + (long)bx * s + sx) * g.C + c;
}

template <bool OUT_F32, int MAX_V4_PER_LANE>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, long rows, RowMap g, void* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out,
              const float* __restrict__ rowmod_add, int group_rows) {
  const int lane = threadIdx.x & 31;
  const int nv = g.cols >> 2;
  for (long r = (long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5); r < rows;
       r += (long)gridDim.x * LN_WARPS) {
    float4 xv[MAX_V4_PER_LANE];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V4_PER_LANE; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        xv[i] = *reinterpret_cast<const float4*>(x + row_v4_offset(g, r, v));
        s += xv[i].x + xv[i].y + xv[i].z + xv[i].w;
      }
    }
    const float mean = warp_sum(s) / g.cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V4_PER_LANE; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / g.cols + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
    const float* add = rowmod_add ? rowmod_add + (r % group_rows) * (long)g.cols : nullptr;
#pragma unroll
    for (int i = 0; i < MAX_V4_PER_LANE; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        const float4 gm = *reinterpret_cast<const float4*>(gamma + v * 4);
        const float4 bt = *reinterpret_cast<const float4*>(beta + v * 4);
        float4 o;
        o.x = (xv[i].x - mean) * rstd * gm.x + bt.x;
        o.y = (xv[i].y - mean) * rstd * gm.y + bt.y;
        o.z = (xv[i].z - mean) * rstd * gm.z + bt.z;
        o.w = (xv[i].w - mean) * rstd * gm.w + bt.w;
        if (add) {
          const float4 a = *reinterpret_cast<const float4*>(add + v * 4);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        if (OUT_F32) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + r * (long)g.cols + v * 4) = o;
        } else {
          *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + r * (long)g.cols + v * 4) =
              make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
        }
      }
    }
  }
}

// Backward.  One warp per row, NV = ceil(cols / 128) float4 per lane (exact, so registers stay low enough
// for >= 2 CTAs per SM: the kernel is HBM-bound and needs the loads of many rows in flight).  dy stays packed
// (bf16) in registers between the two passes; dgamma/dbeta partials live in registers across the rows a warp
// processes, are combined per CTA in shared memory, then one atomicAdd per column per CTA.
template <bool DY_F32, int NV>
__global__ void __launch_bounds__(LN_THREADS, (NV <= 2) ? 3 : ((NV <= 6) ? 2 : 1))
ln_bwd_kernel(const void* __restrict__ dy, const float* __restrict__ x,
              const float* __restrict__ gamma, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, long rows, RowMap g,
              const float* dx_in, float* dx_out,  // may alias (in-place accumulate)
              bf16* __restrict__ dx_bf16, const float* __restrict__ row_scale, int rows_per_scale,
              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float s_part[];  // [2][cols]
  const int lane = threadIdx.x & 31;
  const int nv = g.cols >> 2;
  for (int i = threadIdx.x; i < 2 * g.cols; i += LN_THREADS) s_part[i] = 0.f;
  __syncthreads();

  float4 dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (long r = (long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5); r < rows;
       r += (long)gridDim.x * LN_WARPS) {
    float4 xv[NV];
    float4 dvf[DY_F32 ? NV : 1];
    uint2 dvh[DY_F32 ? 1 : NV];
    // issue every load of the row first
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        xv[i] = *reinterpret_cast<const float4*>(x + row_v4_offset(g, r, v));
        if (DY_F32)
          dvf[DY_F32 ? i : 0] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + r * (long)g.cols + v * 4);
        else
          dvh[DY_F32 ? 0 : i] = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy) + r * (long)g.cols + v * 4);
      }
    }
    const float mean = mean_in[r], rstd = rstd_in[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        float4 d;
        if (DY_F32) d = dvf[DY_F32 ? i : 0];
        else {
          const float2 a = unpack_bf16(dvh[DY_F32 ? 0 : i].x), b = unpack_bf16(dvh[DY_F32 ? 0 : i].y);
          d = make_float4(a.x, a.y, b.x, b.y);
        }
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + v * 4));
        // xv <- xhat
        xv[i] = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd,
                            (xv[i].w - mean) * rstd);
        dg[i].x += d.x * xv[i].x; dg[i].y += d.y * xv[i].y;
        dg[i].z += d.z * xv[i].z; dg[i].w += d.w * xv[i].w;
        db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
        const float gx = d.x * gm.x, gy = d.y * gm.y, gz = d.z * gm.z, gw = d.w * gm.w;
        s1 += gx + gy + gz + gw;
        s2 += gx * xv[i].x + gy * xv[i].y + gz * xv[i].z + gw * xv[i].w;
      }
    }
    s1 = warp_sum(s1) / g.cols;
    s2 = warp_sum(s2) / g.cols;
    const float rs = row_scale ? row_scale[r / rows_per_scale] : 1.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        float4 d;
        if (DY_F32) d = dvf[DY_F32 ? i : 0];
        else {
          const float2 a = unpack_bf16(dvh[DY_F32 ? 0 : i].x), b = unpack_bf16(dvh[DY_F32 ? 0 : i].y);
          d = make_float4(a.x, a.y, b.x, b.y);
        }
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + v * 4));
        float4 o;
        o.x = rstd * (d.x * gm.x - s1 - xv[i].x * s2);
        o.y = rstd * (d.y * gm.y - s1 - xv[i].y * s2);
        o.z = rstd * (d.z * gm.z - s1 - xv[i].z * s2);
        o.w = rstd * (d.w * gm.w - s1 - xv[i].w * s2);
        const long off = row_v4_offset(g, r, v);
        if (dx_in) {
          const float4 p = *reinterpret_cast<const float4*>(dx_in + off);
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        *reinterpret_cast<float4*>(dx_out + off) = o;
        if (dx_bf16) {
          *reinterpret_cast<uint2*>(dx_bf16 + off) =
              make_uint2(pack_bf16(o.x * rs, o.y * rs), pack_bf16(o.z * rs, o.w * rs));
        }
      }
    }
  }
  // CTA-level reduction of the parameter gradients
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      float* pg = s_part + v * 4;
      float* pb = s_part + g.cols + v * 4;
      atomicAdd(pg + 0, dg[i].x); atomicAdd(pg + 1, dg[i].y);
      atomicAdd(pg + 2, dg[i].z); atomicAdd(pg + 3, dg[i].w);
      atomicAdd(pb + 0, db[i].x); atomicAdd(pb + 1, db[i].y);
      atomicAdd(pb + 2, db[i].z); atomicAdd(pb + 3, db[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.cols; i += LN_THREADS) {
    if (dgamma) atomicAdd(dgamma + i, s_part[i]);
    if (dbeta) atomicAdd(dbeta + i, s_part[g.cols + i]);
  }
}

int check_geom(const char* who, long rows, int cols, int patch_s, int Hin, int Win, RowMap* g) {
  VTB_CHECK(rows > 0 && cols > 0, -1, "%s: bad shape rows=%ld cols=%d", who, rows, cols);
  VTB_CHECK(cols % 4 == 0 && cols <= MAX_COLS, -1,
            "%s: cols=%d must be a multiple of 4 and <= %d", who, cols, MAX_COLS);
  g->patch_s = patch_s; g->Hin = Hin; g->Win = Win; g->cols = cols; g->C = cols;
  if (patch_s > 1) {
    VTB_CHECK(Hin % patch_s == 0 && Win % patch_s == 0 && cols % (patch_s * patch_s) == 0, -1,
              "%s: patchify geometry s=%d H=%d W=%d cols=%d", who, patch_s, Hin, Win, cols);
    g->C = cols / (patch_s * patch_s);
    VTB_CHECK(g->C % 4 == 0, -1, "%s: patchify needs C %% 4 == 0 (C=%d)", who, g->C);
    VTB_CHECK(rows % ((long)(Hin / patch_s) * (Win / patch_s)) == 0, -1, "%s: rows vs geometry", who);
  }
  return 0;
}

}  // namespace

extern "C" int vtb_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps,
                                 int64_t rows, int32_t cols, int32_t patch_s, int32_t Hin,
                                 int32_t Win, void* y, int32_t y_f32, float* mean, float* rstd,
                                 const float* rowmod_add, int32_t group_rows,
                                 vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RowMap g;
  int rc = check_geom("vtb_layernorm_fwd", rows, cols, patch_s, Hin, Win, &g);
  if (rc) return rc;
  VTB_CHECK(x && gamma && beta && y, -1, "vtb_layernorm_fwd: null pointer");
  VTB_CHECK(!rowmod_add || group_rows > 0, -1, "vtb_layernorm_fwd: group_rows");
  long blocks = (rows + LN_WARPS - 1) / LN_WARPS;
  const long cap = (long)vtb_num_sms() * 16;
  if (cap > 0 && blocks > cap) blocks = cap;
#define LN_FWD(F32, MV)                                                                       \
  ln_fwd_kernel<F32, MV><<<(int)blocks, LN_THREADS, 0, stream>>>(x, gamma, beta, eps, rows, g, y, \
                                                                 mean, rstd, rowmod_add, group_rows)
  if (y_f32) {
    if (cols <= 512) LN_FWD(true, 4); else if (cols <= 1024) LN_FWD(true, 8); else LN_FWD(true, 12);
  } else {
    if (cols <= 512) LN_FWD(false, 4); else if (cols <= 1024) LN_FWD(false, 8); else LN_FWD(false, 12);
  }
#undef LN_FWD
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int vtb_layernorm_bwd(const void* dy, int32_t dy_f32, const float* x, const float* gamma,
                                 const float* mean, const float* rstd, int64_t rows, int32_t cols,
                                 int32_t patch_s, int32_t Hin, int32_t Win, const float* dx_in,
                                 float* dx_out, void* dx_bf16, const float* row_scale,
                                 int32_t rows_per_scale, float* dgamma, float* dbeta,
                                 vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RowMap g;
  int rc = check_geom("vtb_layernorm_bwd", rows, cols, patch_s, Hin, Win, &g);
  if (rc) return rc;
  VTB_CHECK(dy && x && gamma && mean && rstd && dx_out, -1, "vtb_layernorm_bwd: null pointer");
  VTB_CHECK(!row_scale || rows_per_scale > 0, -1, "vtb_layernorm_bwd: rows_per_scale");
  long blocks = (rows + LN_WARPS - 1) / LN_WARPS;
  const long cap = (long)vtb_num_sms() * 6;  // few CTAs => few global atomics per column; 2-3 resident per SM
  if (cap > 0 && blocks > cap) blocks = cap;
  const size_t smem = 2 * (size_t)cols * sizeof(float);
#define LN_BWD(F32, MV)                                                \
  ln_bwd_kernel<F32, MV><<<(int)blocks, LN_THREADS, smem, stream>>>(   \
      dy, x, gamma, mean, rstd, rows, g, dx_in, dx_out, reinterpret_cast<bf16*>(dx_bf16), \
      row_scale, rows_per_scale, dgamma, dbeta)
#define LN_BWD_NV(F32)                                   \
  do {                                                   \
    const int nvl = (cols / 4 + 31) / 32;                \
    if (nvl <= 1) LN_BWD(F32, 1);                        \
    else if (nvl <= 2) LN_BWD(F32, 2);                   \
    else if (nvl <= 3) LN_BWD(F32, 3);                   \
    else if (nvl <= 4) LN_BWD(F32, 4);                   \
    else if (nvl <= 6) LN_BWD(F32, 6);                   \
    else if (nvl <= 8) LN_BWD(F32, 8);                   \
    else LN_BWD(F32, 12);                                \
  } while (0)
  if (dy_f32) LN_BWD_NV(true); else LN_BWD_NV(false);
#undef LN_BWD_NV
#undef LN_BWD
  VTB_LAUNCH_CHECK();
  return 0;
}
