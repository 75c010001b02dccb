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

// Address of float4 #v of logical row r = row_base(r) + seg_off(v): the patchify gather (swin:15-22) splits into a
// per-row part (three 32-bit divisions per ROW) and a per-lane part that does not depend on the row at all, so the
// kernels below compute seg_off once per thread and row_base once per row instead of dividing per element.
__device__ __forceinline__ long row_base(const RowMap& g, long r) {
  if (g.patch_s <= 1) return r * (long)g.cols;
  const int s = g.patch_s;
  const unsigned Ho = g.Hin / s, Wo = g.Win / s;
  const unsigned ri = (unsigned)r;       // rows < 2^31 (checked on the host)
  const unsigned t = ri / Wo, bx = ri - t * Wo;
  const unsigned b = t / Ho, by = t - b * Ho;
  return (((long)b * g.Hin + (long)by * s) * g.Win + (long)bx * s) * g.C;
}
__device__ __forceinline__ int seg_off(const RowMap& g, int v) {
  const int f = v * 4;
  if (g.patch_s <= 1) return f;
  const int s = g.patch_s;
  const int seg = f / g.C;  // sy*s + sx
  const int c = f - seg * g.C;
  const int sy = seg / s, sx = seg - sy * s;
  return (sy * g.Win + sx) * g.C + c;
}

template <bool OUT_F32, int MAX_V4_PER_LANE>
__global__ void __launch_bounds__(LN_THREADS)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, long rows, RowMap g, void* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out,
              const float* __restrict__ rowmod_add, int group_rows) {
  const int lane = threadIdx.x & 31;
  const int nv = g.cols >> 2;
  int soff[MAX_V4_PER_LANE];
#pragma unroll
  for (int i = 0; i < MAX_V4_PER_LANE; ++i) soff[i] = seg_off(g, lane + i * 32);
  for (long r = (long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5); r < rows;
       r += (long)gridDim.x * LN_WARPS) {
    float4 xv[MAX_V4_PER_LANE];
    float s = 0.f;
    const float* xr = x + row_base(g, r);
#pragma unroll
    for (int i = 0; i < MAX_V4_PER_LANE; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        xv[i] = *reinterpret_cast<const float4*>(xr + soff[i]);
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
  int soff[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    soff[i] = seg_off(g, lane + i * 32);
  }

  for (long r = (long)blockIdx.x * LN_WARPS + (threadIdx.x >> 5); r < rows;
       r += (long)gridDim.x * LN_WARPS) {
    float4 xv[NV];
    float4 dvf[DY_F32 ? NV : 1];
    uint2 dvh[DY_F32 ? 1 : NV];
    const long rbase = row_base(g, r);
    // issue every load of the row first
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        xv[i] = *reinterpret_cast<const float4*>(x + rbase + soff[i]);
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
        const long off = rbase + soff[i];
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


// ------------------------------------------------------------------------------------------------------------
// Streaming variants for dense rows (the 2-per-block LayerNorms of every model): HBM-bound, so the inputs are
// staged through shared memory by bulk async copies (cp.async.bulk + mbarrier) issued several tiles ahead by a
// producer warp — the bytes in flight per SM no longer depend on registers or on where the consumer warps are
// in their reduction chains.  One persistent CTA per SM: 1 producer warp + GROUPS x 4 consumer warps; tile i of
// a CTA is consumed by group (i % GROUPS), one warp per row, and handed back through an `empty` mbarrier.
constexpr int ST_MAX_STAGES = 12;
bool g_ln_stream = true;  // vtb_set_option("ln_stream", 0) falls back to the register-resident kernels (A/B timing)
constexpr int ST_GROUP_WARPS = 4;

// sum over the LPR lanes that share a row (LPR = 32: the whole warp; LPR = 8: four rows per warp for narrow rows,
// so that all lanes load and the shuffle chain is 3 steps instead of 5)
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct StreamGeom {
  long rows;
  int cols, tile_rows, stages;
  long n_tiles;
};

template <int NV, bool OUT_F32, int GROUPS, int LPR>
__global__ void __launch_bounds__((GROUPS * ST_GROUP_WARPS + 1) * 32, 1)
ln_fwd_stream_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, StreamGeom g, void* __restrict__ y,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(st_smem);
  uint64_t* empty = full + ST_MAX_STAGES;
  float* ring = reinterpret_cast<float*>(st_smem + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = g.cols >> 2;
  const size_t tile_elems = (size_t)g.tile_rows * g.cols;
  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ST_GROUP_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == GROUPS * ST_GROUP_WARPS) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        const long r0 = tile * g.tile_rows;
        const uint32_t bytes = (uint32_t)(min((long)g.tile_rows, g.rows - r0) * g.cols * 4);
        mbar_expect_tx(&full[s], bytes);
        bulk_g2s(ring + s * tile_elems, x + r0 * g.cols, bytes, &full[s]);
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
    }
    return;
  }
  const int grp = warp / ST_GROUP_WARPS, wl = warp % ST_GROUP_WARPS;
  constexpr int RPW = 32 / LPR;                 // rows a warp works on at a time
  const int sub = lane / LPR, ll = lane % LPR;  // row slot inside the warp, lane inside the row
  constexpr bool GB_REGS = NV <= 6;  // affine parameters held in registers when they fit
  float4 gm[GB_REGS ? NV : 1], bt[GB_REGS ? NV : 1];
  if (GB_REGS) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = ll + i * LPR;
      if (v < nv) {
        gm[GB_REGS ? i : 0] = __ldg(reinterpret_cast<const float4*>(gamma) + v);
        bt[GB_REGS ? i : 0] = __ldg(reinterpret_cast<const float4*>(beta) + v);
      }
    }
  }
  const float inv_cols = 1.f / g.cols;
  long it = 0;
  int s = 0;
  uint32_t ph = 0;
  for (long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
    if ((int)(it % GROUPS) == grp) {
      mbar_wait(&full[s], ph);
      const float* xs = ring + s * tile_elems;
      const long r0 = tile * g.tile_rows;
      const int rows_here = (int)min((long)g.tile_rows, g.rows - r0);
      for (int rb = wl * RPW; rb < rows_here; rb += ST_GROUP_WARPS * RPW) {
        const int rl = rb + sub;
        const bool rv = rl < rows_here;  // (warp-uniform when LPR == 32)
        const float4* xr = reinterpret_cast<const float4*>(xs + (size_t)(rv ? rl : rb) * g.cols);
        float4 xv[NV];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = ll + i * LPR;
          if (v < nv) {
            xv[i] = xr[v];
            sum += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
          }
        }
        const float mean = group_sum<LPR>(sum) * inv_cols;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = ll + i * LPR;
          if (v < nv) {
            const float a = xv[i].x - mean, b = xv[i].y - mean, c = xv[i].z - mean, d = xv[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
          }
        }
        const float rstd = rsqrtf(group_sum<LPR>(q) * inv_cols + eps);
        const long r = r0 + rl;
        if (ll == 0 && rv) {
          if (mean_out) mean_out[r] = mean;
          if (rstd_out) rstd_out[r] = rstd;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int v = ll + i * LPR;
          if (v < nv && rv) {
            const float4 gmi = GB_REGS ? gm[GB_REGS ? i : 0] : __ldg(reinterpret_cast<const float4*>(gamma) + v);
            const float4 bti = GB_REGS ? bt[GB_REGS ? i : 0] : __ldg(reinterpret_cast<const float4*>(beta) + v);
            float4 o;
            o.x = (xv[i].x - mean) * rstd * gmi.x + bti.x;
            o.y = (xv[i].y - mean) * rstd * gmi.y + bti.y;
            o.z = (xv[i].z - mean) * rstd * gmi.z + bti.z;
            o.w = (xv[i].w - mean) * rstd * gmi.w + bti.w;
            if (OUT_F32) {
              reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + r * (long)g.cols)[v] = o;
            } else {
              reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(y) + r * (long)g.cols)[v] =
                  make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (++s == g.stages) { s = 0; ph ^= 1; }
  }
}

// Backward.  Stage = x (f32) | dy (bf16 or f32) | dx_in (f32, optional) tiles.  Extra fused output for the
// caller's NEXT backward step: dx_bf16 = bf16(dx_out * row_scale) and (COLSUM) its column sums, which is the
// operand / bias gradient of the Linear that produced this residual stream (DropPath scale folded in).
template <int NV, bool DY_F32, bool COLSUM, int GROUPS, int LPR>
__global__ void __launch_bounds__((GROUPS * ST_GROUP_WARPS + 1) * 32, 1)
ln_bwd_stream_kernel(const void* __restrict__ dy, const float* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, StreamGeom g, const float* dx_in, float* dx_out,
                     bf16* __restrict__ dx_bf16, const float* __restrict__ row_scale, int rows_per_scale,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx_colsum) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(st_smem);
  uint64_t* empty = full + ST_MAX_STAGES;
  float* s_part = reinterpret_cast<float*>(st_smem + 256);  // [3][cols]
  uint8_t* ring = st_smem + 256 + 3 * (size_t)g.cols * sizeof(float);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = g.cols >> 2;
  const size_t tile_elems = (size_t)g.tile_rows * g.cols;
  const size_t dy_bytes = tile_elems * (DY_F32 ? 4 : 2);
  const size_t stage_bytes = tile_elems * 4 + dy_bytes + (dx_in ? tile_elems * 4 : 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ST_GROUP_WARPS);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 3 * g.cols; i += blockDim.x) s_part[i] = 0.f;
  __syncthreads();
  if (warp == GROUPS * ST_GROUP_WARPS) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        const long r0 = tile * g.tile_rows;
        const long n = min((long)g.tile_rows, g.rows - r0) * g.cols;
        uint8_t* st = ring + s * stage_bytes;
        mbar_expect_tx(&full[s], (uint32_t)(n * (4 + (DY_F32 ? 4 : 2) + (dx_in ? 4 : 0))));
        bulk_g2s(st, x + r0 * g.cols, (uint32_t)(n * 4), &full[s]);
        bulk_g2s(st + tile_elems * 4, reinterpret_cast<const uint8_t*>(dy) + r0 * g.cols * (DY_F32 ? 4 : 2),
                 (uint32_t)(n * (DY_F32 ? 4 : 2)), &full[s]);
        if (dx_in) bulk_g2s(st + tile_elems * 4 + dy_bytes, dx_in + r0 * g.cols, (uint32_t)(n * 4), &full[s]);
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int grp = warp / ST_GROUP_WARPS, wl = warp % ST_GROUP_WARPS;
    constexpr int RPW = 32 / LPR;                 // rows a warp works on at a time
    const int sub = lane / LPR, ll = lane % LPR;  // row slot inside the warp, lane inside the row
    float4 dg[NV], db[NV], dc[COLSUM ? NV : 1];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (COLSUM) dc[COLSUM ? i : 0] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float inv_cols = 1.f / g.cols;
    long it = 0;
    int s = 0;
    uint32_t ph = 0;
    for (long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++it) {
      if ((int)(it % GROUPS) == grp) {
        const long r0 = tile * g.tile_rows;
        const int rows_here = (int)min((long)g.tile_rows, g.rows - r0);
        // row statistics of this warp's first row: issued before the wait so the latency hides behind it
        float mean = 0.f, rstd = 0.f, rs = 1.f;
        if (wl * RPW + sub < rows_here) {
          mean = __ldg(mean_in + r0 + wl * RPW + sub);
          rstd = __ldg(rstd_in + r0 + wl * RPW + sub);
          if (row_scale) rs = __ldg(row_scale + (r0 + wl * RPW + sub) / rows_per_scale);
        }
        mbar_wait(&full[s], ph);
        const uint8_t* st = ring + s * stage_bytes;
        for (int rb = wl * RPW; rb < rows_here; rb += ST_GROUP_WARPS * RPW) {
          const bool rv = rb + sub < rows_here;  // (warp-uniform when LPR == 32)
          const int rl = rv ? rb + sub : rb;     // idle row slots re-read a valid row and write nothing
          const long r = r0 + rl;
          float mean_n = 0.f, rstd_n = 0.f, rs_n = 1.f;
          if (rb + sub + ST_GROUP_WARPS * RPW < rows_here) {
            mean_n = __ldg(mean_in + r + ST_GROUP_WARPS * RPW);
            rstd_n = __ldg(rstd_in + r + ST_GROUP_WARPS * RPW);
            if (row_scale) rs_n = __ldg(row_scale + (r + ST_GROUP_WARPS * RPW) / rows_per_scale);
          }
          const float4* xr = reinterpret_cast<const float4*>(st) + (size_t)rl * nv;
          const uint8_t* dyr = st + tile_elems * 4 + (size_t)rl * g.cols * (DY_F32 ? 4 : 2);
          constexpr bool KEEP_GD = NV <= 4;  // gamma*dy kept between the passes, or rebuilt from smem when registers are short
          float4 xh[NV], gd[KEEP_GD ? NV : 1];
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int v = ll + i * LPR;
            if (v < nv && rv) {
              float4 d;
              if (DY_F32) d = reinterpret_cast<const float4*>(dyr)[v];
              else {
                const uint2 h = reinterpret_cast<const uint2*>(dyr)[v];
                const float2 a = unpack_bf16(h.x), b = unpack_bf16(h.y);
                d = make_float4(a.x, a.y, b.x, b.y);
              }
              const float4 xv = xr[v];
              const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + v);
              xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd,
                                  (xv.w - mean) * rstd);
              dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y;
              dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
              db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
              const float4 t = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
              if (KEEP_GD) gd[KEEP_GD ? i : 0] = t;
              s1 += (t.x + t.y) + (t.z + t.w);
              s2 += (t.x * xh[i].x + t.y * xh[i].y) + (t.z * xh[i].z + t.w * xh[i].w);
            }
          }
          s1 = group_sum<LPR>(s1) * inv_cols;
          s2 = group_sum<LPR>(s2) * inv_cols;
          const float4* pin = reinterpret_cast<const float4*>(st + tile_elems * 4 + dy_bytes) + (size_t)rl * nv;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int v = ll + i * LPR;
            if (v < nv && rv) {
              float4 t;
              if (KEEP_GD) t = gd[KEEP_GD ? i : 0];
              else {
                float4 d;
                if (DY_F32) d = reinterpret_cast<const float4*>(dyr)[v];
                else {
                  const uint2 h = reinterpret_cast<const uint2*>(dyr)[v];
                  const float2 a = unpack_bf16(h.x), b = unpack_bf16(h.y);
                  d = make_float4(a.x, a.y, b.x, b.y);
                }
                const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + v);
                t = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
              }
              float4 o;
              o.x = rstd * (t.x - s1 - xh[i].x * s2);
              o.y = rstd * (t.y - s1 - xh[i].y * s2);
              o.z = rstd * (t.z - s1 - xh[i].z * s2);
              o.w = rstd * (t.w - s1 - xh[i].w * s2);
              if (dx_in) {
                const float4 p = pin[v];
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
              }
              reinterpret_cast<float4*>(dx_out + r * (long)g.cols)[v] = o;
              if (dx_bf16) {
                const uint32_t lo = pack_bf16(o.x * rs, o.y * rs), hi = pack_bf16(o.z * rs, o.w * rs);
                reinterpret_cast<uint2*>(dx_bf16 + r * (long)g.cols)[v] = make_uint2(lo, hi);
                if (COLSUM) {
                  const float2 f0 = unpack_bf16(lo), f1 = unpack_bf16(hi);
                  float4& c = dc[COLSUM ? i : 0];
                  c.x += f0.x; c.y += f0.y; c.z += f1.x; c.w += f1.y;
                }
              }
            }
          }
          mean = mean_n; rstd = rstd_n; rs = rs_n;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      if (++s == g.stages) { s = 0; ph ^= 1; }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = ll + i * LPR;
      if (v < nv) {
        float* pg = s_part + v * 4;
        float* pb = s_part + g.cols + v * 4;
        atomicAdd(pg + 0, dg[i].x); atomicAdd(pg + 1, dg[i].y);
        atomicAdd(pg + 2, dg[i].z); atomicAdd(pg + 3, dg[i].w);
        atomicAdd(pb + 0, db[i].x); atomicAdd(pb + 1, db[i].y);
        atomicAdd(pb + 2, db[i].z); atomicAdd(pb + 3, db[i].w);
        if (COLSUM) {
          float* pc = s_part + 2 * g.cols + v * 4;
          const float4 c = dc[COLSUM ? i : 0];
          atomicAdd(pc + 0, c.x); atomicAdd(pc + 1, c.y); atomicAdd(pc + 2, c.z); atomicAdd(pc + 3, c.w);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.cols; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, s_part[i]);
    if (dbeta) atomicAdd(dbeta + i, s_part[g.cols + i]);
    if (COLSUM) atomicAdd(dx_colsum + i, s_part[2 * g.cols + i]);
  }
}

// tile / ring geometry for the streaming kernels; returns the dynamic smem size (0 => not applicable).
// The ring depth is a multiple of the number of consumer groups so that every stage is always consumed by the
// same group: a group then waits on the phases of its own barriers strictly in order (an mbarrier parity wait is
// only meaningful for a waiter that has seen the previous phase complete).
size_t stream_geom(long rows, int cols, size_t bytes_per_elem, size_t fixed_bytes, int groups, StreamGeom* g) {
  int tr = (int)(24576 / ((size_t)cols * 4)) / ST_GROUP_WARPS * ST_GROUP_WARPS;
  if (tr < ST_GROUP_WARPS) tr = ST_GROUP_WARPS;
  if (tr > 64) tr = 64;
  const size_t budget = 200 * 1024 - fixed_bytes;  // + 256 B of mbarriers stays under the 201 KB opt-in
  int stages = 0;
  for (;;) {
    const size_t stage = (size_t)tr * cols * bytes_per_elem;
    stages = (int)(budget / stage);
    if (stages > ST_MAX_STAGES) stages = ST_MAX_STAGES;
    stages = stages / groups * groups;
    if (stages >= 2 * groups || tr <= ST_GROUP_WARPS) break;
    tr = (tr / 2 + ST_GROUP_WARPS - 1) / ST_GROUP_WARPS * ST_GROUP_WARPS;
  }
  if (stages < groups) return 0;
  g->rows = rows; g->cols = cols; g->tile_rows = tr; g->stages = stages;
  g->n_tiles = (rows + tr - 1) / tr;
  return 256 + fixed_bytes + (size_t)stages * tr * cols * bytes_per_elem;
}

int check_geom(const char* who, long rows, int cols, int patch_s, int Hin, int Win, RowMap* g) {
  VTB_CHECK(rows > 0 && rows < (1L << 31) && cols > 0, -1, "%s: bad shape rows=%ld cols=%d", who, rows, cols);
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

void vtb_ln_stream_set(bool on) { g_ln_stream = on; }

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
  const bool aligned = (((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0;
  if (patch_s <= 1 && !rowmod_add && aligned && g_ln_stream) {
    StreamGeom sg;
    constexpr int G = 3;  // 13 warps: the register file then allows 128 registers per thread
    const size_t smem = stream_geom(rows, cols, 4, 0, G, &sg);
    if (smem) {
      const int grid = (int)(sg.n_tiles < vtb_num_sms() ? sg.n_tiles : vtb_num_sms());
      const bool narrow = cols <= 128;  // <= 32 float4 per row: four rows per warp, 8 lanes each
      const int nvl = narrow ? (cols / 4 + 7) / 8 : (cols / 4 + 31) / 32;
#define LN_FWD_ST(F32, NVV, LPRV)                                                                        \
  do {                                                                                                   \
    auto kern = ln_fwd_stream_kernel<NVV, F32, G, LPRV>;                                                 \
    VTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));      \
    kern<<<grid, (G * ST_GROUP_WARPS + 1) * 32, smem, stream>>>(x, gamma, beta, eps, sg, y, mean, rstd); \
  } while (0)
#define LN_FWD_ST_NV(F32)                        \
  do {                                           \
    if (narrow) {                                \
      if (nvl <= 1) LN_FWD_ST(F32, 1, 8);        \
      else if (nvl <= 2) LN_FWD_ST(F32, 2, 8);   \
      else if (nvl <= 3) LN_FWD_ST(F32, 3, 8);   \
      else LN_FWD_ST(F32, 4, 8);                 \
    } else if (nvl <= 2) LN_FWD_ST(F32, 2, 32);  \
    else if (nvl <= 3) LN_FWD_ST(F32, 3, 32);    \
    else if (nvl <= 4) LN_FWD_ST(F32, 4, 32);    \
    else if (nvl <= 6) LN_FWD_ST(F32, 6, 32);    \
    else if (nvl <= 8) LN_FWD_ST(F32, 8, 32);    \
    else LN_FWD_ST(F32, 12, 32);                 \
  } while (0)
      if (y_f32) LN_FWD_ST_NV(true); else LN_FWD_ST_NV(false);
#undef LN_FWD_ST_NV
#undef LN_FWD_ST
      VTB_LAUNCH_CHECK();
      return 0;
    }
  }
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
                                 int32_t rows_per_scale, float* dgamma, float* dbeta, float* dx_colsum,
                                 vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RowMap g;
  int rc = check_geom("vtb_layernorm_bwd", rows, cols, patch_s, Hin, Win, &g);
  if (rc) return rc;
  VTB_CHECK(dy && x && gamma && mean && rstd && dx_out, -1, "vtb_layernorm_bwd: null pointer");
  VTB_CHECK(!row_scale || rows_per_scale > 0, -1, "vtb_layernorm_bwd: rows_per_scale");
  VTB_CHECK(!dx_colsum || dx_bf16, -1, "vtb_layernorm_bwd: dx_colsum needs dx_bf16");
  const bool aligned = (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx_in | (uintptr_t)dx_out | (uintptr_t)dx_bf16 |
                         (uintptr_t)gamma) & 15) == 0;
  const bool stream_ok = patch_s <= 1 && aligned && cols <= 768 && g_ln_stream;
  VTB_CHECK(!dx_colsum || stream_ok, -1,
            "vtb_layernorm_bwd: dx_colsum needs dense 16-byte aligned rows with cols <= 768");
  if (stream_ok) {
    StreamGeom sg;
    const bool narrow = cols <= 128;  // <= 32 float4 per row: four rows per warp, 8 lanes each
    const int nvl = narrow ? (cols / 4 + 7) / 8 : (cols / 4 + 31) / 32;
    const int groups = (!narrow && nvl > 4 && dx_colsum) ? 2 : 3;  // == G of the instantiation chosen below
    const size_t smem = stream_geom(rows, cols, 4 + (dy_f32 ? 4 : 2) + (dx_in ? 4 : 0), 3 * (size_t)cols * 4, groups, &sg);
    VTB_CHECK(smem != 0, -1, "vtb_layernorm_bwd: stream geometry");
    const int grid = (int)(sg.n_tiles < vtb_num_sms() ? sg.n_tiles : vtb_num_sms());
    // 3 consumer groups (13 warps, 128 registers) up to 512 columns; 2 groups (9 warps, 168 registers) above, where
    // the per-lane accumulators of dgamma / dbeta / column sums need the room
#define LN_BWD_ST(F32, CS, NVV, LPRV)                                                                    \
  do {                                                                                                   \
    constexpr int G = (LPRV == 32 && NVV > 4 && CS) ? 2 : 3;                                             \
    auto kern = ln_bwd_stream_kernel<NVV, F32, CS, G, LPRV>;                                             \
    VTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));      \
    kern<<<grid, (G * ST_GROUP_WARPS + 1) * 32, smem, stream>>>(                                         \
        dy, x, gamma, mean, rstd, sg, dx_in, dx_out, reinterpret_cast<bf16*>(dx_bf16), row_scale,        \
        rows_per_scale, dgamma, dbeta, dx_colsum);                                                       \
  } while (0)
#define LN_BWD_ST_NV(F32, CS)                        \
  do {                                               \
    if (narrow) {                                    \
      if (nvl <= 1) LN_BWD_ST(F32, CS, 1, 8);        \
      else if (nvl <= 2) LN_BWD_ST(F32, CS, 2, 8);   \
      else if (nvl <= 3) LN_BWD_ST(F32, CS, 3, 8);   \
      else LN_BWD_ST(F32, CS, 4, 8);                 \
    } else if (nvl <= 2) LN_BWD_ST(F32, CS, 2, 32);  \
    else if (nvl <= 3) LN_BWD_ST(F32, CS, 3, 32);    \
    else if (nvl <= 4) LN_BWD_ST(F32, CS, 4, 32);    \
    else LN_BWD_ST(F32, CS, 6, 32);                  \
  } while (0)
    if (dy_f32) { if (dx_colsum) LN_BWD_ST_NV(true, true); else LN_BWD_ST_NV(true, false); }
    else        { if (dx_colsum) LN_BWD_ST_NV(false, true); else LN_BWD_ST_NV(false, false); }
#undef LN_BWD_ST_NV
#undef LN_BWD_ST
    VTB_LAUNCH_CHECK();
    return 0;
  }
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
