// Fused multi-head attention for every variant on the path (global / shifted-window / halo / SRA),
// forward and backward, through one geometry descriptor (include/vtb200.h: vtb_attn_params).
// Scores and probabilities stay in registers; window partition, cyclic shift, halo gather with zero
// padding and the relative-position bias / mask are folded into the load addressing, so none of the
// reference's roll / reshape / unfold / masked_fill / softmax kernels (and their HBM traffic) exist.
//
// This file holds the dispatcher (vtb_attention_fwd / vtb_attention_bwd) and the mma.sync m16n8k16 kernels that remain
// for the shapes the tcgen05 kernels do not cover: Halo blocks (49 x 169), dh = 64 windows (Twins-LSA), global attention
// with more than 256 queries in backward (PVT stages 1-2, streaming kernels) and unaligned / odd shapes.  Global
// attention with dh = 64 and <= 256 tokens runs in attention_tc.cu, shifted-window attention with dh = 32 in
// attention_win_tc.cu (both tcgen05 / TMEM).
#include "common.cuh"
#include "../../include/vtb200.h"
#include <math.h>

namespace {

constexpr int BQ = 64;    // rows per CTA tile (4 warps x 16)
constexpr int BKV = 64;   // keys per inner block
constexpr int NTHREADS = 128;
constexpr int MAX_POS = 512;  // relative-position table rows handled in smem (swin 169, halo 253)

struct Geom {
  int mode, batch, heads, dh, nq, nkv, Hs, Ws, window, shift, halo;
  int nwx, nw;  // windows per row / per image
};

__device__ __forceinline__ long q_token(const Geom& g, int grp, int i) {
  if (i >= g.nq) return -1;
  if (g.mode == VTB_ATTN_GLOBAL) return (long)grp * g.nq + i;
  const int b = grp / g.nw, wi = grp - b * g.nw;
  const int wy = wi / g.nwx, wx = wi - wy * g.nwx;
  const int ty = i / g.window, tx = i - ty * g.window;
  int y = wy * g.window + ty, x = wx * g.window + tx;
  if (g.mode == VTB_ATTN_WINDOW && g.shift) {
    y += g.shift; if (y >= g.Hs) y -= g.Hs;
    x += g.shift; if (x >= g.Ws) x -= g.Ws;
  }
  return ((long)b * g.Hs + y) * g.Ws + x;
}
__device__ __forceinline__ long kv_token(const Geom& g, int grp, int j) {
  if (j >= g.nkv) return -1;
  if (g.mode == VTB_ATTN_GLOBAL) return (long)grp * g.nkv + j;
  if (g.mode == VTB_ATTN_WINDOW) return q_token(g, grp, j);
  const int b = grp / g.nw, wi = grp - b * g.nw;
  const int by = wi / g.nwx, bx = wi - by * g.nwx;
  const int K = g.window + 2 * g.halo;
  const int ky = j / K, kx = j - ky * K;
  const int y = by * g.window - g.halo + ky, x = bx * g.window - g.halo + kx;
  if (y < 0 || y >= g.Hs || x < 0 || x >= g.Ws) return -1;  // zero-padded slot (halo:75-80)
  return ((long)b * g.Hs + y) * g.Ws + x;
}

// cooperative load of `rows` (<=64) token rows of DH bf16 into padded smem; missing rows -> zeros
template <int DH>
__device__ __forceinline__ void load_rows(bf16 (*dst)[DH + 8], const bf16* base, int ld, int head_off,
                                          const long* toks) {
  constexpr int CH = DH / 8;  // 16-byte chunks per row
  for (int c = threadIdx.x; c < 64 * CH; c += NTHREADS) {
    const int r = c / CH, cc = c - r * CH;
    const long t = toks[r];
    const bf16* src = base + (t < 0 ? 0 : t) * (long)ld + head_off + cc * 8;
    cp_async16(smem_u32(&dst[r][cc * 8]), src, t >= 0);
  }
}

// score bias for element (i, j) of group grp / head h (clamped indices for padding rows)
struct BiasCtx {
  const float* table;   // smem copy of rel_bias[:, h] or null
  const int* pos;       // global [nq, nkv]
  const uint8_t* mask;  // global, already offset to this group's [nq, mask_ld] slice, or null
  int nq, nkv, mld;
};
__device__ __forceinline__ float score_bias(const BiasCtx& c, int i, int j, bool& masked) {
  masked = false;
  if (i >= c.nq) i = c.nq - 1;
  float b = 0.f;
  if (c.table) b = c.table[__ldg(c.pos + (long)i * c.nkv + j)];
  if (c.mask && c.mask[(long)i * c.mld + j]) masked = true;
  return b;
}

// ------------------------------------------------------------------------------------ forward
template <int DH>
__global__ void __launch_bounds__(NTHREADS)
attn_fwd_kernel(vtb_attn_params p, Geom g, int q_tiles) {
  __shared__ __align__(16) bf16 sQ[BQ][DH + 8];
  __shared__ __align__(16) bf16 sK[BKV][DH + 8];
  __shared__ __align__(16) bf16 sV[BKV][DH + 8];
  __shared__ long sQtok[BQ], sKtok[BKV];
  __shared__ float sTab[MAX_POS];

  const int qt = blockIdx.x % q_tiles;
  const int gh = blockIdx.x / q_tiles;
  const int h = gh % g.heads;
  const int grp = gh / g.heads;
  const int i0 = qt * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;

  const bf16* Q = reinterpret_cast<const bf16*>(p.q);
  const bf16* K = reinterpret_cast<const bf16*>(p.k);
  const bf16* V = reinterpret_cast<const bf16*>(p.v);

  if (threadIdx.x < BQ) sQtok[threadIdx.x] = q_token(g, grp, i0 + threadIdx.x);
  if (p.rel_bias) {
    for (int t = threadIdx.x; t < p.n_pos; t += NTHREADS) sTab[t] = p.rel_bias[(long)t * g.heads + h];
  }
  __syncthreads();
  load_rows<DH>(sQ, Q, p.ldq, h * DH, sQtok);
  cp_async_commit();

  BiasCtx bc;
  bc.table = p.rel_bias ? sTab : nullptr;
  bc.pos = p.pos;
  bc.mld = p.mask_ld > 0 ? p.mask_ld : g.nkv;
  bc.mask = p.mask ? p.mask + (long)(grp % p.n_mask) * g.nq * bc.mld : nullptr;
  bc.nq = g.nq; bc.nkv = g.nkv;

  float o[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[DH / 16][4];
  bool q_loaded = false;

  for (int j0 = 0; j0 < g.nkv; j0 += BKV) {
    __syncthreads();  // previous block's smem reads done
    if (threadIdx.x < BKV) sKtok[threadIdx.x] = kv_token(g, grp, j0 + threadIdx.x);
    __syncthreads();
    load_rows<DH>(sK, K, p.ldk, h * DH, sKtok);
    load_rows<DH>(sV, V, p.ldv, h * DH, sKtok);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (!q_loaded) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk)
        ldsm_x4(qf[kk], smem_u32(&sQ[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]));
      q_loaded = true;
    }
    // S = Q K^T  (16 x 64 per warp)
    float s[BKV / 8][4];
#pragma unroll
    for (int n = 0; n < BKV / 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int n2 = 0; n2 < BKV / 16; ++n2) {
        uint32_t kb[4];
        ldsm_x4(kb, smem_u32(&sK[n2 * 16 + (lane & 7) + ((lane >> 4) << 3)][kk * 16 + ((lane >> 3) & 1) * 8]));
        uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
        mma_bf16_16816(s[2 * n2], qf[kk], b0);
        mma_bf16_16816(s[2 * n2 + 1], qf[kk], b1);
      }
    }
    // scale + bias + mask, online softmax
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < BKV / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i0 + warp * 16 + gq + (e >> 1) * 8;
        const int j = j0 + n * 8 + 2 * tq + (e & 1);
        float v = s[n][e] * p.scale;
        if (j >= g.nkv) v = -INFINITY;
        else if (bc.table || bc.mask) {
          bool masked;
          v += score_bias(bc, i, j, masked);
          if (masked) v = -INFINITY;
        }
        s[n][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = __expf(m_run[r] - m_use[r]);  // exp(-inf) = 0 on the first block
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < BKV / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = __expf(s[n][e] - m_use[e >> 1]);
        s[n][e] = pv;
        rs[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      o[n][0] *= corr[0]; o[n][1] *= corr[0]; o[n][2] *= corr[1]; o[n][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int k2 = 0; k2 < BKV / 16; ++k2) {
      uint32_t pa[4] = {pack_bf16(s[2 * k2][0], s[2 * k2][1]), pack_bf16(s[2 * k2][2], s[2 * k2][3]),
                        pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]),
                        pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3])};
#pragma unroll
      for (int d2 = 0; d2 < DH / 16; ++d2) {
        uint32_t vb[4];
        ldsm_x4_t(vb, smem_u32(&sV[k2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][d2 * 16 + (lane >> 4) * 8]));
        uint32_t b0[2] = {vb[0], vb[1]}, b1[2] = {vb[2], vb[3]};
        mma_bf16_16816(o[2 * d2], pa, b0);
        mma_bf16_16816(o[2 * d2 + 1], pa, b1);
      }
    }
  }

  // finalise: O / l, write; lse = m + log(l)
  bf16* O = reinterpret_cast<bf16*>(p.o);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = warp * 16 + gq + r * 8;
    const long tok = sQtok[row];
    if (tok < 0) continue;
    const float inv = 1.f / l_run[r];
    bf16* dst = O + tok * (long)p.ldo + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * tq) =
          pack_bf16(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
    }
    if (tq == 0 && p.lse)
      p.lse[((long)grp * g.heads + h) * g.nq + i0 + row] = m_run[r] + __logf(l_run[r]);
  }
}

// ------------------------------------------------------------------------------------ backward: dQ
template <int DH>
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_dq_kernel(vtb_attn_params p, Geom g, int q_tiles) {
  const int n_pos = p.n_pos;
  __shared__ __align__(16) bf16 sQ[BQ][DH + 8];
  __shared__ __align__(16) bf16 sdO[BQ][DH + 8];
  __shared__ __align__(16) bf16 sK[BKV][DH + 8];
  __shared__ __align__(16) bf16 sV[BKV][DH + 8];
  __shared__ long sQtok[BQ], sKtok[BKV];
  __shared__ float sTab[MAX_POS], sdTab[MAX_POS];
  __shared__ float sDelta[BQ], sLse[BQ];

  const int qt = blockIdx.x % q_tiles;
  const int gh = blockIdx.x / q_tiles;
  const int h = gh % g.heads;
  const int grp = gh / g.heads;
  const int i0 = qt * BQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;

  const bf16* Q = reinterpret_cast<const bf16*>(p.q);
  const bf16* K = reinterpret_cast<const bf16*>(p.k);
  const bf16* V = reinterpret_cast<const bf16*>(p.v);
  const bf16* O = reinterpret_cast<const bf16*>(p.o);
  const bf16* dO = reinterpret_cast<const bf16*>(p.dout);

  if (threadIdx.x < BQ) sQtok[threadIdx.x] = q_token(g, grp, i0 + threadIdx.x);
  if (p.rel_bias) {
    for (int t = threadIdx.x; t < n_pos; t += NTHREADS) {
      sTab[t] = p.rel_bias[(long)t * g.heads + h];
      sdTab[t] = 0.f;
    }
  }
  __syncthreads();
  load_rows<DH>(sQ, Q, p.ldq, h * DH, sQtok);
  load_rows<DH>(sdO, dO, p.lddo, h * DH, sQtok);
  cp_async_commit();
  // delta_i = sum_d dO[i,d] * O[i,d]; two threads per row
  {
    const int row = threadIdx.x >> 1, half = threadIdx.x & 1;
    const long tok = sQtok[row];
    float acc = 0.f;
    if (tok >= 0) {
      const bf16* a = dO + tok * (long)p.lddo + h * DH + half * (DH / 2);
      const bf16* b = O + tok * (long)p.ldo + h * DH + half * (DH / 2);
#pragma unroll
      for (int d = 0; d < DH / 2; d += 8) {
        const uint4 ra = *reinterpret_cast<const uint4*>(a + d);
        const uint4 rb = *reinterpret_cast<const uint4*>(b + d);
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fa = unpack_bf16(wa[q]), fb = unpack_bf16(wb[q]);
          acc += fa.x * fb.x + fa.y * fb.y;
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (half == 0) {
      sDelta[row] = acc;
      const long li = ((long)grp * g.heads + h) * g.nq + i0 + row;
      if (tok >= 0) {
        p.delta[li] = acc;
        sLse[row] = p.lse[li];
      } else {
        sLse[row] = 0.f;
      }
    }
  }

  BiasCtx bc;
  bc.table = p.rel_bias ? sTab : nullptr;
  bc.pos = p.pos;
  bc.mld = p.mask_ld > 0 ? p.mask_ld : g.nkv;
  bc.mask = p.mask ? p.mask + (long)(grp % p.n_mask) * g.nq * bc.mld : nullptr;
  bc.nq = g.nq; bc.nkv = g.nkv;

  float dq[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
  uint32_t qf[DH / 16][4], dof[DH / 16][4];
  bool loaded = false;
  float lse_r[2], dl_r[2];

  for (int j0 = 0; j0 < g.nkv; j0 += BKV) {
    __syncthreads();
    if (threadIdx.x < BKV) sKtok[threadIdx.x] = kv_token(g, grp, j0 + threadIdx.x);
    __syncthreads();
    load_rows<DH>(sK, K, p.ldk, h * DH, sKtok);
    load_rows<DH>(sV, V, p.ldv, h * DH, sKtok);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (!loaded) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        ldsm_x4(qf[kk], smem_u32(&sQ[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]));
        ldsm_x4(dof[kk], smem_u32(&sdO[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]));
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        lse_r[r] = sLse[warp * 16 + gq + r * 8];
        dl_r[r] = sDelta[warp * 16 + gq + r * 8];
      }
      loaded = true;
    }
    float s[BKV / 8][4], dp[BKV / 8][4];
#pragma unroll
    for (int n = 0; n < BKV / 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int n2 = 0; n2 < BKV / 16; ++n2) {
        uint32_t kb[4], vb[4];
        const int rr = n2 * 16 + (lane & 7) + ((lane >> 4) << 3), cc = kk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(kb, smem_u32(&sK[rr][cc]));
        ldsm_x4(vb, smem_u32(&sV[rr][cc]));
        uint32_t k0[2] = {kb[0], kb[1]}, k1[2] = {kb[2], kb[3]};
        uint32_t v0[2] = {vb[0], vb[1]}, v1[2] = {vb[2], vb[3]};
        mma_bf16_16816(s[2 * n2], qf[kk], k0);
        mma_bf16_16816(s[2 * n2 + 1], qf[kk], k1);
        mma_bf16_16816(dp[2 * n2], dof[kk], v0);
        mma_bf16_16816(dp[2 * n2 + 1], dof[kk], v1);
      }
    }
    // dS = P * (dP - delta)
#pragma unroll
    for (int n = 0; n < BKV / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        const int i = i0 + warp * 16 + gq + r * 8;
        const int j = j0 + n * 8 + 2 * tq + (e & 1);
        float ds = 0.f;
        if (j < g.nkv && i < g.nq) {
          float v = s[n][e] * p.scale;
          bool masked = false;
          if (bc.table || bc.mask) v += score_bias(bc, i, j, masked);
          if (!masked) {
            const float pv = __expf(v - lse_r[r]);
            ds = pv * (dp[n][e] - dl_r[r]);
            if (p.drel_bias && bc.table) atomicAdd(&sdTab[__ldg(bc.pos + (long)i * g.nkv + j)], ds);
          }
        }
        s[n][e] = ds;
      }
    }
    // dQ += dS K
#pragma unroll
    for (int k2 = 0; k2 < BKV / 16; ++k2) {
      uint32_t pa[4] = {pack_bf16(s[2 * k2][0], s[2 * k2][1]), pack_bf16(s[2 * k2][2], s[2 * k2][3]),
                        pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]),
                        pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3])};
#pragma unroll
      for (int d2 = 0; d2 < DH / 16; ++d2) {
        uint32_t kb[4];
        ldsm_x4_t(kb, smem_u32(&sK[k2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][d2 * 16 + (lane >> 4) * 8]));
        uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
        mma_bf16_16816(dq[2 * d2], pa, b0);
        mma_bf16_16816(dq[2 * d2 + 1], pa, b1);
      }
    }
  }

  bf16* dQ = reinterpret_cast<bf16*>(p.dq);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = warp * 16 + gq + r * 8;
    const long tok = sQtok[row];
    if (tok < 0) continue;
    bf16* dst = dQ + tok * (long)p.lddq + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n)
      *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * tq) =
          pack_bf16(dq[n][2 * r] * p.scale, dq[n][2 * r + 1] * p.scale);
  }
  if (p.drel_bias && p.rel_bias) {
    __syncthreads();
    for (int t = threadIdx.x; t < n_pos; t += NTHREADS) {
      const float v = sdTab[t];
      if (v != 0.f) atomicAdd(p.drel_bias + (long)t * g.heads + h, v);
    }
  }
}

// ------------------------------------------------------------------------------------ backward: dK, dV
template <int DH>
__global__ void __launch_bounds__(NTHREADS)
attn_bwd_dkv_kernel(vtb_attn_params p, Geom g, int kv_tiles) {
  const int n_pos = p.n_pos;
  __shared__ __align__(16) bf16 sK[BKV][DH + 8];
  __shared__ __align__(16) bf16 sV[BKV][DH + 8];
  __shared__ __align__(16) bf16 sQ[BQ][DH + 8];
  __shared__ __align__(16) bf16 sdO[BQ][DH + 8];
  __shared__ long sQtok[BQ], sKtok[BKV];
  __shared__ float sTab[MAX_POS];
  __shared__ float sDelta[BQ], sLse[BQ];

  const int kt = blockIdx.x % kv_tiles;
  const int gh = blockIdx.x / kv_tiles;
  const int h = gh % g.heads;
  const int grp = gh / g.heads;
  const int j0 = kt * BKV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;

  const bf16* Q = reinterpret_cast<const bf16*>(p.q);
  const bf16* K = reinterpret_cast<const bf16*>(p.k);
  const bf16* V = reinterpret_cast<const bf16*>(p.v);
  const bf16* dO = reinterpret_cast<const bf16*>(p.dout);

  if (threadIdx.x < BKV) sKtok[threadIdx.x] = kv_token(g, grp, j0 + threadIdx.x);
  if (p.rel_bias)
    for (int t = threadIdx.x; t < n_pos; t += NTHREADS) sTab[t] = p.rel_bias[(long)t * g.heads + h];
  __syncthreads();
  load_rows<DH>(sK, K, p.ldk, h * DH, sKtok);
  load_rows<DH>(sV, V, p.ldv, h * DH, sKtok);
  cp_async_commit();

  BiasCtx bc;
  bc.table = p.rel_bias ? sTab : nullptr;
  bc.pos = p.pos;
  bc.mld = p.mask_ld > 0 ? p.mask_ld : g.nkv;
  bc.mask = p.mask ? p.mask + (long)(grp % p.n_mask) * g.nq * bc.mld : nullptr;
  bc.nq = g.nq; bc.nkv = g.nkv;

  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  uint32_t kf[DH / 16][4], vf[DH / 16][4];
  bool loaded = false;

  for (int i0 = 0; i0 < g.nq; i0 += BQ) {
    __syncthreads();
    if (threadIdx.x < BQ) {
      const long tok = q_token(g, grp, i0 + threadIdx.x);
      sQtok[threadIdx.x] = tok;
      const long li = ((long)grp * g.heads + h) * g.nq + i0 + threadIdx.x;
      sLse[threadIdx.x] = tok >= 0 ? p.lse[li] : 0.f;
      sDelta[threadIdx.x] = tok >= 0 ? p.delta[li] : 0.f;
    }
    __syncthreads();
    load_rows<DH>(sQ, Q, p.ldq, h * DH, sQtok);
    load_rows<DH>(sdO, dO, p.lddo, h * DH, sQtok);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (!loaded) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        ldsm_x4(kf[kk], smem_u32(&sK[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]));
        ldsm_x4(vf[kk], smem_u32(&sV[warp * 16 + (lane & 15)][kk * 16 + (lane >> 4) * 8]));
      }
      loaded = true;
    }
    // S^T = K Q^T, dP^T = V dO^T   (16 keys x 64 queries per warp)
    float s[BQ / 8][4], dp[BQ / 8][4];
#pragma unroll
    for (int n = 0; n < BQ / 8; ++n) {
      s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int n2 = 0; n2 < BQ / 16; ++n2) {
        uint32_t qb[4], ob[4];
        const int rr = n2 * 16 + (lane & 7) + ((lane >> 4) << 3), cc = kk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(qb, smem_u32(&sQ[rr][cc]));
        ldsm_x4(ob, smem_u32(&sdO[rr][cc]));
        uint32_t q0[2] = {qb[0], qb[1]}, q1[2] = {qb[2], qb[3]};
        uint32_t o0[2] = {ob[0], ob[1]}, o1[2] = {ob[2], ob[3]};
        mma_bf16_16816(s[2 * n2], kf[kk], q0);
        mma_bf16_16816(s[2 * n2 + 1], kf[kk], q1);
        mma_bf16_16816(dp[2 * n2], vf[kk], o0);
        mma_bf16_16816(dp[2 * n2 + 1], vf[kk], o1);
      }
    }
    // P^T and dS^T; rows = keys (j), cols = queries (i)
#pragma unroll
    for (int n = 0; n < BQ / 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + warp * 16 + gq + (e >> 1) * 8;
        const int il = n * 8 + 2 * tq + (e & 1);
        const int i = i0 + il;
        float pv = 0.f, ds = 0.f;
        if (j < g.nkv && i < g.nq) {
          float v = s[n][e] * p.scale;
          bool masked = false;
          if (bc.table || bc.mask) v += score_bias(bc, i, j, masked);
          if (!masked) {
            pv = __expf(v - sLse[il]);
            ds = pv * (dp[n][e] - sDelta[il]);
          }
        }
        s[n][e] = pv;
        dp[n][e] = ds;
      }
    }
    // dV += P^T dO ; dK += dS^T Q
#pragma unroll
    for (int k2 = 0; k2 < BQ / 16; ++k2) {
      uint32_t pa[4] = {pack_bf16(s[2 * k2][0], s[2 * k2][1]), pack_bf16(s[2 * k2][2], s[2 * k2][3]),
                        pack_bf16(s[2 * k2 + 1][0], s[2 * k2 + 1][1]),
                        pack_bf16(s[2 * k2 + 1][2], s[2 * k2 + 1][3])};
      uint32_t da[4] = {pack_bf16(dp[2 * k2][0], dp[2 * k2][1]), pack_bf16(dp[2 * k2][2], dp[2 * k2][3]),
                        pack_bf16(dp[2 * k2 + 1][0], dp[2 * k2 + 1][1]),
                        pack_bf16(dp[2 * k2 + 1][2], dp[2 * k2 + 1][3])};
#pragma unroll
      for (int d2 = 0; d2 < DH / 16; ++d2) {
        uint32_t ob[4], qb[4];
        const int rr = k2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, cc = d2 * 16 + (lane >> 4) * 8;
        ldsm_x4_t(ob, smem_u32(&sdO[rr][cc]));
        ldsm_x4_t(qb, smem_u32(&sQ[rr][cc]));
        uint32_t o0[2] = {ob[0], ob[1]}, o1[2] = {ob[2], ob[3]};
        uint32_t q0[2] = {qb[0], qb[1]}, q1[2] = {qb[2], qb[3]};
        mma_bf16_16816(dv[2 * d2], pa, o0);
        mma_bf16_16816(dv[2 * d2 + 1], pa, o1);
        mma_bf16_16816(dk[2 * d2], da, q0);
        mma_bf16_16816(dk[2 * d2 + 1], da, q1);
      }
    }
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = warp * 16 + gq + r * 8;
    const long tok = sKtok[row];
    if (tok < 0) continue;
    if (p.dkv_f32) {
      float* dKp = reinterpret_cast<float*>(p.dk) + tok * (long)p.lddk + h * DH;
      float* dVp = reinterpret_cast<float*>(p.dv) + tok * (long)p.lddv + h * DH;
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) {
        atomicAdd(dKp + n * 8 + 2 * tq, dk[n][2 * r] * p.scale);
        atomicAdd(dKp + n * 8 + 2 * tq + 1, dk[n][2 * r + 1] * p.scale);
        atomicAdd(dVp + n * 8 + 2 * tq, dv[n][2 * r]);
        atomicAdd(dVp + n * 8 + 2 * tq + 1, dv[n][2 * r + 1]);
      }
    } else {
      bf16* dKp = reinterpret_cast<bf16*>(p.dk) + tok * (long)p.lddk + h * DH;
      bf16* dVp = reinterpret_cast<bf16*>(p.dv) + tok * (long)p.lddv + h * DH;
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) {
        *reinterpret_cast<uint32_t*>(dKp + n * 8 + 2 * tq) =
            pack_bf16(dk[n][2 * r] * p.scale, dk[n][2 * r + 1] * p.scale);
        *reinterpret_cast<uint32_t*>(dVp + n * 8 + 2 * tq) = pack_bf16(dv[n][2 * r], dv[n][2 * r + 1]);
      }
    }
  }
}

int make_geom(const vtb_attn_params* p, Geom* g, long* groups, const char* who) {
  VTB_CHECK(p != nullptr, -1, "%s: null params", who);
  VTB_CHECK(p->dh == 32 || p->dh == 64, -1, "%s: dh must be 32 or 64 (got %d)", who, p->dh);
  VTB_CHECK(p->batch > 0 && p->heads > 0 && p->nq > 0 && p->nkv > 0, -1, "%s: bad sizes", who);
  g->mode = p->mode; g->batch = p->batch; g->heads = p->heads; g->dh = p->dh;
  g->nq = p->nq; g->nkv = p->nkv; g->Hs = p->Hs; g->Ws = p->Ws;
  g->window = p->window; g->shift = p->shift; g->halo = p->halo;
  g->nwx = 1; g->nw = 1;
  if (p->mode == VTB_ATTN_GLOBAL) {
    *groups = p->batch;
  } else {
    VTB_CHECK(p->mode == VTB_ATTN_WINDOW || p->mode == VTB_ATTN_HALO, -1, "%s: bad mode", who);
    VTB_CHECK(p->window > 0 && p->Hs % p->window == 0 && p->Ws % p->window == 0, -1,
              "%s: window %d must divide %dx%d", who, p->window, p->Hs, p->Ws);
    g->nwx = p->Ws / p->window;
    g->nw = (p->Hs / p->window) * g->nwx;
    *groups = (long)p->batch * g->nw;
    VTB_CHECK(p->nq == p->window * p->window, -1, "%s: nq must be window^2", who);
    const int kk = p->window + 2 * (p->mode == VTB_ATTN_HALO ? p->halo : 0);
    VTB_CHECK(p->nkv == kk * kk, -1, "%s: nkv must be %d", who, kk * kk);
    VTB_CHECK(p->shift >= 0 && p->shift < p->window, -1, "%s: bad shift", who);
  }
  VTB_CHECK(p->q && p->k && p->v, -1, "%s: null q/k/v", who);
  VTB_CHECK(p->ldq % 8 == 0 && p->ldk % 8 == 0 && p->ldv % 8 == 0 &&
                (((uintptr_t)p->q | (uintptr_t)p->k | (uintptr_t)p->v) & 15) == 0,
            -1, "%s: q/k/v must be 16-byte aligned rows", who);
  VTB_CHECK((p->rel_bias == nullptr) == (p->pos == nullptr), -1, "%s: rel_bias and pos go together", who);
  VTB_CHECK(!p->mask || p->n_mask > 0, -1, "%s: n_mask", who);
  VTB_CHECK(!p->mask || p->mask_ld == 0 || p->mask_ld >= p->nkv, -1, "%s: mask_ld", who);
  VTB_CHECK(!p->rel_bias || (p->n_pos > 0 && p->n_pos <= MAX_POS), -1,
            "%s: rel_bias rows n_pos=%d must be in (0, %d]", who, p->n_pos, MAX_POS);
  return 0;
}

}  // namespace

bool vtb_attn_halo_dkv_ok(const vtb_attn_params* p);
int vtb_attn_halo_dkv(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream);
bool vtb_attn_wp_ok(const vtb_attn_params* p, bool bwd);
int vtb_attn_wp_fwd(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream);
int vtb_attn_wp_bwd(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream);
bool vtb_attn_tc_fwd_ok(const vtb_attn_params* p);
int vtb_attn_tc_fwd(const vtb_attn_params* p, cudaStream_t stream);
bool vtb_attn_tc_bwd_ok(const vtb_attn_params* p);
int vtb_attn_tc_bwd(const vtb_attn_params* p, cudaStream_t stream);
bool vtb_attn_wt_ok(const vtb_attn_params* p, bool bwd);
int vtb_attn_wt_fwd(const vtb_attn_params* p, cudaStream_t stream);
int vtb_attn_wt_bwd(const vtb_attn_params* p, cudaStream_t stream);
bool vtb_attn_ht_ok(const vtb_attn_params* p, bool bwd);
size_t vtb_attn_ht_ws_bytes(const vtb_attn_params* p);
int vtb_attn_ht_fwd(const vtb_attn_params* p, cudaStream_t stream);
int vtb_attn_ht_bwd(const vtb_attn_params* p, cudaStream_t stream);
bool vtb_attn_resident_ok(const vtb_attn_params* p);
int vtb_attn_resident_fwd(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream);
int vtb_attn_resident_bwd(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream, int skip_dkv = 0);

extern "C" int vtb_attention_fwd(const vtb_attn_params* p, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Geom g;
  long groups;
  int rc = make_geom(p, &g, &groups, "vtb_attention_fwd");
  if (rc) return rc;
  VTB_CHECK(p->o && p->ldo % 2 == 0, -1, "vtb_attention_fwd: o");
  if (vtb_attn_tc_fwd_ok(p)) return vtb_attn_tc_fwd(p, stream);  // tcgen05 / TMEM path (global, dh 64, <= 256 keys)
  if (vtb_attn_wt_ok(p, false)) return vtb_attn_wt_fwd(p, stream);  // tcgen05 window tiles (two windows per 128-row tile, dh 32)
  if (vtb_attn_ht_ok(p, false)) return vtb_attn_ht_fwd(p, stream);  // tcgen05 halo tiles (two blocks per 128-row tile, dh 32)
  if (vtb_attn_wp_ok(p, false)) return vtb_attn_wp_fwd(p, g, groups, stream);  // one warp per (window, head)
  if (vtb_attn_resident_ok(p)) return vtb_attn_resident_fwd(p, g, groups, stream);
  const int q_tiles = (p->nq + BQ - 1) / BQ;
  const long blocks = groups * p->heads * q_tiles;
  VTB_CHECK(blocks < (1L << 31), -1, "vtb_attention_fwd: grid too large");
  if (p->dh == 64) attn_fwd_kernel<64><<<(unsigned)blocks, NTHREADS, 0, stream>>>(*p, g, q_tiles);
  else             attn_fwd_kernel<32><<<(unsigned)blocks, NTHREADS, 0, stream>>>(*p, g, q_tiles);
  VTB_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t vtb_attention_bwd_workspace_bytes(const vtb_attn_params* p) {
  if (p == nullptr || p->mode != VTB_ATTN_HALO || p->dh != 32 || p->window <= 0 || p->dkv_f32) return 0;
  vtb_attn_params q = *p;  // would the tcgen05 halo kernels take this problem if it had a workspace?
  q.ws = reinterpret_cast<void*>(16);
  q.ws_bytes = INT64_MAX;
  q.dout = q.o; q.dq = q.o; q.dk = q.o; q.dv = q.o;  // alignment of the gradient buffers is checked at the call
  q.lddo = q.lddq = q.lddk = q.lddv = 8;
  return vtb_attn_ht_ok(&q, true) ? (int64_t)vtb_attn_ht_ws_bytes(p) : 0;
}

extern "C" int vtb_attention_bwd(const vtb_attn_params* p, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Geom g;
  long groups;
  int rc = make_geom(p, &g, &groups, "vtb_attention_bwd");
  if (rc) return rc;
  VTB_CHECK(p->o && p->dout && p->dq && p->dk && p->dv && p->lse && p->delta, -1,
            "vtb_attention_bwd: null pointer");
  VTB_CHECK(p->lddo % 8 == 0 && p->ldo % 8 == 0 && (((uintptr_t)p->dout | (uintptr_t)p->o) & 15) == 0,
            -1, "vtb_attention_bwd: o/dout alignment");
  VTB_CHECK(p->lddq % 2 == 0 && (p->dkv_f32 || (p->lddk % 2 == 0 && p->lddv % 2 == 0)), -1,
            "vtb_attention_bwd: dq/dk/dv leading dims");
  const vtb_attn_params& q = *p;
  if (vtb_attn_tc_bwd_ok(p)) return vtb_attn_tc_bwd(p, stream);  // tcgen05 / TMEM path
  if (vtb_attn_wt_ok(p, true)) return vtb_attn_wt_bwd(p, stream);
  if (vtb_attn_ht_ok(p, true)) return vtb_attn_ht_bwd(p, stream);
  if (vtb_attn_wp_ok(p, true)) return vtb_attn_wp_bwd(p, g, groups, stream);
  if (vtb_attn_resident_ok(p) && vtb_attn_halo_dkv_ok(p)) {
    // halo: dQ + bias gradient query-centric (resident kernel, phase A only), dK / dV key-centric without atomics
    int rc2 = vtb_attn_resident_bwd(p, g, groups, stream, 1);
    if (rc2) return rc2;
    return vtb_attn_halo_dkv(p, g, groups, stream);
  }
  if (vtb_attn_resident_ok(p)) return vtb_attn_resident_bwd(p, g, groups, stream);
  const int q_tiles = (p->nq + BQ - 1) / BQ;
  const int kv_tiles = (p->nkv + BKV - 1) / BKV;
  const long b1 = groups * p->heads * q_tiles, b2 = groups * p->heads * kv_tiles;
  VTB_CHECK(b1 < (1L << 31) && b2 < (1L << 31), -1, "vtb_attention_bwd: grid too large");
  if (p->dh == 64) {
    attn_bwd_dq_kernel<64><<<(unsigned)b1, NTHREADS, 0, stream>>>(q, g, q_tiles);
    VTB_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<64><<<(unsigned)b2, NTHREADS, 0, stream>>>(q, g, kv_tiles);
  } else {
    attn_bwd_dq_kernel<32><<<(unsigned)b1, NTHREADS, 0, stream>>>(q, g, q_tiles);
    VTB_LAUNCH_CHECK();
    attn_bwd_dkv_kernel<32><<<(unsigned)b2, NTHREADS, 0, stream>>>(q, g, kv_tiles);
  }
  VTB_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// "Resident" attention (nq <= 256 and nkv <= 256: ViT/DeiT global, Swin/Twins windows, Halo, PVT stage 4).
// One CTA per (group, head); Q/K/V (and dO) are loaded ONCE into XOR-swizzled shared memory, the combined
// relative-position-bias + mask tile is built once per CTA, and each warp owns one 16-row tile: there is
// no block barrier inside the compute loops and no tile is padded beyond the next multiple of 16.
// Backward is ONE kernel: phase A (warp = 16 queries) produces dQ and the bias gradient, phase B
// (warp = 16 keys) produces dK and dV from the same resident operands.
// =====================================================================================================
namespace {

constexpr int RES_MAX = 208;      // 13 warps x 16 rows
constexpr int RES_THREADS = 416;

// element offset of 16-byte chunk `c` of row `r` in a [rows][DH] bf16 tile, XOR-swizzled so that ldmatrix
// (8 rows x 16 B) and the cooperative 16-byte stores are bank-conflict free without padding.
template <int DH>
__device__ __forceinline__ int soff(int r, int c) {
  if (DH == 64) return r * 64 + ((c ^ (r & 7)) << 3);
  return r * 32 + ((c ^ ((r >> 1) & 3)) << 3);
}

template <int DH>
__device__ __forceinline__ void res_load_rows(bf16* dst, int rows16, const bf16* base, int ld, int head_off,
                                              const int* toks) {
  constexpr int CH = DH / 8;
  for (int c = threadIdx.x; c < rows16 * CH; c += blockDim.x) {
    const int r = c / CH, cc = c - r * CH;
    const int t = toks[r];
    const bf16* src = base + (long)(t < 0 ? 0 : t) * ld + head_off + cc * 8;
    cp_async16(smem_u32(dst + soff<DH>(r, cc)), src, t >= 0);
  }
}

struct ResSmem {
  bf16 *q, *k, *v, *dO;
  float *bias, *lse, *delta, *dtab, *tab;
  float* dsum;   // [nq][bias_ld] (bwd): running sum of dS over every group this CTA processes (bias gradient)
  uint8_t* maskb;
  int *qtok, *ktok;
  int nq16, nkv16, bias_ld;
};

// nq_loc = query rows staged by this CTA (forward may split the queries of a (group, head) over CTAs)
template <int DH>
__device__ __forceinline__ ResSmem res_carve(uint8_t* base, int nq_loc, int nq, int nkv, bool bwd, bool has_bias,
                                             int n_pos) {
  ResSmem s;
  s.nq16 = (nq_loc + 15) & ~15;
  s.nkv16 = (nkv + 15) & ~15;
  s.bias_ld = (nkv + 3) & ~3;
  uint8_t* p = base;
  s.q = reinterpret_cast<bf16*>(p); p += s.nq16 * DH * 2;
  s.k = reinterpret_cast<bf16*>(p); p += s.nkv16 * DH * 2;
  s.v = reinterpret_cast<bf16*>(p); p += s.nkv16 * DH * 2;
  s.dO = reinterpret_cast<bf16*>(p); if (bwd) p += s.nq16 * DH * 2;
  s.bias = reinterpret_cast<float*>(p); if (has_bias) p += (size_t)nq * s.bias_ld * 4;
  s.lse = reinterpret_cast<float*>(p); if (bwd) p += s.nq16 * 4;
  s.delta = reinterpret_cast<float*>(p); if (bwd) p += s.nq16 * 4;
  s.dtab = reinterpret_cast<float*>(p); if (bwd && has_bias) p += n_pos * 4;
  s.tab = reinterpret_cast<float*>(p); if (has_bias) p += n_pos * 4;   // rel_bias[:, h] staged once per CTA
  s.qtok = reinterpret_cast<int*>(p); p += s.nq16 * 4;
  s.ktok = reinterpret_cast<int*>(p); p += s.nkv16 * 4;
  s.dsum = reinterpret_cast<float*>(p); if (bwd && has_bias) p += (size_t)nq * s.bias_ld * 4;
  s.maskb = p;                                                          // [nq][bias_ld] u8 mask of the current group
  return s;
}

size_t res_smem_bytes(int DH, int nq_loc, int nq, int nkv, bool bwd, bool has_bias, bool has_table, int n_pos) {
  const int nq16 = (nq_loc + 15) & ~15, nkv16 = (nkv + 15) & ~15, bld = (nkv + 3) & ~3;
  size_t b = (size_t)(nq16 + 2 * nkv16) * DH * 2 + (size_t)(nq16 + nkv16) * 4;
  if (bwd) b += (size_t)nq16 * DH * 2 + (size_t)nq16 * 8;
  if (has_bias) b += (size_t)nq * bld * 4 + (size_t)n_pos * 4 + (size_t)nq * bld;  // bias tile, table, mask bytes
  if (bwd && has_table) b += (size_t)n_pos * 4 + (size_t)nq * bld * 4;
  return b + 16;
}

// Once per CTA: stage rel_bias[:, h] in shared memory, then build the group-independent bias tile
// bias[i][j] = table[pos[i][j]] (and clear the backward's dS accumulation tile).  The table goes through
// smem first so that no global load depends on another one.
template <int DH>
__device__ __forceinline__ void res_once(const vtb_attn_params& p, const Geom& g, int h, const ResSmem& s, bool bwd) {
  const bool has_bias = p.rel_bias || p.mask;
  if (!has_bias) return;
  if (p.rel_bias)
    for (int t = threadIdx.x; t < p.n_pos; t += blockDim.x) {
      s.tab[t] = __ldg(p.rel_bias + (long)t * g.heads + h);
      if (bwd) s.dtab[t] = 0.f;
    }
  __syncthreads();
  for (int e = threadIdx.x; e < g.nq * g.nkv; e += blockDim.x) {
    const int i = e / g.nkv, j = e - i * g.nkv;
    float b = 0.f;
    if (p.rel_bias) {
      const int pi = __ldg(p.pos + e);
      b = s.tab[pi];
      if (bwd) s.dsum[i * s.bias_ld + j] = 0.f;
    }
    s.bias[i * s.bias_ld + j] = b;
  }
}

// Per group: token indices, the group's mask bytes, and the cp.async row loads.
template <int DH>
__device__ __forceinline__ void res_group(const vtb_attn_params& p, const Geom& g, int grp, int h,
                                          const ResSmem& s, bool bwd, int qbase, int q_end) {
  for (int i = threadIdx.x; i < s.nq16; i += blockDim.x)
    s.qtok[i] = (qbase + i < q_end) ? (int)q_token(g, grp, qbase + i) : -1;
  for (int j = threadIdx.x; j < s.nkv16; j += blockDim.x) s.ktok[j] = (int)kv_token(g, grp, j);
  if (p.mask) {
    const int mld = p.mask_ld > 0 ? p.mask_ld : g.nkv;
    const uint8_t* mask = p.mask + (long)(grp % p.n_mask) * g.nq * mld;
    for (int e = threadIdx.x; e < g.nq * g.nkv; e += blockDim.x) {
      const int i = e / g.nkv, j = e - i * g.nkv;
      s.maskb[i * s.bias_ld + j] = mask[(long)i * mld + j];
    }
  }
  __syncthreads();  // token indices visible
  res_load_rows<DH>(s.q, s.nq16, reinterpret_cast<const bf16*>(p.q), p.ldq, h * DH, s.qtok);
  res_load_rows<DH>(s.k, s.nkv16, reinterpret_cast<const bf16*>(p.k), p.ldk, h * DH, s.ktok);
  res_load_rows<DH>(s.v, s.nkv16, reinterpret_cast<const bf16*>(p.v), p.ldv, h * DH, s.ktok);
  if (bwd) res_load_rows<DH>(s.dO, s.nq16, reinterpret_cast<const bf16*>(p.dout), p.lddo, h * DH, s.qtok);
  cp_async_commit();
}

// bias (+ -inf where masked) of element (i, j)
__device__ __forceinline__ float res_bias(const ResSmem& s, bool has_mask, int i, int j) {
  const int o = i * s.bias_ld + j;
  return (has_mask && s.maskb[o]) ? -INFINITY : s.bias[o];
}

// A fragments (16 rows x DH) of a swizzled tile
template <int DH>
__device__ __forceinline__ void res_ld_a(uint32_t (&f)[DH / 16][4], const bf16* tile, int r0, int lane) {
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk)
    ldsm_x4(f[kk], smem_u32(tile + soff<DH>(r0 + (lane & 15), kk * 2 + (lane >> 4))));
}
// B fragments for two adjacent n8 tiles (16 "n" rows starting at n0), contraction chunk kk (16 wide), non-transposed
template <int DH>
__device__ __forceinline__ void res_ld_b(uint32_t (&r)[4], const bf16* tile, int n0, int kk, int lane) {
  ldsm_x4(r, smem_u32(tile + soff<DH>(n0 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1))));
}
// transposed B fragments: contraction rows k0..k0+15 of the tile, output columns d2*16..+15
template <int DH>
__device__ __forceinline__ void res_ld_bt(uint32_t (&r)[4], const bf16* tile, int k0, int d2, int lane) {
  ldsm_x4_t(r, smem_u32(tile + soff<DH>(k0 + (lane & 7) + (((lane >> 3) & 1) << 3), d2 * 2 + (lane >> 4))));
}

template <int DH>
__global__ void __launch_bounds__(RES_THREADS, 1)
attn_res_fwd_kernel(vtb_attn_params p, Geom g, int nsplit, int q_per_cta, int groups, int nchunks) {
  extern __shared__ __align__(16) uint8_t res_smem[];
  // blockIdx -> (query split, head, chunk); the CTA then walks groups chunk, chunk + nchunks, ... (persistent:
  // the per-head bias table / tile is staged once per CTA)
  const int split = blockIdx.x % nsplit;
  const int gh = blockIdx.x / nsplit;
  const int h = gh % g.heads;
  const int chunk = gh / g.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const bool has_bias = p.rel_bias || p.mask;
  const bool has_mask = p.mask != nullptr;
  const int qbase = split * q_per_cta;
  const int q_end = min(g.nq, qbase + q_per_cta);
  const ResSmem s = res_carve<DH>(res_smem, q_per_cta, g.nq, g.nkv, false, has_bias, p.n_pos);
  res_once<DH>(p, g, h, s, false);
  const int r0 = warp * 16;
  for (int grp = chunk; grp < groups; grp += nchunks) {
  __syncthreads();  // previous group's shared-memory reads are done (and the one-time tiles are visible)
  res_group<DH>(p, g, grp, h, s, false, qbase, q_end);
  cp_async_wait<0>();
  __syncthreads();
  if (qbase + r0 < q_end) {
  uint32_t qf[DH / 16][4];
  res_ld_a<DH>(qf, s.q, r0, lane);
  float o[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int j0 = 0; j0 < g.nkv; j0 += 64) {
    const int npair = min(4, (s.nkv16 - j0) >> 4);  // valid 16-key groups in this block
    float sc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int n2 = 0; n2 < 4; ++n2) {
        if (n2 < npair) {
          uint32_t kb[4];
          res_ld_b<DH>(kb, s.k, j0 + n2 * 16, kk, lane);
          uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
          mma_bf16_16816(sc[2 * n2], qf[kk], b0);
          mma_bf16_16816(sc[2 * n2 + 1], qf[kk], b1);
        }
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = qbase + r0 + gq + (e >> 1) * 8;
        const int j = j0 + n * 8 + 2 * tq + (e & 1);
        float v = sc[n][e] * p.scale;
        if (j >= g.nkv) v = -INFINITY;
        else if (has_bias) v += res_bias(s, has_mask, min(i, g.nq - 1), j);
        sc[n][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[r] = __expf(m_run[r] - m_use[r]);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = __expf(sc[n][e] - m_use[e >> 1]);
        sc[n][e] = pv;
        rs[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      o[n][0] *= corr[0]; o[n][1] *= corr[0]; o[n][2] *= corr[1]; o[n][3] *= corr[1];
    }
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      if (k2 < npair) {
        uint32_t pa[4] = {pack_bf16(sc[2 * k2][0], sc[2 * k2][1]), pack_bf16(sc[2 * k2][2], sc[2 * k2][3]),
                          pack_bf16(sc[2 * k2 + 1][0], sc[2 * k2 + 1][1]),
                          pack_bf16(sc[2 * k2 + 1][2], sc[2 * k2 + 1][3])};
#pragma unroll
        for (int d2 = 0; d2 < DH / 16; ++d2) {
          uint32_t vb[4];
          res_ld_bt<DH>(vb, s.v, j0 + k2 * 16, d2, lane);
          uint32_t b0[2] = {vb[0], vb[1]}, b1[2] = {vb[2], vb[3]};
          mma_bf16_16816(o[2 * d2], pa, b0);
          mma_bf16_16816(o[2 * d2 + 1], pa, b1);
        }
      }
    }
  }

  bf16* O = reinterpret_cast<bf16*>(p.o);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = r0 + gq + r * 8;
    const int tok = s.qtok[row];
    if (tok < 0) continue;
    const float inv = 1.f / l_run[r];
    bf16* dst = O + (long)tok * p.ldo + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n)
      *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * tq) = pack_bf16(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
    if (tq == 0 && p.lse) p.lse[((long)grp * g.heads + h) * g.nq + qbase + row] = m_run[r] + __logf(l_run[r]);
  }
  }  // active warp
  }  // groups
}

template <int DH>
__global__ void __launch_bounds__(RES_THREADS, 1)
attn_res_bwd_kernel(vtb_attn_params p, Geom g, int groups, int nchunks, int skip_dkv) {
  extern __shared__ __align__(16) uint8_t res_smem[];
  const int h = blockIdx.x % g.heads;
  const int chunk = blockIdx.x / g.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const bool has_bias = p.rel_bias || p.mask;
  const bool has_tab = p.rel_bias != nullptr;
  const bool has_mask = p.mask != nullptr;
  const ResSmem s = res_carve<DH>(res_smem, g.nq, g.nq, g.nkv, true, has_bias, has_tab ? p.n_pos : 0);
  res_once<DH>(p, g, h, s, true);
  for (int grp = chunk; grp < groups; grp += nchunks) {
  __syncthreads();  // previous group's shared-memory reads are done (and the one-time tiles are visible)
  res_group<DH>(p, g, grp, h, s, true, 0, g.nq);
  // delta_i = sum_d dO[i,d] O[i,d] and lse_i, two threads per row (straight from global; O is not staged)
  {
    const bf16* O = reinterpret_cast<const bf16*>(p.o);
    const bf16* dO = reinterpret_cast<const bf16*>(p.dout);
    for (int t = threadIdx.x; t < 2 * s.nq16; t += blockDim.x) {
      const int row = t >> 1, half = t & 1;
      const int tok = s.qtok[row];
      float acc = 0.f;
      if (tok >= 0) {
        const bf16* a = dO + (long)tok * p.lddo + h * DH + half * (DH / 2);
        const bf16* b = O + (long)tok * p.ldo + h * DH + half * (DH / 2);
#pragma unroll
        for (int d = 0; d < DH / 2; d += 8) {
          const uint4 ra = *reinterpret_cast<const uint4*>(a + d);
          const uint4 rb = *reinterpret_cast<const uint4*>(b + d);
          const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 fa = unpack_bf16(wa[q]), fb = unpack_bf16(wb[q]);
            acc += fa.x * fb.x + fa.y * fb.y;
          }
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);  // partner thread t^1 is in the same warp
      if (half == 0) {
        s.delta[row] = acc;
        s.lse[row] = tok >= 0 ? p.lse[((long)grp * g.heads + h) * g.nq + row] : 0.f;
        // also kept in the caller's workspace: the key-centric Halo dK / dV kernel that follows reads it from there
        if (tok >= 0 && p.delta) p.delta[((long)grp * g.heads + h) * g.nq + row] = acc;
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ------------------------------------------------------------------ phase A: dQ (+ bias gradient)
  {
    const int r0 = warp * 16;
    if (r0 < s.nq16) {
      float lse_r[2], dl_r[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) { lse_r[r] = s.lse[r0 + gq + r * 8]; dl_r[r] = s.delta[r0 + gq + r * 8]; }
      float dq[DH / 8][4];
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
      for (int j0 = 0; j0 < g.nkv; j0 += 64) {
        const int npair = min(4, (s.nkv16 - j0) >> 4);
        float sc[8][4], dp[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
          dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          // A fragments are re-read from shared memory per k-step (keeps the kernel under 128 registers)
          uint32_t qf[4], dof[4];
          ldsm_x4(qf, smem_u32(s.q + soff<DH>(r0 + (lane & 15), kk * 2 + (lane >> 4))));
          ldsm_x4(dof, smem_u32(s.dO + soff<DH>(r0 + (lane & 15), kk * 2 + (lane >> 4))));
#pragma unroll
          for (int n2 = 0; n2 < 4; ++n2) {
            if (n2 < npair) {
              uint32_t kb[4], vb[4];
              res_ld_b<DH>(kb, s.k, j0 + n2 * 16, kk, lane);
              res_ld_b<DH>(vb, s.v, j0 + n2 * 16, kk, lane);
              uint32_t k0[2] = {kb[0], kb[1]}, k1[2] = {kb[2], kb[3]};
              uint32_t v0[2] = {vb[0], vb[1]}, v1[2] = {vb[2], vb[3]};
              mma_bf16_16816(sc[2 * n2], qf, k0);
              mma_bf16_16816(sc[2 * n2 + 1], qf, k1);
              mma_bf16_16816(dp[2 * n2], dof, v0);
              mma_bf16_16816(dp[2 * n2 + 1], dof, v1);
            }
          }
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = e >> 1;
            const int i = r0 + gq + r * 8;
            const int j = j0 + n * 8 + 2 * tq + (e & 1);
            float ds = 0.f;
            if (j < g.nkv && i < g.nq) {
              float v = sc[n][e] * p.scale;
              if (has_bias) v += res_bias(s, has_mask, i, j);
              const float pv = __expf(v - lse_r[r]);  // exp(-inf) = 0 for masked entries
              ds = pv * (dp[n][e] - dl_r[r]);
              // bias gradient: element (i, j) belongs to exactly this thread in every group the CTA processes, so the
              // running sum is a plain read-modify-write (shared-memory float atomics are CAS loops: they were most of
              // this kernel's time on Halo blocks); table indices are applied once per CTA after the last group
              if (has_tab && p.drel_bias) s.dsum[i * s.bias_ld + j] += ds;
            }
            sc[n][e] = ds;
          }
        }
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          if (k2 < npair) {
            uint32_t pa[4] = {pack_bf16(sc[2 * k2][0], sc[2 * k2][1]), pack_bf16(sc[2 * k2][2], sc[2 * k2][3]),
                              pack_bf16(sc[2 * k2 + 1][0], sc[2 * k2 + 1][1]),
                              pack_bf16(sc[2 * k2 + 1][2], sc[2 * k2 + 1][3])};
#pragma unroll
            for (int d2 = 0; d2 < DH / 16; ++d2) {
              uint32_t kb[4];
              res_ld_bt<DH>(kb, s.k, j0 + k2 * 16, d2, lane);
              uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
              mma_bf16_16816(dq[2 * d2], pa, b0);
              mma_bf16_16816(dq[2 * d2 + 1], pa, b1);
            }
          }
        }
      }
      bf16* dQ = reinterpret_cast<bf16*>(p.dq);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int tok = s.qtok[r0 + gq + r * 8];
        if (tok < 0) continue;
        bf16* dst = dQ + (long)tok * p.lddq + h * DH;
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * tq) =
              pack_bf16(dq[n][2 * r] * p.scale, dq[n][2 * r + 1] * p.scale);
      }
    }
  }

  // ------------------------------------------------------------------ phase B: dK, dV (rows = keys)
  if (!skip_dkv) {
    const int c0 = warp * 16;
    if (c0 < s.nkv16) {
      constexpr int PB = (DH == 64) ? 2 : 4;  // 16-query groups per pass (register budget: 128/thread)
      float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) {
        dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
        dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
      }
      for (int i0 = 0; i0 < g.nq; i0 += 16 * PB) {
        const int npair = min(PB, (s.nq16 - i0) >> 4);
        float sc[2 * PB][4], dp[2 * PB][4];
#pragma unroll
        for (int n = 0; n < 2 * PB; ++n) {
          sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
          dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          uint32_t kf[4], vf[4];
          ldsm_x4(kf, smem_u32(s.k + soff<DH>(c0 + (lane & 15), kk * 2 + (lane >> 4))));
          ldsm_x4(vf, smem_u32(s.v + soff<DH>(c0 + (lane & 15), kk * 2 + (lane >> 4))));
#pragma unroll
          for (int n2 = 0; n2 < PB; ++n2) {
            if (n2 < npair) {
              uint32_t qb[4], ob[4];
              res_ld_b<DH>(qb, s.q, i0 + n2 * 16, kk, lane);
              res_ld_b<DH>(ob, s.dO, i0 + n2 * 16, kk, lane);
              uint32_t q0[2] = {qb[0], qb[1]}, q1[2] = {qb[2], qb[3]};
              uint32_t o0[2] = {ob[0], ob[1]}, o1[2] = {ob[2], ob[3]};
              mma_bf16_16816(sc[2 * n2], kf, q0);
              mma_bf16_16816(sc[2 * n2 + 1], kf, q1);
              mma_bf16_16816(dp[2 * n2], vf, o0);
              mma_bf16_16816(dp[2 * n2 + 1], vf, o1);
            }
          }
        }
#pragma unroll
        for (int n = 0; n < 2 * PB; ++n) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = c0 + gq + (e >> 1) * 8;
            const int i = i0 + n * 8 + 2 * tq + (e & 1);
            float pv = 0.f, ds = 0.f;
            if (j < g.nkv && i < g.nq) {
              float v = sc[n][e] * p.scale;
              if (has_bias) v += res_bias(s, has_mask, i, j);
              pv = __expf(v - s.lse[i]);
              ds = pv * (dp[n][e] - s.delta[i]);
            }
            sc[n][e] = pv;
            dp[n][e] = ds;
          }
        }
#pragma unroll
        for (int k2 = 0; k2 < PB; ++k2) {
          if (k2 < npair) {
            uint32_t pa[4] = {pack_bf16(sc[2 * k2][0], sc[2 * k2][1]), pack_bf16(sc[2 * k2][2], sc[2 * k2][3]),
                              pack_bf16(sc[2 * k2 + 1][0], sc[2 * k2 + 1][1]),
                              pack_bf16(sc[2 * k2 + 1][2], sc[2 * k2 + 1][3])};
            uint32_t da[4] = {pack_bf16(dp[2 * k2][0], dp[2 * k2][1]), pack_bf16(dp[2 * k2][2], dp[2 * k2][3]),
                              pack_bf16(dp[2 * k2 + 1][0], dp[2 * k2 + 1][1]),
                              pack_bf16(dp[2 * k2 + 1][2], dp[2 * k2 + 1][3])};
#pragma unroll
            for (int d2 = 0; d2 < DH / 16; ++d2) {
              uint32_t ob[4], qb[4];
              res_ld_bt<DH>(ob, s.dO, i0 + k2 * 16, d2, lane);
              res_ld_bt<DH>(qb, s.q, i0 + k2 * 16, d2, lane);
              uint32_t o0[2] = {ob[0], ob[1]}, o1[2] = {ob[2], ob[3]};
              uint32_t q0[2] = {qb[0], qb[1]}, q1[2] = {qb[2], qb[3]};
              mma_bf16_16816(dv[2 * d2], pa, o0);
              mma_bf16_16816(dv[2 * d2 + 1], pa, o1);
              mma_bf16_16816(dk[2 * d2], da, q0);
              mma_bf16_16816(dk[2 * d2 + 1], da, q1);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int tok = s.ktok[c0 + gq + r * 8];
        if (tok < 0) continue;
        if (p.dkv_f32) {
          float* dKp = reinterpret_cast<float*>(p.dk) + (long)tok * p.lddk + h * DH;
          float* dVp = reinterpret_cast<float*>(p.dv) + (long)tok * p.lddv + h * DH;
#pragma unroll
          for (int n = 0; n < DH / 8; ++n) {
            atomicAdd(dKp + n * 8 + 2 * tq, dk[n][2 * r] * p.scale);
            atomicAdd(dKp + n * 8 + 2 * tq + 1, dk[n][2 * r + 1] * p.scale);
            atomicAdd(dVp + n * 8 + 2 * tq, dv[n][2 * r]);
            atomicAdd(dVp + n * 8 + 2 * tq + 1, dv[n][2 * r + 1]);
          }
        } else {
          bf16* dKp = reinterpret_cast<bf16*>(p.dk) + (long)tok * p.lddk + h * DH;
          bf16* dVp = reinterpret_cast<bf16*>(p.dv) + (long)tok * p.lddv + h * DH;
#pragma unroll
          for (int n = 0; n < DH / 8; ++n) {
            *reinterpret_cast<uint32_t*>(dKp + n * 8 + 2 * tq) =
                pack_bf16(dk[n][2 * r] * p.scale, dk[n][2 * r + 1] * p.scale);
            *reinterpret_cast<uint32_t*>(dVp + n * 8 + 2 * tq) = pack_bf16(dv[n][2 * r], dv[n][2 * r + 1]);
          }
        }
      }
    }
  }

  }  // groups: the dS tile keeps accumulating across them; folded into the table and flushed once per CTA
  if (has_tab && p.drel_bias) {
    __syncthreads();
    for (int e = threadIdx.x; e < g.nq * g.nkv; e += blockDim.x) {
      const int i = e / g.nkv, j = e - i * g.nkv;
      const float v = s.dsum[i * s.bias_ld + j];
      if (v != 0.f) atomicAdd(&s.dtab[__ldg(p.pos + e)], v);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < p.n_pos; t += blockDim.x) {
      const float v = s.dtab[t];
      if (v != 0.f) atomicAdd(p.drel_bias + (long)t * g.heads + h, v);
    }
  }
}

template <typename K>
int res_set_smem(K kern, size_t bytes) {
  VTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  (void)bytes;
  return 0;
}

}  // namespace

bool vtb_attn_resident_ok(const vtb_attn_params* p) { return p->nq <= RES_MAX && p->nkv <= RES_MAX; }

int vtb_attn_resident_fwd(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream) {
  const bool has_bias = p->rel_bias || p->mask;
  // split the queries of one (group, head) over CTAs so that >= 2-3 CTAs fit an SM (K/V are re-read from L2)
  const int tiles = (p->nq + 15) / 16;
  const int nsplit = tiles > 8 ? 2 : 1;
  const int q_per_cta = ((tiles + nsplit - 1) / nsplit) * 16;
  const size_t smem = res_smem_bytes(p->dh, q_per_cta, p->nq, p->nkv, false, has_bias, p->rel_bias != nullptr, p->n_pos);
  VTB_CHECK(smem <= 227 * 1024, -1, "vtb_attention_fwd: resident tile needs %zu B of shared memory", smem);
  const int warps = q_per_cta / 16;
  static bool set64 = false, set32 = false;
  if (p->dh == 64 && !set64) { int rc = res_set_smem(attn_res_fwd_kernel<64>, smem); if (rc) return rc; set64 = true; }
  if (p->dh == 32 && !set32) { int rc = res_set_smem(attn_res_fwd_kernel<32>, smem); if (rc) return rc; set32 = true; }
  // persistent over groups when a bias table / mask has to be staged per CTA; one group per CTA otherwise
  long nchunks = groups;
  if (has_bias) {  // one resident wave: CTAs per SM from the occupancy calculator
    int per_sm = 1;
    if (p->dh == 64) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn_res_fwd_kernel<64>, warps * 32, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn_res_fwd_kernel<32>, warps * 32, smem);
    if (per_sm < 1) per_sm = 1;
    const long target = (long)vtb_num_sms() * per_sm / ((long)p->heads * nsplit);
    nchunks = target < 1 ? 1 : (target < groups ? target : groups);
  }
  const long blocks = nchunks * p->heads * nsplit;
  VTB_CHECK(blocks < (1L << 31) && groups < (1L << 31), -1, "vtb_attention_fwd: grid too large");
  if (p->dh == 64) {
    attn_res_fwd_kernel<64><<<(unsigned)blocks, warps * 32, smem, stream>>>(*p, g, nsplit, q_per_cta, (int)groups, (int)nchunks);
  } else {
    attn_res_fwd_kernel<32><<<(unsigned)blocks, warps * 32, smem, stream>>>(*p, g, nsplit, q_per_cta, (int)groups, (int)nchunks);
  }
  VTB_LAUNCH_CHECK();
  return 0;
}

int vtb_attn_resident_bwd(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream, int skip_dkv) {
  const bool has_bias = p->rel_bias || p->mask;
  const size_t smem = res_smem_bytes(p->dh, p->nq, p->nq, p->nkv, true, has_bias, p->rel_bias != nullptr, p->n_pos);
  VTB_CHECK(smem <= 227 * 1024, -1, "vtb_attention_bwd: resident tile needs %zu B of shared memory", smem);
  const int nmax = (skip_dkv || p->nq > p->nkv) ? p->nq : p->nkv;  // phase A only needs one warp per 16 queries
  const int warps = (nmax + 15) / 16;
  static bool set64 = false, set32 = false;
  if (p->dh == 64 && !set64) { int rc = res_set_smem(attn_res_bwd_kernel<64>, smem); if (rc) return rc; set64 = true; }
  if (p->dh == 32 && !set32) { int rc = res_set_smem(attn_res_bwd_kernel<32>, smem); if (rc) return rc; set32 = true; }
  long nchunks = groups;
  if (has_bias) {
    int per_sm = 1;
    if (p->dh == 64) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn_res_bwd_kernel<64>, warps * 32, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn_res_bwd_kernel<32>, warps * 32, smem);
    if (per_sm < 1) per_sm = 1;
    const long target = (long)vtb_num_sms() * per_sm / (long)p->heads;
    nchunks = target < 1 ? 1 : (target < groups ? target : groups);
  }
  const long blocks = nchunks * p->heads;
  VTB_CHECK(blocks < (1L << 31) && groups < (1L << 31), -1, "vtb_attention_bwd: grid too large");
  if (p->dh == 64) {
    attn_res_bwd_kernel<64><<<(unsigned)blocks, warps * 32, smem, stream>>>(*p, g, (int)groups, (int)nchunks, skip_dkv);
  } else {
    attn_res_bwd_kernel<32><<<(unsigned)blocks, warps * 32, smem, stream>>>(*p, g, (int)groups, (int)nchunks, skip_dkv);
  }
  VTB_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// Warp-per-problem attention for small problems (nq <= 64 and nkv <= 64: 7x7 windows of Swin / Twins).
// With 49 tokens a (window, head) problem is only four 16-row MMA tiles; spreading it over four warps makes
// every warp pay the whole per-problem overhead (token indices, loads, barriers) for a quarter of the work
// (ncu: 11 k warp-instructions per problem, < 5 % of them MMAs).  Here ONE warp owns a problem end to end,
// walks its tiles sequentially out of its private shared-memory slice, and loops over problems with no block
// barrier at all; the four warps of a CTA share the per-head bias tile.  Mask rows arrive with 16-byte loads
// (mask_ld = padded row pitch).
// =====================================================================================================
namespace {

constexpr int WP_WARPS = 4;
constexpr int WP_LD = 64;  // pitch of the bias / mask / pos tiles (key slots padded to 64)

struct WpSmem {
  float* bias;           // [nq][64]   shared by the CTA (table[pos], -inf beyond nkv)
  unsigned short* pos;   // [nq][64]   (bwd)
  float* tab;            // [n_pos]
  float* dtab;           // [n_pos]    (bwd)
  bf16 *q, *k, *v, *dO;  // per warp [64][DH]
  uint8_t* maskb;        // per warp [nq][64]
  float *lse, *delta;    // per warp [64] (bwd)
  int *qtok, *ktok;      // per warp [64]
};

template <int DH>
size_t wp_smem_bytes(int nq, bool bwd, int n_pos) {
  size_t shared = (size_t)nq * WP_LD * 4 + (size_t)n_pos * 4 * (bwd ? 2 : 1) + (bwd ? (size_t)nq * WP_LD * 2 : 0);
  size_t per_warp = (size_t)(bwd ? 4 : 3) * 64 * DH * 2 + (size_t)nq * WP_LD + 2 * 64 * 4 + (bwd ? 2 * 64 * 4 : 0);
  return ((shared + 15) & ~(size_t)15) + WP_WARPS * ((per_warp + 15) & ~(size_t)15) + 32;
}

template <int DH>
__device__ __forceinline__ WpSmem wp_carve(uint8_t* base, int nq, bool bwd, int n_pos, int warp) {
  WpSmem s;
  uint8_t* p = base;
  s.bias = reinterpret_cast<float*>(p); p += (size_t)nq * WP_LD * 4;
  s.tab = reinterpret_cast<float*>(p); p += (size_t)n_pos * 4;
  s.dtab = reinterpret_cast<float*>(p); if (bwd) p += (size_t)n_pos * 4;
  s.pos = reinterpret_cast<unsigned short*>(p); if (bwd) p += (size_t)nq * WP_LD * 2;
  p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
  size_t per_warp = (size_t)(bwd ? 4 : 3) * 64 * DH * 2 + (size_t)nq * WP_LD + 2 * 64 * 4 + (bwd ? 2 * 64 * 4 : 0);
  per_warp = (per_warp + 15) & ~(size_t)15;
  p += (size_t)warp * per_warp;
  s.q = reinterpret_cast<bf16*>(p); p += 64 * DH * 2;
  s.k = reinterpret_cast<bf16*>(p); p += 64 * DH * 2;
  s.v = reinterpret_cast<bf16*>(p); p += 64 * DH * 2;
  s.dO = reinterpret_cast<bf16*>(p); if (bwd) p += 64 * DH * 2;
  s.maskb = p; p += (size_t)nq * WP_LD;
  s.qtok = reinterpret_cast<int*>(p); p += 64 * 4;
  s.ktok = reinterpret_cast<int*>(p); p += 64 * 4;
  s.lse = reinterpret_cast<float*>(p); if (bwd) p += 64 * 4;
  s.delta = reinterpret_cast<float*>(p);
  return s;
}

// once per CTA: per-head table -> smem, bias tile (and pos tile) with -inf padding beyond nkv
__device__ __forceinline__ void wp_once(const vtb_attn_params& p, const Geom& g, int h, const WpSmem& s, bool bwd) {
  if (p.rel_bias)
    for (int t = threadIdx.x; t < p.n_pos; t += blockDim.x) {
      s.tab[t] = __ldg(p.rel_bias + (long)t * g.heads + h);
      if (bwd) s.dtab[t] = 0.f;
    }
  __syncthreads();
  for (int e = threadIdx.x; e < g.nq * WP_LD; e += blockDim.x) {
    const int i = e / WP_LD, j = e - i * WP_LD;
    float b = (j < g.nkv) ? 0.f : -INFINITY;
    unsigned short pi = 0;
    if (p.rel_bias && j < g.nkv) {
      pi = (unsigned short)__ldg(p.pos + i * g.nkv + j);
      b = s.tab[pi];
    }
    s.bias[e] = b;
    if (bwd) s.pos[e] = pi;
  }
  __syncthreads();
}

// per problem, executed by ONE warp: token indices, mask rows, cp.async row loads
template <int DH, int NL = 32>
__device__ __forceinline__ void wp_load(const vtb_attn_params& p, const Geom& g, int grp, int h, const WpSmem& s,
                                        bool bwd, int lane, int mask_ld, int pair_bar = 0) {
  for (int i = lane; i < 64; i += NL) {
    s.qtok[i] = (int)q_token(g, grp, i);
    s.ktok[i] = (int)kv_token(g, grp, i);
  }
  if (p.mask) {
    const uint8_t* m = p.mask + (long)(grp % p.n_mask) * g.nq * mask_ld;
    if (mask_ld == WP_LD) {
      for (int c = lane; c < g.nq * (WP_LD / 16); c += NL)
        cp_async16(smem_u32(s.maskb + c * 16), m + c * 16, true);
    } else {
      for (int e = lane; e < g.nq * g.nkv; e += NL) {
        const int i = e / g.nkv, j = e - i * g.nkv;
        s.maskb[i * WP_LD + j] = m[(long)i * mask_ld + j];
      }
    }
  }
  if (NL == 32) __syncwarp();
  else asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
  constexpr int CH = DH / 8;
  for (int c = lane; c < 64 * CH; c += NL) {
    const int r = c / CH, cc = c - r * CH;
    const int tq_ = s.qtok[r], tk_ = s.ktok[r];
    const long oq = (long)(tq_ < 0 ? 0 : tq_), ok = (long)(tk_ < 0 ? 0 : tk_);
    cp_async16(smem_u32(s.q + soff<DH>(r, cc)), reinterpret_cast<const bf16*>(p.q) + oq * p.ldq + h * DH + cc * 8, tq_ >= 0);
    cp_async16(smem_u32(s.k + soff<DH>(r, cc)), reinterpret_cast<const bf16*>(p.k) + ok * p.ldk + h * DH + cc * 8, tk_ >= 0);
    cp_async16(smem_u32(s.v + soff<DH>(r, cc)), reinterpret_cast<const bf16*>(p.v) + ok * p.ldv + h * DH + cc * 8, tk_ >= 0);
    if (bwd)
      cp_async16(smem_u32(s.dO + soff<DH>(r, cc)), reinterpret_cast<const bf16*>(p.dout) + oq * p.lddo + h * DH + cc * 8, tq_ >= 0);
  }
  cp_async_commit();
}

template <int DH>
__global__ void __launch_bounds__(WP_WARPS * 32, 2)
attn_wp_fwd_kernel(vtb_attn_params p, Geom g, int groups, int nchunks, int mask_ld) {
  extern __shared__ __align__(16) uint8_t wp_smem[];
  const int h = blockIdx.x % g.heads;
  const int chunk = blockIdx.x / g.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const bool has_mask = p.mask != nullptr;
  const WpSmem s = wp_carve<DH>(wp_smem, g.nq, false, p.rel_bias ? p.n_pos : 0, warp);
  wp_once(p, g, h, s, false);
  const int mtiles = (g.nq + 15) >> 4;
  const int npair = (g.nkv + 15) >> 4;  // 16-key groups that hold real keys
  const float sl2 = p.scale * 1.4426950408889634f;
  bf16* O = reinterpret_cast<bf16*>(p.o);

  for (int grp = chunk * WP_WARPS + warp; grp < groups; grp += nchunks * WP_WARPS) {
    __syncwarp();  // previous problem's reads of this warp's slice are done
    wp_load<DH>(p, g, grp, h, s, false, lane, mask_ld);
    cp_async_wait<0>();
    __syncwarp();
    for (int mt = 0; mt < mtiles; ++mt) {
      const int r0 = mt * 16;
      float sc[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) { sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        uint32_t qf[4];
        ldsm_x4(qf, smem_u32(s.q + soff<DH>(r0 + (lane & 15), kk * 2 + (lane >> 4))));
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
          if (n2 < npair) {
            uint32_t kb[4];
            res_ld_b<DH>(kb, s.k, n2 * 16, kk, lane);
            uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
            mma_bf16_16816(sc[2 * n2], qf, b0);
            mma_bf16_16816(sc[2 * n2 + 1], qf, b1);
          }
        }
      }
      // logits in the log2 domain: s*scale*log2e + bias*log2e (bias tile already holds -inf for key padding)
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = min(r0 + gq + r * 8, g.nq - 1);
        const float* brow = s.bias + i * WP_LD + 2 * tq;
        const uint8_t* mrow = s.maskb + i * WP_LD + 2 * tq;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float v = fmaf(sc[n][2 * r + e], sl2, brow[n * 8 + e] * 1.4426950408889634f);
            if (has_mask && mrow[n * 8 + e]) v = -INFINITY;
            sc[n][2 * r + e] = v;
            mx[r] = fmaxf(mx[r], v);
          }
        }
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_use = (mx[r] == -INFINITY) ? 0.f : mx[r];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pv = exp2f(sc[n][2 * r + e] - m_use);
            sc[n][2 * r + e] = pv;
            rs[r] += pv;
          }
        }
        rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
        rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      }
      float o[DH / 8][4];
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        if (k2 < npair) {
          uint32_t pa[4] = {pack_bf16(sc[2 * k2][0], sc[2 * k2][1]), pack_bf16(sc[2 * k2][2], sc[2 * k2][3]),
                            pack_bf16(sc[2 * k2 + 1][0], sc[2 * k2 + 1][1]),
                            pack_bf16(sc[2 * k2 + 1][2], sc[2 * k2 + 1][3])};
#pragma unroll
          for (int d2 = 0; d2 < DH / 16; ++d2) {
            uint32_t vb[4];
            res_ld_bt<DH>(vb, s.v, k2 * 16, d2, lane);
            uint32_t b0[2] = {vb[0], vb[1]}, b1[2] = {vb[2], vb[3]};
            mma_bf16_16816(o[2 * d2], pa, b0);
            mma_bf16_16816(o[2 * d2 + 1], pa, b1);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = r0 + gq + r * 8;
        const int tok = s.qtok[row];
        if (tok < 0) continue;
        const float inv = 1.f / rs[r];
        bf16* dst = O + (long)tok * p.ldo + h * DH;
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * tq) = pack_bf16(o[n][2 * r] * inv, o[n][2 * r + 1] * inv);
        if (tq == 0 && p.lse)
          p.lse[((long)grp * g.heads + h) * g.nq + row] = (mx[r] + log2f(rs[r])) * 0.6931471805599453f;
      }
    }
  }
}

// Backward: TWO warps per problem share one shared-memory slice — warp role 0 runs phase A (dQ + bias gradient),
// role 1 runs phase B (dK, dV) at the same time; they meet on a 64-thread named barrier around the loads.
template <int DH>
__global__ void __launch_bounds__(WP_WARPS * 64, 2)
attn_wp_bwd_kernel(vtb_attn_params p, Geom g, int groups, int nchunks, int mask_ld) {
  extern __shared__ __align__(16) uint8_t wp_smem[];
  const int h = blockIdx.x % g.heads;
  const int chunk = blockIdx.x / g.heads;
  const int warp = (threadIdx.x >> 5) >> 1, role = (threadIdx.x >> 5) & 1, lane = threadIdx.x & 31;
  const int lane64 = role * 32 + lane;  // index inside the warp pair
  const int gq = lane >> 2, tq = lane & 3;
  const bool has_mask = p.mask != nullptr;
  const bool has_tab = p.rel_bias != nullptr && p.drel_bias != nullptr;
  const WpSmem s = wp_carve<DH>(wp_smem, g.nq, true, p.rel_bias ? p.n_pos : 0, warp);
  wp_once(p, g, h, s, true);
  const int qtiles = (g.nq + 15) >> 4, ktiles = (g.nkv + 15) >> 4;
  const float sl2 = p.scale * 1.4426950408889634f;
  constexpr float L2E = 1.4426950408889634f;
  const bf16* Og = reinterpret_cast<const bf16*>(p.o);
  const bf16* dOg = reinterpret_cast<const bf16*>(p.dout);

  for (int grp = chunk * WP_WARPS + warp; grp < groups; grp += nchunks * WP_WARPS) {
    asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");  // both warps are done with the previous problem
    wp_load<DH, 64>(p, g, grp, h, s, true, lane64, mask_ld, warp + 1);
    // delta_i = sum_d dO[i,d] O[i,d];  lse2_i = lse_i log2(e) (+inf for padding rows so that p = 0)
    {
      const int row = lane64;
      const int tok = s.qtok[row];
      float acc = 0.f, l2 = INFINITY;
      if (tok >= 0) {
        const bf16* a = dOg + (long)tok * p.lddo + h * DH;
        const bf16* b = Og + (long)tok * p.ldo + h * DH;
#pragma unroll
        for (int d = 0; d < DH; d += 8) {
          const uint4 ra = *reinterpret_cast<const uint4*>(a + d);
          const uint4 rb = *reinterpret_cast<const uint4*>(b + d);
          const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 fa = unpack_bf16(wa[q]), fb = unpack_bf16(wb[q]);
            acc += fa.x * fb.x + fa.y * fb.y;
          }
        }
        l2 = p.lse[((long)grp * g.heads + h) * g.nq + row] * L2E;
      }
      s.delta[row] = acc;
      s.lse[row] = l2;
    }
    cp_async_wait<0>();
    asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");  // tiles, lse2, delta visible to the pair

    // ---------------------------------------------------------------- phase A: dQ (+ bias gradient), rows = queries
    if (role == 0)
    for (int mt = 0; mt < qtiles; ++mt) {
      const int r0 = mt * 16;
      float sc[8][4], dp[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
        dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        uint32_t qf[4], dof[4];
        ldsm_x4(qf, smem_u32(s.q + soff<DH>(r0 + (lane & 15), kk * 2 + (lane >> 4))));
        ldsm_x4(dof, smem_u32(s.dO + soff<DH>(r0 + (lane & 15), kk * 2 + (lane >> 4))));
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
          if (n2 < ktiles) {
            uint32_t kb[4], vb[4];
            res_ld_b<DH>(kb, s.k, n2 * 16, kk, lane);
            res_ld_b<DH>(vb, s.v, n2 * 16, kk, lane);
            uint32_t k0[2] = {kb[0], kb[1]}, k1[2] = {kb[2], kb[3]};
            uint32_t v0[2] = {vb[0], vb[1]}, v1[2] = {vb[2], vb[3]};
            mma_bf16_16816(sc[2 * n2], qf, k0);
            mma_bf16_16816(sc[2 * n2 + 1], qf, k1);
            mma_bf16_16816(dp[2 * n2], dof, v0);
            mma_bf16_16816(dp[2 * n2 + 1], dof, v1);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = min(r0 + gq + r * 8, g.nq - 1);
        const float l2 = s.lse[r0 + gq + r * 8], dl = s.delta[r0 + gq + r * 8];
        const float* brow = s.bias + i * WP_LD + 2 * tq;
        const uint8_t* mrow = s.maskb + i * WP_LD + 2 * tq;
        const unsigned short* prow = s.pos + i * WP_LD + 2 * tq;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float v = fmaf(sc[n][2 * r + e], sl2, fmaf(brow[n * 8 + e], L2E, -l2));
            if (has_mask && mrow[n * 8 + e]) v = -INFINITY;
            const float ds = exp2f(v) * (dp[n][2 * r + e] - dl);  // exp2(-inf) = 0: masked / padded entries vanish
            sc[n][2 * r + e] = ds;
            if (has_tab && ds != 0.f) atomicAdd(&s.dtab[prow[n * 8 + e]], ds);
          }
        }
      }
      float dq[DH / 8][4];
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        if (k2 < ktiles) {
          uint32_t pa[4] = {pack_bf16(sc[2 * k2][0], sc[2 * k2][1]), pack_bf16(sc[2 * k2][2], sc[2 * k2][3]),
                            pack_bf16(sc[2 * k2 + 1][0], sc[2 * k2 + 1][1]),
                            pack_bf16(sc[2 * k2 + 1][2], sc[2 * k2 + 1][3])};
#pragma unroll
          for (int d2 = 0; d2 < DH / 16; ++d2) {
            uint32_t kb[4];
            res_ld_bt<DH>(kb, s.k, k2 * 16, d2, lane);
            uint32_t b0[2] = {kb[0], kb[1]}, b1[2] = {kb[2], kb[3]};
            mma_bf16_16816(dq[2 * d2], pa, b0);
            mma_bf16_16816(dq[2 * d2 + 1], pa, b1);
          }
        }
      }
      bf16* dQ = reinterpret_cast<bf16*>(p.dq);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int tok = s.qtok[r0 + gq + r * 8];
        if (tok < 0) continue;
        bf16* dst = dQ + (long)tok * p.lddq + h * DH;
#pragma unroll
        for (int n = 0; n < DH / 8; ++n)
          *reinterpret_cast<uint32_t*>(dst + n * 8 + 2 * tq) =
              pack_bf16(dq[n][2 * r] * p.scale, dq[n][2 * r + 1] * p.scale);
      }
    }

    // ---------------------------------------------------------------- phase B: dK, dV, rows = keys
    if (role == 1)
    for (int kt = 0; kt < ktiles; ++kt) {
      const int c0 = kt * 16;
      float sc[8][4], dp[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
        dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        uint32_t kf[4], vf[4];
        ldsm_x4(kf, smem_u32(s.k + soff<DH>(c0 + (lane & 15), kk * 2 + (lane >> 4))));
        ldsm_x4(vf, smem_u32(s.v + soff<DH>(c0 + (lane & 15), kk * 2 + (lane >> 4))));
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
          if (n2 < qtiles) {
            uint32_t qb[4], ob[4];
            res_ld_b<DH>(qb, s.q, n2 * 16, kk, lane);
            res_ld_b<DH>(ob, s.dO, n2 * 16, kk, lane);
            uint32_t q0[2] = {qb[0], qb[1]}, q1[2] = {qb[2], qb[3]};
            uint32_t o0[2] = {ob[0], ob[1]}, o1[2] = {ob[2], ob[3]};
            mma_bf16_16816(sc[2 * n2], kf, q0);
            mma_bf16_16816(sc[2 * n2 + 1], kf, q1);
            mma_bf16_16816(dp[2 * n2], vf, o0);
            mma_bf16_16816(dp[2 * n2 + 1], vf, o1);
          }
        }
      }
      // element (key j = c0 + gq + r*8, query i = n*8 + 2tq + e): bias / mask are indexed [i][j]
#pragma unroll
      for (int n = 0; n < 8; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = n * 8 + 2 * tq + e;
          const int ic = min(i, g.nq - 1);
          const float l2 = s.lse[i], dl = s.delta[i];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int j = c0 + gq + r * 8;
            float v = fmaf(sc[n][2 * r + e], sl2, fmaf(s.bias[ic * WP_LD + j], L2E, -l2));
            if (has_mask && s.maskb[ic * WP_LD + j]) v = -INFINITY;
            const float pv = exp2f(v);
            sc[n][2 * r + e] = pv;
            dp[n][2 * r + e] = pv * (dp[n][2 * r + e] - dl);
          }
        }
      }
      float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
      for (int n = 0; n < DH / 8; ++n) {
        dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
        dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
      }
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        if (k2 < qtiles) {
          uint32_t pa[4] = {pack_bf16(sc[2 * k2][0], sc[2 * k2][1]), pack_bf16(sc[2 * k2][2], sc[2 * k2][3]),
                            pack_bf16(sc[2 * k2 + 1][0], sc[2 * k2 + 1][1]),
                            pack_bf16(sc[2 * k2 + 1][2], sc[2 * k2 + 1][3])};
          uint32_t da[4] = {pack_bf16(dp[2 * k2][0], dp[2 * k2][1]), pack_bf16(dp[2 * k2][2], dp[2 * k2][3]),
                            pack_bf16(dp[2 * k2 + 1][0], dp[2 * k2 + 1][1]),
                            pack_bf16(dp[2 * k2 + 1][2], dp[2 * k2 + 1][3])};
#pragma unroll
          for (int d2 = 0; d2 < DH / 16; ++d2) {
            uint32_t ob[4], qb[4];
            res_ld_bt<DH>(ob, s.dO, k2 * 16, d2, lane);
            res_ld_bt<DH>(qb, s.q, k2 * 16, d2, lane);
            uint32_t o0[2] = {ob[0], ob[1]}, o1[2] = {ob[2], ob[3]};
            uint32_t q0[2] = {qb[0], qb[1]}, q1[2] = {qb[2], qb[3]};
            mma_bf16_16816(dv[2 * d2], pa, o0);
            mma_bf16_16816(dv[2 * d2 + 1], pa, o1);
            mma_bf16_16816(dk[2 * d2], da, q0);
            mma_bf16_16816(dk[2 * d2 + 1], da, q1);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int tok = s.ktok[c0 + gq + r * 8];
        if (tok < 0) continue;
        bf16* dKp = reinterpret_cast<bf16*>(p.dk) + (long)tok * p.lddk + h * DH;
        bf16* dVp = reinterpret_cast<bf16*>(p.dv) + (long)tok * p.lddv + h * DH;
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) {
          *reinterpret_cast<uint32_t*>(dKp + n * 8 + 2 * tq) =
              pack_bf16(dk[n][2 * r] * p.scale, dk[n][2 * r + 1] * p.scale);
          *reinterpret_cast<uint32_t*>(dVp + n * 8 + 2 * tq) = pack_bf16(dv[n][2 * r], dv[n][2 * r + 1]);
        }
      }
    }
  }
  if (has_tab) {
    __syncthreads();
    for (int t = threadIdx.x; t < p.n_pos; t += blockDim.x) {
      const float v = s.dtab[t];
      if (v != 0.f) atomicAdd(p.drel_bias + (long)t * g.heads + h, v);
    }
  }
}

bool g_attn_wp = true;

template <int DH, bool BWD>
int wp_launch(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream) {
  const size_t smem = wp_smem_bytes<DH>(p->nq, BWD, p->rel_bias ? p->n_pos : 0);
  VTB_CHECK(smem <= 227 * 1024, -1, "vtb_attention: warp-per-problem tile needs %zu B of shared memory", smem);
  auto kern = BWD ? attn_wp_bwd_kernel<DH> : attn_wp_fwd_kernel<DH>;
  static bool set = false;
  if (!set) {
    VTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    set = true;
  }
  const int threads = BWD ? WP_WARPS * 64 : WP_WARPS * 32;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  if (per_sm < 1) per_sm = 1;
  long nchunks = (long)vtb_num_sms() * per_sm / p->heads;
  const long need = (groups + WP_WARPS - 1) / WP_WARPS;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > need) nchunks = need;
  const long blocks = nchunks * p->heads;
  VTB_CHECK(blocks < (1L << 31) && groups < (1L << 31), -1, "vtb_attention: grid too large");
  const int mask_ld = p->mask ? (p->mask_ld > 0 ? p->mask_ld : p->nkv) : 0;
  kern<<<(unsigned)blocks, threads, smem, stream>>>(*p, g, (int)groups, (int)nchunks, mask_ld);
  VTB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

void vtb_attn_wp_set(bool on) { g_attn_wp = on; }

bool vtb_attn_wp_ok(const vtb_attn_params* p, bool bwd) {
  return g_attn_wp && p->mode != VTB_ATTN_HALO && p->nq <= 64 && p->nkv <= 64 && !(bwd && p->dkv_f32) &&
         (!p->rel_bias || p->n_pos <= MAX_POS);
}

int vtb_attn_wp_fwd(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream) {
  return p->dh == 64 ? wp_launch<64, false>(p, g, groups, stream) : wp_launch<32, false>(p, g, groups, stream);
}
int vtb_attn_wp_bwd(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream) {
  return p->dh == 64 ? wp_launch<64, true>(p, g, groups, stream) : wp_launch<32, true>(p, g, groups, stream);
}

// =====================================================================================================
// Halo attention backward, key-centric dK / dV (halo_transformer.py:74-106 adjoint).
// A key/value token lies in the (W+2h)^2 halo of up to (2*ceil(h/W)+1)^2 blocks, so the query-centric kernel has
// to scatter dK/dV with fp32 atomics (21 k atomics per (block, head)).  Here each CTA owns the W^2 tokens of ONE
// block as keys and walks the neighbouring blocks whose halo covers them as queries: for neighbour (dy, dx) the
// centre key (cy, cx) sits in slot (cy - W*dy + h, cx - W*dx + h) of that neighbour's halo window (valid when
// inside [0, W+2h)^2).  S^T, P^T, dS^T are recomputed per neighbour from the neighbour's Q / dO / O / lse;
// dK, dV accumulate in registers and are written once, in bf16, without atomics.
// =====================================================================================================
namespace {

template <int DH>
__global__ void __launch_bounds__(128, 3)
attn_halo_dkv_kernel(vtb_attn_params p, Geom g, int groups, int nchunks) {
  extern __shared__ __align__(16) uint8_t hk_smem[];
  const int h = blockIdx.x % g.heads;
  const int chunk = blockIdx.x / g.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int W = g.window, HL = g.halo, KW = W + 2 * HL;
  const int nq = g.nq;                       // W*W tokens per block
  const int nby = g.Hs / W;
  const int reach = (HL + W - 1) / W;        // neighbour blocks whose halo can reach this block
  const float sl2 = p.scale * 1.4426950408889634f;
  constexpr float L2E = 1.4426950408889634f;

  // smem: K, V (own tokens), Q, dO (current neighbour) [64][DH]; lse2, delta [64]; bias tile f32 [nq][nkv] (x log2 e)
  bf16* sK = reinterpret_cast<bf16*>(hk_smem);
  bf16* sV = sK + 64 * DH;
  bf16* sQ = sV + 64 * DH;
  bf16* sdO = sQ + 64 * DH;
  float* sLse = reinterpret_cast<float*>(sdO + 64 * DH);
  float* sDelta = sLse + 64;
  int* sTok = reinterpret_cast<int*>(sDelta + 64);          // [64] own tokens
  int* sNTok = sTok + 64;                                   // [64] neighbour tokens
  float* sBias = reinterpret_cast<float*>(sNTok + 64);      // [nq][nkv], built once per CTA (fixed head)

  constexpr float L2E_ = 1.4426950408889634f;
  for (int e = threadIdx.x; e < nq * g.nkv; e += blockDim.x)
    sBias[e] = __ldg(p.rel_bias + (long)__ldg(p.pos + e) * g.heads + h) * L2E_;

  // persistent over the blocks of this head: the bias tile is built once per CTA
  for (int grp = chunk; grp < groups; grp += nchunks) {
  const int b = grp / g.nw, wi = grp - b * g.nw;
  const int by = wi / g.nwx, bx = wi - by * g.nwx;
  __syncthreads();  // previous block's tiles fully consumed (and the bias tile visible)
  for (int i = threadIdx.x; i < 64; i += blockDim.x) {
    int tok = -1;
    if (i < nq) {
      const int cy = i / W, cx = i - cy * W;
      tok = (b * g.Hs + by * W + cy) * g.Ws + bx * W + cx;
    }
    sTok[i] = tok;
  }
  __syncthreads();
  {
    constexpr int CH = DH / 8;
    for (int c = threadIdx.x; c < 64 * CH; c += blockDim.x) {
      const int r = c / CH, cc = c - r * CH;
      const int tok = sTok[r];
      const long o = (long)(tok < 0 ? 0 : tok);
      cp_async16(smem_u32(sK + soff<DH>(r, cc)), reinterpret_cast<const bf16*>(p.k) + o * p.ldk + h * DH + cc * 8, tok >= 0);
      cp_async16(smem_u32(sV + soff<DH>(r, cc)), reinterpret_cast<const bf16*>(p.v) + o * p.ldv + h * DH + cc * 8, tok >= 0);
    }
    cp_async_commit();
  }

  const int c0 = warp * 16;  // this warp's 16 centre keys
  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int n = 0; n < DH / 8; ++n) {
    dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
  }
  const bf16* Og = reinterpret_cast<const bf16*>(p.o);
  const bf16* dOg = reinterpret_cast<const bf16*>(p.dout);

  for (int dy = -reach; dy <= reach; ++dy) {
    const int ny = by + dy;
    if (ny < 0 || ny >= nby) continue;         // CTA-uniform
    for (int dx = -reach; dx <= reach; ++dx) {
      const int nx = bx + dx;
      if (nx < 0 || nx >= g.nwx) continue;
      const int ngrp = b * g.nw + ny * g.nwx + nx;
      __syncthreads();  // previous neighbour's tiles fully consumed
      // neighbour query tokens, lse2 and delta (= rowsum(dO o O), left in the workspace by the dQ kernel)
      if (threadIdx.x < 64) {
        const int row = threadIdx.x;
        int tok = -1;
        if (row < nq) {
          const int ty = row / W, tx = row - ty * W;
          tok = (b * g.Hs + ny * W + ty) * g.Ws + nx * W + tx;
        }
        const long li = ((long)ngrp * g.heads + h) * nq + row;
        sNTok[row] = tok;
        sDelta[row] = tok >= 0 ? __ldg(p.delta + li) : 0.f;
        sLse[row] = tok >= 0 ? __ldg(p.lse + li) * L2E : INFINITY;
      }
      __syncthreads();
      {
        constexpr int CH = DH / 8;
        for (int c = threadIdx.x; c < 64 * CH; c += blockDim.x) {
          const int r = c / CH, cc = c - r * CH;
          const int tok = sNTok[r];
          const long o = (long)(tok < 0 ? 0 : tok);
          cp_async16(smem_u32(sQ + soff<DH>(r, cc)), reinterpret_cast<const bf16*>(p.q) + o * p.ldq + h * DH + cc * 8, tok >= 0);
          cp_async16(smem_u32(sdO + soff<DH>(r, cc)), dOg + o * p.lddo + h * DH + cc * 8, tok >= 0);
        }
        cp_async_commit();
      }
      cp_async_wait<0>();
      __syncthreads();

      // S^T = K Q^T, dP^T = V dO^T  (16 centre keys x 64 neighbour queries per warp)
      float sc[8][4], dp[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
        dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        uint32_t kf[4], vf[4];
        ldsm_x4(kf, smem_u32(sK + soff<DH>(c0 + (lane & 15), kk * 2 + (lane >> 4))));
        ldsm_x4(vf, smem_u32(sV + soff<DH>(c0 + (lane & 15), kk * 2 + (lane >> 4))));
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
          uint32_t qb[4], ob[4];
          res_ld_b<DH>(qb, sQ, n2 * 16, kk, lane);
          res_ld_b<DH>(ob, sdO, n2 * 16, kk, lane);
          uint32_t q0[2] = {qb[0], qb[1]}, q1[2] = {qb[2], qb[3]};
          uint32_t o0[2] = {ob[0], ob[1]}, o1[2] = {ob[2], ob[3]};
          mma_bf16_16816(sc[2 * n2], kf, q0);
          mma_bf16_16816(sc[2 * n2 + 1], kf, q1);
          mma_bf16_16816(dp[2 * n2], vf, o0);
          mma_bf16_16816(dp[2 * n2 + 1], vf, o1);
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int c = c0 + gq + r * 8;          // centre key index
        const int cy = c / W, cx = c - cy * W;
        const int ky = cy - W * dy + HL, kx = cx - W * dx + HL;
        const bool kvalid = (c < nq) && ky >= 0 && ky < KW && kx >= 0 && kx < KW;
        const int slot = kvalid ? ky * KW + kx : 0;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int t = n * 8 + 2 * tq + e;   // neighbour query index
            float pv = 0.f, ds = 0.f;
            if (kvalid && t < nq) {
              pv = exp2f(fmaf(sc[n][2 * r + e], sl2, sBias[t * g.nkv + slot] - sLse[t]));
              ds = pv * (dp[n][2 * r + e] - sDelta[t]);
            }
            sc[n][2 * r + e] = pv;
            dp[n][2 * r + e] = ds;
          }
        }
      }
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        uint32_t pa[4] = {pack_bf16(sc[2 * k2][0], sc[2 * k2][1]), pack_bf16(sc[2 * k2][2], sc[2 * k2][3]),
                          pack_bf16(sc[2 * k2 + 1][0], sc[2 * k2 + 1][1]),
                          pack_bf16(sc[2 * k2 + 1][2], sc[2 * k2 + 1][3])};
        uint32_t da[4] = {pack_bf16(dp[2 * k2][0], dp[2 * k2][1]), pack_bf16(dp[2 * k2][2], dp[2 * k2][3]),
                          pack_bf16(dp[2 * k2 + 1][0], dp[2 * k2 + 1][1]),
                          pack_bf16(dp[2 * k2 + 1][2], dp[2 * k2 + 1][3])};
#pragma unroll
        for (int d2 = 0; d2 < DH / 16; ++d2) {
          uint32_t ob[4], qb[4];
          res_ld_bt<DH>(ob, sdO, k2 * 16, d2, lane);
          res_ld_bt<DH>(qb, sQ, k2 * 16, d2, lane);
          uint32_t o0[2] = {ob[0], ob[1]}, o1[2] = {ob[2], ob[3]};
          uint32_t q0[2] = {qb[0], qb[1]}, q1[2] = {qb[2], qb[3]};
          mma_bf16_16816(dv[2 * d2], pa, o0);
          mma_bf16_16816(dv[2 * d2 + 1], pa, o1);
          mma_bf16_16816(dk[2 * d2], da, q0);
          mma_bf16_16816(dk[2 * d2 + 1], da, q1);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int tok = sTok[c0 + gq + r * 8];
    if (tok < 0) continue;
    bf16* dKp = reinterpret_cast<bf16*>(p.dk) + (long)tok * p.lddk + h * DH;
    bf16* dVp = reinterpret_cast<bf16*>(p.dv) + (long)tok * p.lddv + h * DH;
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
      *reinterpret_cast<uint32_t*>(dKp + n * 8 + 2 * tq) = pack_bf16(dk[n][2 * r] * p.scale, dk[n][2 * r + 1] * p.scale);
      *reinterpret_cast<uint32_t*>(dVp + n * 8 + 2 * tq) = pack_bf16(dv[n][2 * r], dv[n][2 * r + 1]);
    }
  }
  }  // blocks
}

}  // namespace

// =====================================================================================================
// Validation mode (vtb_attention_fwd_f32): exact-softmax attention in fp32 on the CUDA cores, every geometry of
// vtb_attn_params (global / shifted window with bias + mask / halo with zero-padded slots).  One warp per (group, head,
// query): online softmax over the keys, lanes hold the head dimension.  q / k / v / o are FLOAT buffers here (leading
// dimensions in elements).  Not a fallback of the bf16 kernels: it exists so that the forward of a whole model can be
// checked against the reference at the north star's rtol 1e-3, which bf16 operands cannot meet.
// =====================================================================================================
namespace {
__global__ void __launch_bounds__(256)
attn_f32_fwd_kernel(vtb_attn_params p, Geom g, long n_rows) {
  const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= n_rows) return;
  const int i = (int)(wid % g.nq);
  const long gh = wid / g.nq;
  const int h = (int)(gh % g.heads);
  const int grp = (int)(gh / g.heads);
  const long tq = q_token(g, grp, i);
  const float* Q = reinterpret_cast<const float*>(p.q);
  const float* K = reinterpret_cast<const float*>(p.k);
  const float* V = reinterpret_cast<const float*>(p.v);
  float* O = reinterpret_cast<float*>(p.o);
  const bool two = g.dh > 32;
  const float q0 = Q[tq * p.ldq + h * g.dh + lane];
  const float q1 = two ? Q[tq * p.ldq + h * g.dh + 32 + lane] : 0.f;
  const uint8_t* mrow = p.mask ? p.mask + ((long)(grp % p.n_mask) * g.nq + i) * (p.mask_ld ? p.mask_ld : g.nkv) : nullptr;
  float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f;
  for (int j = 0; j < g.nkv; ++j) {
    if (mrow && mrow[j]) continue;  // masked_fill(-inf): the key drops out of the softmax
    const long tk = kv_token(g, grp, j);
    float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;  // zero-padded halo slot: score = bias, value = 0
    if (tk >= 0) {
      k0 = K[tk * p.ldk + h * g.dh + lane];
      v0 = V[tk * p.ldv + h * g.dh + lane];
      if (two) {
        k1 = K[tk * p.ldk + h * g.dh + 32 + lane];
        v1 = V[tk * p.ldv + h * g.dh + 32 + lane];
      }
    }
    float sdot = warp_sum(fmaf(q0, k0, q1 * k1)) * p.scale;
    if (p.rel_bias) sdot += p.rel_bias[(long)p.pos[(long)i * g.nkv + j] * g.heads + h];
    const float mn = fmaxf(m, sdot);
    const float corr = expf(m - mn), pj = expf(sdot - mn);  // m = -inf on the first key: corr = 0
    l = l * corr + pj;
    a0 = a0 * corr + pj * v0;
    a1 = a1 * corr + pj * v1;
    m = mn;
  }
  const float inv = l > 0.f ? 1.f / l : 0.f;
  O[tq * p.ldo + h * g.dh + lane] = a0 * inv;
  if (two) O[tq * p.ldo + h * g.dh + 32 + lane] = a1 * inv;
  if (p.lse && lane == 0) p.lse[((long)grp * g.heads + h) * g.nq + i] = m + logf(l);
}
}  // namespace

extern "C" int vtb_attention_fwd_f32(const vtb_attn_params* p, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Geom g;
  long groups;
  VTB_CHECK(p != nullptr && p->q && p->k && p->v && p->o, -1, "vtb_attention_fwd_f32: null pointer");
  int rc = make_geom(p, &g, &groups, "vtb_attention_fwd_f32");  // the bf16 layout rules (rows of 8 elements) hold for float rows too
  if (rc) return rc;
  const long n_rows = groups * p->heads * p->nq;
  const long blocks = (n_rows * 32 + 255) / 256;
  VTB_CHECK(blocks < (1L << 31), -1, "vtb_attention_fwd_f32: grid too large");
  attn_f32_fwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*p, g, n_rows);
  VTB_LAUNCH_CHECK();
  return 0;
}

bool vtb_attn_halo_dkv_ok(const vtb_attn_params* p) {
  return p->mode == VTB_ATTN_HALO && !p->dkv_f32 && p->nq <= 64 && p->rel_bias != nullptr && p->n_pos <= MAX_POS &&
         p->delta != nullptr;
}

int vtb_attn_halo_dkv(const vtb_attn_params* p, const Geom& g, long groups, cudaStream_t stream) {
  const size_t smem = (size_t)4 * 64 * p->dh * 2 + 2 * 64 * 4 + 2 * 64 * 4 + (size_t)p->nq * p->nkv * 4 + 16;
  VTB_CHECK(smem <= 227 * 1024, -1, "vtb_attention_bwd: halo dK/dV tile needs %zu B of shared memory", smem);
  static bool set64 = false, set32 = false;
  if (p->dh == 64 && !set64) { VTB_CUDA(cudaFuncSetAttribute(attn_halo_dkv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); set64 = true; }
  if (p->dh == 32 && !set32) { VTB_CUDA(cudaFuncSetAttribute(attn_halo_dkv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); set32 = true; }
  // one resident wave, persistent over the blocks of a head (the bias tile is built once per CTA)
  int per_sm = 1;
  if (p->dh == 64) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn_halo_dkv_kernel<64>, 128, smem);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attn_halo_dkv_kernel<32>, 128, smem);
  if (per_sm < 1) per_sm = 1;
  long nchunks = (long)vtb_num_sms() * per_sm / p->heads;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > groups) nchunks = groups;
  const long blocks = nchunks * p->heads;
  VTB_CHECK(blocks < (1L << 31) && groups < (1L << 31), -1, "vtb_attention_bwd: grid too large");
  if (p->dh == 64) attn_halo_dkv_kernel<64><<<(unsigned)blocks, 128, smem, stream>>>(*p, g, (int)groups, (int)nchunks);
  else attn_halo_dkv_kernel<32><<<(unsigned)blocks, 128, smem, stream>>>(*p, g, (int)groups, (int)nchunks);
  VTB_LAUNCH_CHECK();
  return 0;
}
