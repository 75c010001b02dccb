// tcgen05 / TMEM attention forward for global attention with dh = 64 and <= 256 keys
// (ViT-B / DeiT-S: N = 197 or 37; PVT: 3136/784/196/50 queries x 49/50 keys).
//
// One CTA per (image, head, 128-query tile), 2 CTAs per SM (80 KB smem, 256 TMEM columns each):
//   warp 0   : TMEM alloc, barrier init, TMA loads of Q / K / V straight out of the fused projection buffer
//              (3-D tensor maps [image][token][column]: tokens past the end of an image are zero-filled)
//   warp 1   : UMMA issuer.  S[128 x Ns] = Q K^T (4 x tcgen05.mma, operands in smem, K-major),
//              then O[128 x 64] = P V with P read FROM TMEM as the A operand and V as an MN-major smem
//              operand (the very tile TMA delivered: no transpose)
//   warps 2-5: softmax, one query row per thread, directly on TMEM (no shuffles, no smem): pass 1 row max,
//              pass 2 exp2 / row sum / bf16 pack, storing P back into TMEM over the S columns already consumed;
//              then the epilogue O / l -> bf16 -> global and lse.
// Scores and probabilities never touch shared or global memory.
#include "common.cuh"
#include "../../include/vtb200.h"
#include <math.h>

namespace {

constexpr int TQ = 128;        // query rows per CTA
constexpr int DH = 64;
constexpr int MAXK = 256;      // key slots per CTA (TMA boxes of 128 rows x 2)
constexpr int TC_THREADS = 192;
constexpr int O_COL = 128;     // O accumulator columns [128,192): inside S's footprint, written only after S is consumed
constexpr int TMEM_COLS_ATT = 256;
constexpr int SMEM_ATT = TQ * 128 + 2 * MAXK * 128 + 1024 + 128;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_att = nullptr;

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
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
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(TC_THREADS, 2)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, bf16* __restrict__ O, int ldo,
                   float* __restrict__ lse, int heads, int nq, int nkv, int q_tiles, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                       // [128 rows][128 B]  K-major A
  uint8_t* sK = sQ + TQ * 128;              // [256 rows][128 B]  K-major B (rows = keys)
  uint8_t* sV = sK + MAXK * 128;            // [256 rows][128 B]  MN-major B (rows = keys = contraction)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + MAXK * 128);
  uint64_t* bar_qk = bars;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int h = bh % heads;
  const int b = bh / heads;
  const int i0 = qt * TQ;
  const int ns = (nkv + 15) & ~15;          // UMMA N of the score tile / contraction length of P V

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv);
      mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 4); mbar_init(bar_o, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS_ATT);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const int nbox = (nkv > 128) ? 2 : 1;  // key rows arrive in boxes of 128 tokens
      mbar_expect_tx(bar_qk, (uint32_t)(TQ * 128 + nbox * 128 * 128));
      tma_load_3d(sQ, &tq, bar_qk, h * DH, i0, b);
      for (int i = 0; i < nbox; ++i) tma_load_3d(sK + i * 128 * 128, &tk, bar_qk, h * DH, i * 128, b);
      mbar_expect_tx(bar_v, (uint32_t)(nbox * 128 * 128));
      for (int i = 0; i < nbox; ++i) tma_load_3d(sV + i * 128 * 128, &tv, bar_v, h * DH, i * 128, b);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // S = Q K^T : M = 128, N = ns, K = 64 (4 x K16), both operands K-major
      const uint32_t idesc_s = umma_idesc_bf16(TQ, ns, 0, 0);
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_bf16(tmem_base, umma_desc_sw128(qa + k * 32, 0, 1024), umma_desc_sw128(ka + k * 32, 0, 1024),
                  idesc_s, k > 0 ? 1u : 0u);
      umma_commit(bar_s);
      // O = P V : A = P from TMEM (bf16 pairs, 8 columns per K16 step), B = V MN-major, N = 64, K = ns
      const uint32_t idesc_o = umma_idesc_bf16(TQ, DH, 0, 1);
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t va = smem_u32(sV);
      for (int k = 0; k < ns / 16; ++k)
        umma_bf16_ts(tmem_base + O_COL, tmem_base + k * 8, umma_desc_sw128(va + k * (16 * 128), 0, 1024), idesc_o,
                     k > 0 ? 1u : 0u);
      umma_commit(bar_o);
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 2..5)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;      // query row in the tile == TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = scale * 1.4426950408889634f;  // exp(x) = exp2(x * log2 e)
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // pass 1: row max of the raw scores (scale > 0, so max commutes with the scaling)
    float mx = -INFINITY;
    for (int c = 0; c < ns; c += 32) {
      if (ns - c >= 32) {
        uint32_t a[32];
        tmem_ld_32x32(t_row + c, a);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c + j < nkv) mx = fmaxf(mx, __uint_as_float(a[j]));
      } else {
        uint32_t a[16];
        tmem_ld_32x16(t_row + c, a);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c + j < nkv) mx = fmaxf(mx, __uint_as_float(a[j]));
      }
    }
    const float mb = mx * sl2;
    // pass 2: p = exp2(s * sl2 - mb); P (bf16 pairs) overwrites the S columns this thread has already consumed
    float sum = 0.f;
    for (int c = 0; c < ns; c += 32) {
      if (ns - c >= 32) {
        uint32_t a[32];
        tmem_ld_32x32(t_row + c, a);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = (c + j < nkv) ? ex2f(fmaf(__uint_as_float(a[j]), sl2, -mb)) : 0.f;
          const float p1 = (c + j + 1 < nkv) ? ex2f(fmaf(__uint_as_float(a[j + 1]), sl2, -mb)) : 0.f;
          sum += p0 + p1;
          pk[j >> 1] = pack_bf16(p0, p1);
        }
        tmem_st_32x16(t_row + (c >> 1), pk);
      } else {
        uint32_t a[16];
        tmem_ld_32x16(t_row + c, a);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float p0 = (c + j < nkv) ? ex2f(fmaf(__uint_as_float(a[j]), sl2, -mb)) : 0.f;
          const float p1 = (c + j + 1 < nkv) ? ex2f(fmaf(__uint_as_float(a[j + 1]), sl2, -mb)) : 0.f;
          sum += p0 + p1;
          pk[j >> 1] = pack_bf16(p0, p1);
        }
#pragma unroll
        for (int j = 8; j < 16; ++j) pk[j] = 0u;
        tmem_st_32x16(t_row + (c >> 1), pk);  // 8 live columns + 8 zero columns (never read by the MMA)
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);
    // epilogue: O / l -> bf16 -> global; lse = max * scale + ln(sum)
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const int i = i0 + row;
    const float inv = 1.f / sum;
#pragma unroll
    for (int hc = 0; hc < 2; ++hc) {
      uint32_t a[32];
      tmem_ld_32x32(t_row + O_COL + hc * 32, a);
      tmem_ld_wait();
      if (i < nq) {
        bf16* dst = O + ((long)b * nq + i) * ldo + h * DH + hc * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          *reinterpret_cast<uint4*>(dst + j) = make_uint4(
              pack_bf16(__uint_as_float(a[j]) * inv, __uint_as_float(a[j + 1]) * inv),
              pack_bf16(__uint_as_float(a[j + 2]) * inv, __uint_as_float(a[j + 3]) * inv),
              pack_bf16(__uint_as_float(a[j + 4]) * inv, __uint_as_float(a[j + 5]) * inv),
              pack_bf16(__uint_as_float(a[j + 6]) * inv, __uint_as_float(a[j + 7]) * inv));
      }
    }
    if (i < nq && lse) lse[((long)b * heads + h) * nq + i] = mx * scale + __logf(sum);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS_ATT);
  }
}


#ifdef VTB_ATTN_TRACE
// timeline of block 0 (debug builds): per warp a private list of {event id, clock} pairs (plain stores, no atomics: the
// probe must not stall the warp), fetched with vtb_debug_attn_trace
__device__ unsigned int g_b2_trace[16 * 2 * 1024];
#define B2_TRACE_DECL unsigned int trace_i__ = 0
#define B2_TRACE(ev)                                                                              \
  do {                                                                                            \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trace_i__ < 1024) {                         \
      unsigned int* p__ = g_b2_trace + ((threadIdx.x >> 5) * 1024 + trace_i__) * 2;              \
      p__[0] = (ev); p__[1] = (unsigned int)clock64();                                            \
      ++trace_i__;                                                                                \
    }                                                                                             \
  } while (0)
#else
#define B2_TRACE_DECL do { } while (0)
#define B2_TRACE(ev) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------
// Forward, version 2: PERSISTENT, one CTA per SM, two query tiles in flight, loads prefetched one work item ahead.
//
// A work item is a pair of 128-query tiles that share one TMA stage (96 KB: Q0 Q1 | K[256] | V[256]):
//   * shared mode  (nq > 128 or nkv > 128: ViT-B / DeiT-S at 224^2, PVT): both tiles belong to one (image, head), the
//     keys / values are loaded ONCE for both (version 1 loaded them once per tile);
//   * paired mode  (nq <= 128 and nkv <= 128: 96^2 crops with 37 tokens, the PVT cls stage): the tiles are two
//     different (image, head) problems, tile t uses key rows [128 t, 128 t + nkv) of the stage.
// warps 0-3: softmax + epilogue of tile 0, warps 4-7: of tile 1 (one query row per thread = one TMEM lane);
// warp 8 lane 0: TMA producer (2-stage ring, full / empty mbarriers);   warp 9 lane 0: UMMA issuer.
// TMEM: tile t owns columns [256 t, 256 t + 256): S at +0 (P written back over it as bf16 pairs), O at +128.
// While set 0 runs its softmax the tensor pipe computes S of tile 1; while set 1 finishes, P V of tile 0 and the
// next item's S of tile 0 are issued, and the next item's Q / K / V have been in flight since the stage was released.
// ---------------------------------------------------------------------------------------------------------------
constexpr int F2_THREADS = 320;
constexpr int F2_STAGE = 6 * 16384;
constexpr int F2_OSTAGE = 8 * 4096;   // per softmax warp: 32 output rows x 128 B, transposed before they are written out
constexpr int SMEM_F2 = 2 * F2_STAGE + F2_OSTAGE + 1024 + 256;

__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// work item -> (image * heads + head, first query token) of tile t, and the number of tiles (1 or 2) of the item
__device__ __forceinline__ int f2_ntile(int item, int ppb, int paired, int n_bh, int nq) {
  if (paired) return (2 * item + 1 < n_bh) ? 2 : 1;
  return ((item % ppb) * 256 + 128 < nq) ? 2 : 1;
}
__device__ __forceinline__ int f2_bh(int item, int t, int ppb, int paired) { return paired ? 2 * item + t : item / ppb; }
__device__ __forceinline__ int f2_q0(int item, int t, int ppb, int paired) { return paired ? 0 : (item % ppb) * 256 + t * 128; }

// row maximum over W (16 or 32) score columns already in registers; MASK: columns >= lim do not count
template <int W, bool MASK>
__device__ __forceinline__ float f2_rowmax(const uint32_t (&a)[W], int lim, float mx) {
  float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < W; j += 4) {
    if (!MASK || j < lim) m0 = fmaxf(m0, __uint_as_float(a[j]));
    if (!MASK || j + 1 < lim) m1 = fmaxf(m1, __uint_as_float(a[j + 1]));
    if (!MASK || j + 2 < lim) m2 = fmaxf(m2, __uint_as_float(a[j + 2]));
    if (!MASK || j + 3 < lim) m3 = fmaxf(m3, __uint_as_float(a[j + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// p = exp2(s * sl2 - mb) for W columns in registers, bf16 pairs stored at `paddr`; returns the row sum of the chunk
template <int W, bool MASK>
__device__ __forceinline__ float f2_probs(const uint32_t (&a)[W], uint32_t paddr, int lim, float sl2, float mb) {
  uint32_t pk[16];
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < W; j += 2) {
    float p0 = ex2f(fmaf(__uint_as_float(a[j]), sl2, -mb));
    float p1 = ex2f(fmaf(__uint_as_float(a[j + 1]), sl2, -mb));
    if (MASK) {
      p0 = (j < lim) ? p0 : 0.f;
      p1 = (j + 1 < lim) ? p1 : 0.f;
    }
    s0 += p0; s1 += p1;
    pk[j >> 1] = pack_bf16(p0, p1);
  }
  if (W == 32) tmem_st_32x16(paddr, pk);
  else tmem_st_32x8(paddr, pk);
  return s0 + s1;
}

__global__ void __launch_bounds__(F2_THREADS, 1)
attn_tc_fwd2_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                    const __grid_constant__ CUtensorMap tv, bf16* __restrict__ O, int ldo, float* __restrict__ lse,
                    int heads, int nq, int nkv, int n_bh, int n_items, int ppb, int paired, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sO = smem + 2 * F2_STAGE;   // [8 warps][32 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * F2_STAGE + F2_OSTAGE);
  uint64_t* bar_full = bars;        // [2] TMA -> issuer: the stage's Q / K / V landed
  uint64_t* bar_empty = bars + 2;   // [2] issuer -> TMA: every MMA reading the stage retired
  uint64_t* bar_s = bars + 4;       // [2] per tile: S complete
  uint64_t* bar_p = bars + 6;       // [2] per tile: P written (4 warps)
  uint64_t* bar_o = bars + 8;       // [2] per tile: O complete
  uint64_t* bar_free = bars + 10;   // [2] per tile: O read out, the tile's TMEM columns may be overwritten (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  B2_TRACE_DECL;
  const int ns = (nkv + 15) & ~15;  // UMMA N of the score tile / contraction length of P V

  // warp roles: 0-3 softmax + epilogue of tile 0, 4-7 of tile 1, 8 TMA producer, 9 UMMA issuer (the arbiter prefers the
  // highest warp id of a scheduler partition: the issuer must not starve behind the softmax warps)
  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_full + i, 1); mbar_init(bar_empty + i, 1); mbar_init(bar_s + i, 1); mbar_init(bar_p + i, 4);
        mbar_init(bar_o + i, 1); mbar_init(bar_free + i, 4);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      const int nbox = (nkv > 128) ? 2 : 1;
      int k = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
        const int s = k & 1;
        mbar_wait(bar_empty + s, (uint32_t)(((k >> 1) & 1) ^ 1));
        const int ntile = f2_ntile(item, ppb, paired, n_bh, nq);
        uint8_t* sQ = smem + s * F2_STAGE;
        uint8_t* sK = sQ + 2 * 16384;
        uint8_t* sV = sK + 2 * 16384;
        const int kv_boxes = paired ? ntile : nbox;
        mbar_expect_tx(bar_full + s, (uint32_t)((ntile + 2 * kv_boxes) * 16384));
        for (int t = 0; t < ntile; ++t) {
          const int bh = f2_bh(item, t, ppb, paired);
          tma_load_3d(sQ + t * 16384, &tq, bar_full + s, (bh % heads) * DH, f2_q0(item, t, ppb, paired), bh / heads);
        }
        for (int i = 0; i < kv_boxes; ++i) {
          const int bh = f2_bh(item, i, ppb, paired);
          const int tok = paired ? 0 : i * 128;
          tma_load_3d(sK + i * 16384, &tk, bar_full + s, (bh % heads) * DH, tok, bh / heads);
          tma_load_3d(sV + i * 16384, &tv, bar_full + s, (bh % heads) * DH, tok, bh / heads);
        }
      }
    }
  } else if (warp == 9) {
    // All 32 lanes run the loop (everything is warp-uniform: addresses and descriptors stay in uniform registers); one
    // elected lane issues the tcgen05 instructions, straight-line (the issuing thread, not the tensor pipe, bounds a run of
    // small MMAs: tools/probes/mma_probe.cu).
    {
      const uint32_t idesc_s = umma_idesc_bf16(TQ, ns, 0, 0);   // S = Q K^T: both operands K-major
      const uint32_t idesc_o = umma_idesc_bf16(TQ, DH, 0, 1);   // O = P V: A = P from TMEM, B = V MN-major
      const int nks = ns >> 4;
      // The two tiles are independent streams  S_t(i) -> softmax -> PV_t(i) -> epilogue -> S_t(i+1) ...; they are served in
      // the fixed order  PV_0(i), S_0(i+1), PV_1(i), S_1(i+1): tile 1 then trails tile 0 by at least (P V + epilogue + S), so
      // the exp-heavy second pass of one warp set overlaps the MUFU-free phases (waits, row max, epilogue) of the other.
      auto issue_s = [&](int t, int s) {
        const uint32_t qa = smem_u32(smem + s * F2_STAGE), ka = qa + 2 * 16384;
        const uint32_t koff = paired ? (uint32_t)t * 16384u : 0u;
        const uint64_t dq = umma_desc_sw128(qa + t * 16384, 0, 1024), dk = umma_desc_sw128(ka + koff, 0, 1024);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk)
            umma_bf16(tmem_base + t * 256, dq + 2 * kk, dk + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(bar_s + t);
        }
        __syncwarp();
        B2_TRACE(30 + t);
      };
      auto issue_pv = [&](int t, int s, bool release) {
        const uint32_t va = smem_u32(smem + s * F2_STAGE) + 4 * 16384;
        const uint32_t koff = paired ? (uint32_t)t * 16384u : 0u;
        const uint64_t dv = umma_desc_sw128(va + koff, 0, 1024);
        const uint32_t d = tmem_base + t * 256 + O_COL, a = tmem_base + t * 256;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 16; ++kk)
            if (kk < nks) umma_bf16_ts(d, a + kk * 8, dv + (uint64_t)(kk * 128), idesc_o, kk > 0 ? 1u : 0u);
          umma_commit(bar_o + t);
          if (release) umma_commit(bar_empty + s);
        }
        __syncwarp();
        B2_TRACE(32 + t);
      };
      int it0 = 0, it1 = 0, k = 0;   // per-tile iteration counters (tile 1 is absent from some items), local item counter
      int item = blockIdx.x;
      if (item < n_items) {
        mbar_wait(bar_full, 0);
        tc_fence_after();
        issue_s(0, 0);
        if (f2_ntile(item, ppb, paired, n_bh, nq) > 1) issue_s(1, 0);
      }
      for (; item < n_items; item += gridDim.x, ++k) {
        const int s = k & 1;
        const int ntile = f2_ntile(item, ppb, paired, n_bh, nq);
        const int nxt = item + gridDim.x;
        const bool has_next = nxt < n_items;
        const int ntile_next = has_next ? f2_ntile(nxt, ppb, paired, n_bh, nq) : 0;
        B2_TRACE(34);
        mbar_wait(bar_p, (uint32_t)(it0 & 1));
        tc_fence_after();
        B2_TRACE(35);
        issue_pv(0, s, ntile == 1);
        ++it0;
        if (has_next) {
          mbar_wait(bar_full + (s ^ 1), (uint32_t)(((k + 1) >> 1) & 1));
          B2_TRACE(36);
          mbar_wait(bar_free, (uint32_t)((it0 & 1) ^ 1));
          tc_fence_after();
          B2_TRACE(37);
          issue_s(0, s ^ 1);
        }
        if (ntile > 1) {
          mbar_wait(bar_p + 1, (uint32_t)(it1 & 1));
          tc_fence_after();
          B2_TRACE(38);
          issue_pv(1, s, true);
          ++it1;
        }
        if (ntile_next > 1) {
          mbar_wait(bar_free + 1, (uint32_t)((it1 & 1) ^ 1));
          tc_fence_after();
          issue_s(1, s ^ 1);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue (warps 0..7)
    const int t = warp >> 2;                   // the tile this warp set serves
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;       // query row in the tile == TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)t * 256u;
    const float sl2 = scale * 1.4426950408889634f;  // exp(x) = exp2(x * log2 e)
    const int nfull = nkv & ~31;               // columns covered by unmasked 32-column chunks
    const int tail = ns - nfull;               // 0, 16 or 32 masked columns
    const int n32 = nfull >> 5;                // unmasked 32-column chunks
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      if (t >= f2_ntile(item, ppb, paired, n_bh, nq)) continue;
      const int bh = f2_bh(item, t, ppb, paired), q0 = f2_q0(item, t, ppb, paired);
      const int rows_valid = min(128, nq - q0);
      const bool active = quarter * 32 < rows_valid;  // warps whose 32 rows are all padding only keep the barriers moving
      B2_TRACE(40);
      mbar_wait(bar_s + t, (uint32_t)(it & 1));
      tc_fence_after();
      B2_TRACE(41);
      float mx = -INFINITY, sum = 0.f;
      if (active) {
        // Both passes keep one TMEM load in flight: chunk c + 1 is requested before chunk c is processed.
        uint32_t a[32], b[32];
        // pass 1: row max of the raw scores (scale > 0, so max commutes with the scaling)
        if (n32 > 0) tmem_ld_32x32(t_row, a);
        for (int c = 0; c < n32; c += 2) {
          tmem_ld_wait();
          if (c + 1 < n32) tmem_ld_32x32(t_row + (c + 1) * 32, b);
          mx = f2_rowmax<32, false>(a, 32, mx);
          if (c + 1 < n32) {
            tmem_ld_wait();
            if (c + 2 < n32) tmem_ld_32x32(t_row + (c + 2) * 32, a);
            mx = f2_rowmax<32, false>(b, 32, mx);
          }
        }
        if (tail == 32) {
          tmem_ld_32x32(t_row + nfull, a);
          tmem_ld_wait();
          mx = f2_rowmax<32, true>(a, nkv - nfull, mx);
        } else if (tail == 16) {
          uint32_t h16[16];
          tmem_ld_32x16(t_row + nfull, h16);
          tmem_ld_wait();
          mx = f2_rowmax<16, true>(h16, nkv - nfull, mx);
        }
        const float mb = mx * sl2;
        B2_TRACE(42);
        // pass 2: p = exp2(s * sl2 - mb); P (bf16 pairs) overwrites the S columns this thread has already consumed
        // (chunk c's P lands in columns [16 c, 16 c + 16), below every column still to be read)
        if (n32 > 0) tmem_ld_32x32(t_row, a);
        for (int c = 0; c < n32; c += 2) {
          tmem_ld_wait();
          if (c + 1 < n32) tmem_ld_32x32(t_row + (c + 1) * 32, b);
          sum += f2_probs<32, false>(a, t_row + c * 16, 32, sl2, mb);
          if (c + 1 < n32) {
            tmem_ld_wait();
            if (c + 2 < n32) tmem_ld_32x32(t_row + (c + 2) * 32, a);
            sum += f2_probs<32, false>(b, t_row + (c + 1) * 16, 32, sl2, mb);
          }
        }
        if (tail == 32) {
          tmem_ld_32x32(t_row + nfull, a);
          tmem_ld_wait();
          sum += f2_probs<32, true>(a, t_row + (nfull >> 1), nkv - nfull, sl2, mb);
        } else if (tail == 16) {
          uint32_t h16[16];
          tmem_ld_32x16(t_row + nfull, h16);
          tmem_ld_wait();
          sum += f2_probs<16, true>(h16, t_row + (nfull >> 1), nkv - nfull, sl2, mb);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + t);
      B2_TRACE(43);
      if (active) {
        // epilogue: O / l -> bf16 -> global; lse = max * scale + ln(sum)
        mbar_wait(bar_o + t, (uint32_t)(it & 1));
        tc_fence_after();
        B2_TRACE(44);
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(t_row + O_COL, a0);
        tmem_ld_32x32(t_row + O_COL + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + t);   // O is in registers: the tile's TMEM columns are free
        // O / l -> bf16.  A thread owns one 128-byte output row, but rows are 2 * ldo bytes apart in global memory: written
        // straight from the registers every store instruction would touch 32 different lines with 16 bytes each.  The warp
        // transposes through its private 4 KB of shared memory instead (16-byte chunks XOR-swizzled by the row: no bank
        // conflicts either way), so that 8 lanes write one full line and an instruction covers 4 whole rows.
        {
          const float inv = 1.f / sum;
          const uint32_t so = smem_u32(sO) + (uint32_t)warp * 4096u;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sts_u4(so + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4),
                   pack_bf16(__uint_as_float(a0[8 * c]) * inv, __uint_as_float(a0[8 * c + 1]) * inv),
                   pack_bf16(__uint_as_float(a0[8 * c + 2]) * inv, __uint_as_float(a0[8 * c + 3]) * inv),
                   pack_bf16(__uint_as_float(a0[8 * c + 4]) * inv, __uint_as_float(a0[8 * c + 5]) * inv),
                   pack_bf16(__uint_as_float(a0[8 * c + 6]) * inv, __uint_as_float(a0[8 * c + 7]) * inv));
            sts_u4(so + (uint32_t)lane * 128u + (uint32_t)(((c + 4) ^ (lane & 7)) << 4),
                   pack_bf16(__uint_as_float(a1[8 * c]) * inv, __uint_as_float(a1[8 * c + 1]) * inv),
                   pack_bf16(__uint_as_float(a1[8 * c + 2]) * inv, __uint_as_float(a1[8 * c + 3]) * inv),
                   pack_bf16(__uint_as_float(a1[8 * c + 4]) * inv, __uint_as_float(a1[8 * c + 5]) * inv),
                   pack_bf16(__uint_as_float(a1[8 * c + 6]) * inv, __uint_as_float(a1[8 * c + 7]) * inv));
          }
          __syncwarp();
          const int b = bh / heads, h = bh - b * heads;
          const int sub = lane >> 3, ch = lane & 7;   // row inside a group of four, 16-byte chunk of the row
          bf16* base = O + ((long)b * nq + q0 + quarter * 32) * ldo + h * DH + ch * 8;
#pragma unroll
          for (int r4 = 0; r4 < 8; ++r4) {
            const int r = r4 * 4 + sub;
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(so + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4)));
            if (quarter * 32 + r < rows_valid) *reinterpret_cast<uint4*>(base + (long)r * ldo) = v;
          }
          __syncwarp();   // the staging rows are rewritten by the next item
          if (row < rows_valid && lse) lse[(long)bh * nq + q0 + row] = mx * scale + __logf(sum);
        }
        B2_TRACE(45);
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_free + t);
      }
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// [images][tokens][cols] bf16 view with a 64-column x `box_rows`-token box
int make_tmap3(CUtensorMap* map, const void* base, uint64_t cols, uint64_t tokens, uint64_t images, uint64_t ld,
               uint32_t box_rows) {
  cuuint64_t dims[3] = {cols, tokens, images};
  cuuint64_t strides[2] = {ld * 2, tokens * ld * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode_att(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vtb_set_error("cuTensorMapEncodeTiled(3d) failed (%d): cols=%llu tokens=%llu images=%llu ld=%llu", (int)r,
                  (unsigned long long)cols, (unsigned long long)tokens, (unsigned long long)images,
                  (unsigned long long)ld);
    return -3;
  }
  return 0;
}

bool g_attn_tc = true;
int g_attn_tc_fwd_version = 2;  // vtb_set_option("attn_tc_fwd_version", 1): the one-CTA-per-tile kernel (A/B timing)
int g_attn_tc_bwd_version = 2;

}  // namespace

void vtb_attn_tc_set(bool on) { g_attn_tc = on; }
void vtb_attn_tc_version_set(int fwd, int bwd) {
  if (fwd > 0) g_attn_tc_fwd_version = fwd;
  if (bwd > 0) g_attn_tc_bwd_version = bwd;
}

bool vtb_attn_tc_fwd_ok(const vtb_attn_params* p) {
  return g_attn_tc && p->mode == VTB_ATTN_GLOBAL && p->dh == DH && p->nkv <= MAXK && p->nkv >= 1 &&
         !p->rel_bias && !p->mask && p->ldo % 8 == 0 && (((uintptr_t)p->o) & 15) == 0;
}

static int ensure_encode() {
  if (!g_encode_att) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    VTB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    VTB_CHECK(fn != nullptr && q == cudaDriverEntryPointSuccess, -2, "cuTensorMapEncodeTiled not available");
    g_encode_att = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return 0;
}

int vtb_attn_tc_fwd(const vtb_attn_params* p, cudaStream_t stream) {
  if (int rc0 = ensure_encode()) return rc0;
  const uint64_t cols = (uint64_t)p->heads * DH;
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap3(&tq, p->q, cols, p->nq, p->batch, p->ldq, TQ))) return rc;
  if ((rc = make_tmap3(&tk, p->k, cols, p->nkv, p->batch, p->ldk, 128))) return rc;
  if ((rc = make_tmap3(&tv, p->v, cols, p->nkv, p->batch, p->ldv, 128))) return rc;
  static bool attr = false;
  if (!attr) {
    VTB_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ATT));
    VTB_CUDA(cudaFuncSetAttribute(attn_tc_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_F2));
    attr = true;
  }
  const int q_tiles = (p->nq + TQ - 1) / TQ;
  const long n_bh = (long)p->batch * p->heads;
  if (g_attn_tc_fwd_version == 1) {
    const long blocks = n_bh * q_tiles;
    VTB_CHECK(blocks < (1L << 31), -1, "vtb_attention_fwd: grid too large");
    attn_tc_fwd_kernel<<<(unsigned)blocks, TC_THREADS, SMEM_ATT, stream>>>(
        tq, tk, tv, reinterpret_cast<bf16*>(p->o), p->ldo, p->lse, p->heads, p->nq, p->nkv, q_tiles, p->scale);
    VTB_LAUNCH_CHECK();
    return 0;
  }
  // persistent kernel: a work item = two 128-query tiles sharing one stage (see attn_tc_fwd2_kernel)
  const int paired = (p->nq <= 128 && p->nkv <= 128) ? 1 : 0;
  const int ppb = (q_tiles + 1) / 2;
  const long n_items = paired ? (n_bh + 1) / 2 : n_bh * ppb;
  VTB_CHECK(n_items < (1L << 31) && n_bh < (1L << 30), -1, "vtb_attention_fwd: too many work items");
  const int grid = (int)(n_items < vtb_num_sms() ? n_items : vtb_num_sms());
  attn_tc_fwd2_kernel<<<grid, F2_THREADS, SMEM_F2, stream>>>(
      tq, tk, tv, reinterpret_cast<bf16*>(p->o), p->ldo, p->lse, p->heads, p->nq, p->nkv, (int)n_bh, (int)n_items, ppb,
      paired, p->scale);
  VTB_LAUNCH_CHECK();
  return 0;
}

// =====================================================================================================
// tcgen05 / TMEM attention BACKWARD (global attention, dh = 64, <= 256 queries and <= 256 keys per image).
// One CTA per (image, head); Q, K, V, dO resident in shared memory (TMA, 128 KB); every product on UMMA:
//   for each key tile kt (128 keys) and query half hq (128 queries):
//     S^T  = K_kt Q_hq^T           dP^T = V_kt dO_hq^T           (M = keys, N = queries, K = dh)
//     math warps (one key row per thread, straight from TMEM):
//        P^T = exp2(S^T sl2 - lse2[q]),   dS^T = scale * P^T (dP^T - delta[q])   -> bf16, 128B-swizzled smem
//     dV_kt += P^T dO_hq            dK_kt += dS^T Q_hq            (A = the smem tile, K-major; B MN-major)
//     dQ_hq += dS K_kt                                             (A = the SAME dS^T tile read MN-major)
// TMEM (512 columns): S^T 0-127 | dP^T 128-255 | dV 256-319 | dK 320-383 | dQ_0 384-447 | dQ_1 448-511.
// Probabilities are recomputed once (not twice as in the two-phase mma.sync kernel) and never leave the SM.
// =====================================================================================================
namespace {

constexpr int BWD_THREADS = 320;   // warp 0 TMA/alloc, warp 1 UMMA issuer, warps 2-9 math (two per TMEM lane quarter)
constexpr int BWD_MATH = 256;
constexpr int TILE_BYTES = 256 * 128;  // one resident operand: 256 token rows x 128 B
constexpr int PT_BYTES = 128 * 256;    // P^T / dS^T tile: 128 key rows x 128 queries (two 64-query blocks)
constexpr int SMEM_BWD = 4 * TILE_BYTES + 2 * PT_BYTES + 2 * 256 * 4 + 1024 + 256;

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                   const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                   const __grid_constant__ CUtensorMap to,
                   const float* __restrict__ lse, bf16* __restrict__ dQ, int lddq, bf16* __restrict__ dK,
                   int lddk, bf16* __restrict__ dV, int lddv, int heads, int nq, int nkv, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sdO = sV + TILE_BYTES;
  uint8_t* sP = sdO + TILE_BYTES;   // [2 blocks of 64 queries][128 key rows][128 B]
  uint8_t* sdS = sP + PT_BYTES;
  float* sLse2 = reinterpret_cast<float*>(sdS + PT_BYTES);  // [256] lse * log2(e); +inf for padding rows
  float* sDelta = sLse2 + 256;                              // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* bar_load = bars;        // TMA: Q, K, V, dO landed
  uint64_t* bar_s = bars + 1;       // S^T, dP^T accumulators complete
  uint64_t* bar_sfree = bars + 2;   // math warps finished reading S^T / dP^T from TMEM (count 4)
  uint64_t* bar_pds = bars + 3;     // P^T / dS^T tiles written + fenced (count 4)
  uint64_t* bar_pdsfree = bars + 4; // gradient MMAs that read the tiles retired
  uint64_t* bar_dkv = bars + 5;     // dV, dK of this key tile complete
  uint64_t* bar_dkvfree = bars + 6; // dV, dK drained by the math warps (count 4)
  uint64_t* bar_dq = bars + 7;      // dQ complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % heads;
  const int b = blockIdx.x / heads;
  const int nhq = (nq + 127) / 128, nkt = (nkv + 127) / 128;
  auto nq_half = [&](int hq) { return min(128, ((nq - hq * 128) + 15) & ~15); };
  auto nk_tile = [&](int kt) { return min(128, ((nkv - kt * 128) + 15) & ~15); };

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
      tma_prefetch_desc(&to);
      mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_sfree, 8); mbar_init(bar_pds, 8);
      mbar_init(bar_pdsfree, 1); mbar_init(bar_dkv, 1); mbar_init(bar_dkvfree, 8); mbar_init(bar_dq, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t C_ST = 0, C_DPT = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

  if (warp == 0) {
    if (lane == 0) {
      // O is only needed for delta = rowsum(dO * O): it lands in the (not yet used) P^T tile
      mbar_expect_tx(bar_load, (uint32_t)((3 * nhq + 2 * nkt) * 128 * 128));
      for (int i = 0; i < nhq; ++i) {
        tma_load_3d(sQ + i * 128 * 128, &tq, bar_load, h * DH, i * 128, b);
        tma_load_3d(sdO + i * 128 * 128, &tdo, bar_load, h * DH, i * 128, b);
        tma_load_3d(sP + i * 128 * 128, &to, bar_load, h * DH, i * 128, b);
      }
      for (int i = 0; i < nkt; ++i) {
        tma_load_3d(sK + i * 128 * 128, &tk, bar_load, h * DH, i * 128, b);
        tma_load_3d(sV + i * 128 * 128, &tv, bar_load, h * DH, i * 128, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);   // dV, dK: A K-major (P^T / dS^T), B MN-major
      const uint32_t idesc_q = umma_idesc_bf16(128, DH, 1, 1);   // dQ   : A MN-major (dS^T as dS), B MN-major
      const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK), va = smem_u32(sV), oa = smem_u32(sdO);
      const uint32_t pa = smem_u32(sP), sa = smem_u32(sdS);
      mbar_wait(bar_load, 0);
      tc_fence_after();
      // score products of iteration `i` (key tile i / nhq, query half i % nhq)
      auto issue_scores = [&](int i) {
        const int kt = i / nhq, hq = i - kt * nhq;
        const uint32_t idesc_s = umma_idesc_bf16(128, nq_half(hq), 0, 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16(tmem_base + C_ST, umma_desc_sw128(ka + kt * 16384 + k * 32, 0, 1024),
                    umma_desc_sw128(qa + hq * 16384 + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16(tmem_base + C_DPT, umma_desc_sw128(va + kt * 16384 + k * 32, 0, 1024),
                    umma_desc_sw128(oa + hq * 16384 + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
      };
      issue_scores(0);
      int it = 0;
      for (int kt = 0; kt < nkt; ++kt) {
        const int nk = nk_tile(kt);
        for (int hq = 0; hq < nhq; ++hq, ++it) {
          const int nqh = nq_half(hq);
          // the math warps have turned S^T / dP^T of this iteration into the P^T / dS^T tiles (and are done with TMEM)
          mbar_wait(bar_pds, (uint32_t)(it & 1));
          mbar_wait(bar_sfree, (uint32_t)(it & 1));
          tc_fence_after();
          // next scores FIRST: the math warps start on them while this iteration's gradient products run
          if (it + 1 < nkt * nhq) issue_scores(it + 1);
          if (hq == 0 && kt > 0) {  // dV / dK accumulators are about to be overwritten: previous tile drained?
            mbar_wait(bar_dkvfree, (uint32_t)((kt - 1) & 1));
            tc_fence_after();
          }
          for (int s = 0; s < nqh / 16; ++s) {  // contraction over the queries of this half
            const uint32_t aoff = (uint32_t)((s >> 2) * 16384 + (s & 3) * 32);
            const uint64_t bq = umma_desc_sw128(qa + hq * 16384 + s * 2048, 0, 1024);
            const uint64_t bo = umma_desc_sw128(oa + hq * 16384 + s * 2048, 0, 1024);
            umma_bf16(tmem_base + C_DV, umma_desc_sw128(pa + aoff, 0, 1024), bo, idesc_g, (hq > 0 || s > 0) ? 1u : 0u);
            umma_bf16(tmem_base + C_DK, umma_desc_sw128(sa + aoff, 0, 1024), bq, idesc_g, (hq > 0 || s > 0) ? 1u : 0u);
          }
          for (int s = 0; s < nk / 16; ++s)  // contraction over the keys of this tile
            umma_bf16(tmem_base + C_DQ + hq * 64, umma_desc_sw128(sa + s * 2048, 16384, 1024),
                      umma_desc_sw128(ka + kt * 16384 + s * 2048, 0, 1024), idesc_q, (kt > 0 || s > 0) ? 1u : 0u);
          umma_commit(bar_pdsfree);
          if (hq == nhq - 1) umma_commit(bar_dkv);
        }
      }
      umma_commit(bar_dq);
    }
  } else {
    // ---------------------------------------------------------------- math + epilogue warps (2..5)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;               // the two warps of a lane quarter alternate 32-column chunks
    const int row = quarter * 32 + lane;            // key row inside the tile == TMEM lane
    const int mt = threadIdx.x - 64;                // 0..255
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t swz = (uint32_t)(row & 7);
    const float sl2 = scale * 1.4426950408889634f;
    // delta_i = sum_d dO[i,d] O[i,d] from the TMA-loaded tiles (both carry the same 128B swizzle, so equal physical
    // chunks hold equal columns; chunk order rotated by the row so that 8 neighbouring rows hit 8 bank groups);
    // lse2_i = lse_i log2(e).  One row per thread.
    mbar_wait(bar_load, 0);
    {
      const int i = mt;
      float acc = 0.f, l2 = INFINITY;
      if (i < nq) {
        const uint8_t* a = sdO + i * 128;
        const uint8_t* c = sP + i * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int ch = ((k ^ (i & 7)) << 4);
          const uint4 ra = *reinterpret_cast<const uint4*>(a + ch);
          const uint4 rc = *reinterpret_cast<const uint4*>(c + ch);
          const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wc[4] = {rc.x, rc.y, rc.z, rc.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 fa = unpack_bf16(wa[t]), fc = unpack_bf16(wc[t]);
            acc += fa.x * fc.x + fa.y * fc.y;
          }
        }
        l2 = lse[((long)b * heads + h) * nq + i] * 1.4426950408889634f;
      }
      sDelta[i] = acc;
      sLse2[i] = l2;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // delta / lse2 visible to all math warps

    // dV, dK of key tile kt: TMEM -> bf16 -> global rows of the keys.  Deferred by one iteration (run after the first
    // math pass of the NEXT key tile) so that the wait for the gradient products hides behind that pass.
    auto drain_dkv = [&](int kt) {
      mbar_wait(bar_dkv, (uint32_t)(kt & 1));
      tc_fence_after();
      {
        const int j = kt * 128 + row;
#pragma unroll
        for (int part = half * 2; part < half * 2 + 2; ++part) {  // half 0: dV cols 0-31, 32-63; half 1: dK
          uint32_t a[32];
          tmem_ld_32x32(t_row + (part < 2 ? C_DV : C_DK) + (part & 1) * 32, a);
          tmem_ld_wait();
          if (j < nkv) {
            bf16* dst = (part < 2 ? dV + ((long)b * nkv + j) * lddv : dK + ((long)b * nkv + j) * lddk) + h * DH + (part & 1) * 32;
#pragma unroll
            for (int e = 0; e < 32; e += 8)
              *reinterpret_cast<uint4*>(dst + e) = make_uint4(
                  pack_bf16(__uint_as_float(a[e]), __uint_as_float(a[e + 1])),
                  pack_bf16(__uint_as_float(a[e + 2]), __uint_as_float(a[e + 3])),
                  pack_bf16(__uint_as_float(a[e + 4]), __uint_as_float(a[e + 5])),
                  pack_bf16(__uint_as_float(a[e + 6]), __uint_as_float(a[e + 7])));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dkvfree);
    };
    int it = 0;
    for (int kt = 0; kt < nkt; ++kt) {
      const bool key_ok = (kt * 128 + row) < nkv;
      for (int hq = 0; hq < nhq; ++hq, ++it) {
        const int nqh = nq_half(hq);
        mbar_wait(bar_s, (uint32_t)(it & 1));
        tc_fence_after();
        bool tiles_free = (it == 0);  // the delta pass above is done with the O rows in the P^T tile (bar.sync)
        for (int c = half * 32; c < nqh; c += 64) {
          uint32_t st[32], dp[32];
          if (nqh - c >= 32) {
            tmem_ld_32x32(t_row + C_ST + c, st);
            tmem_ld_32x32(t_row + C_DPT + c, dp);
          } else {  // 16-column tail
            uint32_t s16[16], d16[16];
            tmem_ld_32x16(t_row + C_ST + c, s16);
            tmem_ld_32x16(t_row + C_DPT + c, d16);
#pragma unroll
            for (int j = 0; j < 16; ++j) { st[j] = s16[j]; dp[j] = d16[j]; st[j + 16] = 0u; dp[j + 16] = 0u; }
          }
          tmem_ld_wait();
          const int lim = min(32, nqh - c);
          uint32_t pp[16], dd[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
            if (key_ok && j < lim) {
              const int i = hq * 128 + c + j;
              p0 = ex2f(fmaf(__uint_as_float(st[j]), sl2, -sLse2[i]));
              p1 = ex2f(fmaf(__uint_as_float(st[j + 1]), sl2, -sLse2[i + 1]));
              d0 = scale * p0 * (__uint_as_float(dp[j]) - sDelta[i]);
              d1 = scale * p1 * (__uint_as_float(dp[j + 1]) - sDelta[i + 1]);
            }
            pp[j >> 1] = pack_bf16(p0, p1);
            dd[j >> 1] = pack_bf16(d0, d1);
          }
          if (!tiles_free) {  // the previous iteration's gradient products no longer read the tiles
            mbar_wait(bar_pdsfree, (uint32_t)((it - 1) & 1));
            tiles_free = true;
          }
          // 32 queries = 4 chunks of 16 B in this thread's row of the 64-query block (c / 64)
          const uint32_t blk = (uint32_t)(c >> 6) * 16384u + (uint32_t)row * 128u;
          const uint32_t ch0 = (uint32_t)((c & 63) >> 3);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (q4 * 8 < lim) {
              const uint32_t off = blk + (((ch0 + q4) ^ swz) << 4);
              *reinterpret_cast<uint4*>(sP + off) = make_uint4(pp[q4 * 4], pp[q4 * 4 + 1], pp[q4 * 4 + 2], pp[q4 * 4 + 3]);
              *reinterpret_cast<uint4*>(sdS + off) = make_uint4(dd[q4 * 4], dd[q4 * 4 + 1], dd[q4 * 4 + 2], dd[q4 * 4 + 3]);
            }
          }
        }
        // TMEM S^T / dP^T consumed -> the issuer may start the next pair; tiles written -> gradient MMAs may start
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar_sfree); mbar_arrive(bar_pds); }
        if (hq == 0 && kt > 0) drain_dkv(kt - 1);
      }
    }
    drain_dkv(nkt - 1);
    // dQ: rows = queries
    mbar_wait(bar_dq, 0);
    tc_fence_after();
    for (int hq = 0; hq < nhq; ++hq) {
      const int i = hq * 128 + row;
      {
        const int part = half;
        uint32_t a[32];
        tmem_ld_32x32(t_row + C_DQ + hq * 64 + part * 32, a);
        tmem_ld_wait();
        if (i < nq) {
          bf16* dst = dQ + ((long)b * nq + i) * lddq + h * DH + part * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 8)
            *reinterpret_cast<uint4*>(dst + e) = make_uint4(
                pack_bf16(__uint_as_float(a[e]), __uint_as_float(a[e + 1])),
                pack_bf16(__uint_as_float(a[e + 2]), __uint_as_float(a[e + 3])),
                pack_bf16(__uint_as_float(a[e + 4]), __uint_as_float(a[e + 5])),
                pack_bf16(__uint_as_float(a[e + 6]), __uint_as_float(a[e + 7])));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace


// =====================================================================================================================
// Backward, version 2: PERSISTENT, one CTA per SM, loads of the next (image, head) problem prefetched into the tile slots
// the current problem has already released; chunked score tiles double-buffered in TMEM so that the tensor pipe and the
// two math warp sets run concurrently.
//
//   warp 0 lane 0 : TMA producer.  Shared memory is a pool of 16 KB tile slots (128 token rows x 128 B, 128B swizzle);
//                   a problem owns up to 8 tiles (Q0 Q1 dO0 dO1 K0 K1 V0 V1).  Slots are handed out from a free list and
//                   recycled in the order the tiles die (K_kt / V_kt after key tile kt, Q_hq / dO_hq after the last key
//                   tile): the UMMA issuer commits one "released" mbarrier per tile.  11 input slots: the next problem's
//                   first tiles land while the current problem is still in its second key tile.
//   warp 1 lane 0 : UMMA issuer.  Work is cut into chunks of <= 64 queries (two per 128-query half):
//                     scores(c): S^T = K_kt Q_c^T, dP^T = V_kt dO_c^T        (M = 128 keys, N = chunk, K = dh)
//                     grads(c) : dV_kt += P^T dO_c (A = P^T from TMEM), dK_kt += dS^T Q_c (A = the smem tile, K-major),
//                                and after the last chunk of a half  dQ_hq += dS K_kt (A = the same tile read MN-major)
//                   issued as  scores(c+2) right behind grads(c): chunk buffers alternate, so while one math set works on
//                   chunk c+1 the pipe runs grads(c) and scores(c+2).
//   warps 2-3     : delta_i = sum_d dO[i,d] O[i,d] and lse_i log2(e) of the NEXT problem, straight from global memory into
//                   a double-buffered shared array (O never occupies a tile slot).
//   warps 4-7 / 8-11 : math set 0 / 1 (one key row per thread = one TMEM lane); set h owns chunk buffer h, i.e. every
//                   other chunk:  P^T = exp2(S^T sl2 - lse2[q])  -> bf16 back into TMEM over the S^T columns already
//                   consumed;  dS^T = P^T (dP^T - delta[q]) -> bf16, 128B-swizzled shared tile.  The softmax scale is
//                   applied when dK / dQ are written out.  Drains of dV / dK (per key tile) and dQ (per problem) are
//                   deferred behind the set's next chunk so that their wait for the tensor pipe is hidden.
// TMEM (512 columns): chunk buffer b: S^T [128 b, +64) and dP^T [128 b + 64, +64) | dV 256 | dK 320 | dQ_0 384 | dQ_1 448.
// =====================================================================================================================
namespace {

constexpr int B2_THREADS = 512;
// registers (512 x 128 at launch): producer / issuer / delta warpgroup 88, drain warpgroup 104, the two math warpgroups 160
// (measured: 272 -> 251 us on ViT-B against 96 / 80 / 168; summing the dqkv columns in the drain warps for the QKV bias
// gradient was tried and cost 55 us per call against 36 us for the separate column-sum pass: not kept)
constexpr int B2_REGS_AUX = 88, B2_REGS_DRAIN = 104, B2_REGS_MATH = 160;
constexpr int B2_SLOT = 16384;
constexpr int B2_KV_SLOTS = 4;                  // ring of key / value tiles   (slots 0-3)
constexpr int B2_QO_SLOTS = 5;                  // ring of query / dO tiles    (slots 4-8)
constexpr int B2_IN_SLOTS = B2_KV_SLOTS + B2_QO_SLOTS;   // then two dS^T buffers of two 64-query blocks each
constexpr uint32_t B2_DP_OFF = 64;   // dP^T columns of a chunk buffer sit this far behind its S^T columns
constexpr int B2_OFF_DS = B2_IN_SLOTS * B2_SLOT;
constexpr int B2_OFF_DELTA = B2_OFF_DS + 4 * B2_SLOT;          // float [4][128]: delta of the last four query tiles
constexpr int B2_OFF_LSE = B2_OFF_DELTA + 4 * 128 * 4;         // float [4][128]: lse * log2(e)
constexpr int B2_OFF_BAR = B2_OFF_LSE + 4 * 128 * 4;           // 40 mbarriers
constexpr int B2_OFF_TMEM = B2_OFF_BAR + 40 * 8;
constexpr int B2_OFF_OST = B2_OFF_TMEM + 64;                     // [4 drain warps][32 rows][64 B] output transposition
constexpr int SMEM_B2 = B2_OFF_OST + 4 * 2048 + 1024;
// barrier indices
constexpr int BB_FULL = 0;      // [9] per input slot: the tile landed
constexpr int BB_REL = 9;       // [9] per input slot: every MMA reading the tile retired
constexpr int BB_S = 18;        // [2] chunk buffer: scores complete
constexpr int BB_MATH = 20;     // [2] chunk buffer: P^T / dS^T written (4 warps)
constexpr int BB_DSFREE = 22;   // [2] dS^T buffer: the dQ product that read it retired
constexpr int BB_DKV = 24;      // dV, dK of a key tile complete
constexpr int BB_DKVFREE = 25;  // ... read out of TMEM (4 drain warps)
constexpr int BB_DQ = 26;       // [2] dQ accumulator (query tile & 1) complete
constexpr int BB_DQFREE = 28;   // [2] ... read out of TMEM (4 drain warps)
constexpr int BB_DFULL = 30;    // [4] delta / lse2 of a query tile written (2 warps)
constexpr int BB_DEMPTY = 34;   // [4] ... no longer needed (8 math warps)

// A tile of one of the two input rings: slot index and how often the slot has been used before (mbarrier phase).
struct B2Tile {
  int slot;
  uint32_t use;
};

// The chunk sequence of one CTA: problems bh = blockIdx.x, + gridDim.x, ...; per problem key tiles (<= 2), query tiles of
// 128 (any number when there is one key tile, <= 2 otherwise), two chunks per query tile.  Every role walks the same
// sequence with its own iterator.
//   key / value tiles: ring of 4 slots, tile number gkv = 2 nkt k + 2 kt (+1 for V)
//   query / dO tiles : ring of 5 slots, tile number gq = 2 nhq k + 2 hq (+1 for dO)
// Both rings are FIFOs: tiles are loaded in that order and die in that order (K_kt / V_kt after key tile kt, Q_hq / dO_hq
// after query tile hq of the LAST key tile).  The long-lived key tiles must not share a ring with the streaming query
// tiles: a query tile waiting for a key tile's slot would wait for the end of the problem it belongs to.
struct B2Iter {
  int bh, k;
  int kt, hq, part;
  int nkt, nhq, n_bh, stride;
  int len_fa, len_fb, len_la, len_lb;   // chunk lengths of a full query tile (64 + 64) and of the last one
  int q_slot;                           // ring position of Q_hq (dO_hq follows)
  uint32_t q_use;
  int q0_slot;                          // ... of Q_0 of this problem (key tile 1 walks the query tiles again)
  uint32_t q0_use;
  __device__ __forceinline__ void init(int first, int stride_, int n_bh_, int nq, int nkv) {
    bh = first; stride = stride_; n_bh = n_bh_; k = 0; kt = hq = part = 0;
    nkt = (nkv + 127) >> 7; nhq = (nq + 127) >> 7;
    const int last = ((nq - (nhq - 1) * 128) + 15) & ~15;
    len_fa = 64; len_fb = 64;
    len_la = min(last, ((last + 31) >> 5) << 4);   // 128 -> 64 + 64, 80 -> 48 + 32, 48 -> 32 + 16, 16 -> 16 + 0
    len_lb = last - len_la;
    q_slot = q0_slot = 0; q_use = q0_use = 0;
  }
  __device__ __forceinline__ bool valid() const { return bh < n_bh; }
  __device__ __forceinline__ bool last_tile() const { return hq == nhq - 1; }
  __device__ __forceinline__ int len() const { return last_tile() ? (part ? len_lb : len_la) : (part ? len_fb : len_fa); }
  __device__ __forceinline__ int qo() const { return part ? (last_tile() ? len_la : len_fa) : 0; }   // first query of the chunk inside its tile
  __device__ __forceinline__ bool last_in_half() const { return part == 1 || (last_tile() ? len_lb : len_fb) == 0; }
  __device__ __forceinline__ bool last_in_kt() const { return last_in_half() && last_tile(); }
  __device__ __forceinline__ bool last_in_problem() const { return last_in_kt() && kt == nkt - 1; }
  __device__ __forceinline__ bool first_in_kt() const { return hq == 0 && part == 0; }
  __device__ __forceinline__ bool dq_final() const { return last_in_half() && kt == nkt - 1; }   // dQ_hq complete, Q_hq / dO_hq dead
  // sequence number of the query tile (delta / lse2 buffer = & 3, dQ accumulator = hq & 1)
  __device__ __forceinline__ uint32_t qseq() const { return (uint32_t)(k * nhq + hq); }
  __device__ __forceinline__ B2Tile tile_k() const {   // V_kt: slot + 1, same use count (pairs never straddle the ring of 4)
    const uint32_t g = (uint32_t)(2 * (nkt * k + kt));
    return B2Tile{(int)(g & 3), g >> 2};
  }
  __device__ __forceinline__ B2Tile tile_q() const { return B2Tile{B2_KV_SLOTS + q_slot, q_use}; }
  __device__ __forceinline__ B2Tile tile_do() const {
    const int s = q_slot + 1;
    return s >= B2_QO_SLOTS ? B2Tile{B2_KV_SLOTS + s - B2_QO_SLOTS, q_use + 1} : B2Tile{B2_KV_SLOTS + s, q_use};
  }
  __device__ __forceinline__ void step_q() {
    q_slot += 2;
    if (q_slot >= B2_QO_SLOTS) { q_slot -= B2_QO_SLOTS; ++q_use; }
  }
  __device__ __forceinline__ void advance() {
    if (!last_in_half()) { part = 1; return; }
    part = 0;
    if (++hq < nhq) { step_q(); return; }
    hq = 0;
    if (++kt < nkt) { q_slot = q0_slot; q_use = q0_use; return; }   // second key tile: the same query tiles again
    kt = 0; bh += stride; ++k;
    step_q();
    q0_slot = q_slot; q0_use = q_use;
  }
};

// One W-column (32 or 16) step of a chunk for one key row: P^T = exp2(S^T sl2 - lse2) and dS^T = P^T (dP^T - delta), both
// as bf16 pairs back into TMEM over the columns just read (A operands of dV / dK); dS^T also into this row of the
// 128B-swizzled shared tile (dQ = dS K reads it MN-major).
//   t_s: TMEM address of the S^T columns (dP^T sits B2_DP_OFF columns further), t_p: where the P^T pairs go
//   l2a / dla: shared addresses of lse2 / delta of the step's first query;  ds_row: shared address of this row in
//   block 0 of the dS^T buffer;  q: first query of the step inside its 128-query tile
template <int W>
__device__ __forceinline__ void b2_step(uint32_t t_s, uint32_t t_p, uint32_t l2a, uint32_t dla, uint32_t ds_row, uint32_t q,
                                        uint32_t swz, float sl2, uint64_t* bar_dsfree, uint32_t dsfree_par, bool& ds_free) {
  uint32_t st[W], dp[W];
  tmem_ld_cols<W>(t_s, st);
  tmem_ld_cols<W>(t_s + B2_DP_OFF, dp);
  tmem_ld_wait();
  uint32_t pp[W / 2], dd[W / 2];
#pragma unroll
  for (int j = 0; j < W; j += 4) {
    const float4 l4 = lds_f4(l2a + j * 4);
    const float4 d4 = lds_f4(dla + j * 4);
    const float p0 = ex2f(fmaf(__uint_as_float(st[j]), sl2, -l4.x));
    const float p1 = ex2f(fmaf(__uint_as_float(st[j + 1]), sl2, -l4.y));
    const float p2 = ex2f(fmaf(__uint_as_float(st[j + 2]), sl2, -l4.z));
    const float p3 = ex2f(fmaf(__uint_as_float(st[j + 3]), sl2, -l4.w));
    pp[j >> 1] = pack_bf16(p0, p1);
    pp[(j >> 1) + 1] = pack_bf16(p2, p3);
    dd[j >> 1] = pack_bf16(p0 * (__uint_as_float(dp[j]) - d4.x), p1 * (__uint_as_float(dp[j + 1]) - d4.y));
    dd[(j >> 1) + 1] = pack_bf16(p2 * (__uint_as_float(dp[j + 2]) - d4.z), p3 * (__uint_as_float(dp[j + 3]) - d4.w));
  }
  if (W == 32) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(t_p),
        "r"(pp[0]), "r"(pp[1]), "r"(pp[2]), "r"(pp[3]), "r"(pp[4]), "r"(pp[5]), "r"(pp[6]), "r"(pp[7]),
        "r"(pp[W / 2 - 8]), "r"(pp[W / 2 - 7]), "r"(pp[W / 2 - 6]), "r"(pp[W / 2 - 5]), "r"(pp[W / 2 - 4]), "r"(pp[W / 2 - 3]),
        "r"(pp[W / 2 - 2]), "r"(pp[W / 2 - 1])
        : "memory");
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(t_p + B2_DP_OFF),
        "r"(dd[0]), "r"(dd[1]), "r"(dd[2]), "r"(dd[3]), "r"(dd[4]), "r"(dd[5]), "r"(dd[6]), "r"(dd[7]),
        "r"(dd[W / 2 - 8]), "r"(dd[W / 2 - 7]), "r"(dd[W / 2 - 6]), "r"(dd[W / 2 - 5]), "r"(dd[W / 2 - 4]), "r"(dd[W / 2 - 3]),
        "r"(dd[W / 2 - 2]), "r"(dd[W / 2 - 1])
        : "memory");
  } else {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t_p), "r"(pp[0]),
                 "r"(pp[1]), "r"(pp[2]), "r"(pp[3]), "r"(pp[4]), "r"(pp[5]), "r"(pp[6]), "r"(pp[7])
                 : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t_p + B2_DP_OFF), "r"(dd[0]),
                 "r"(dd[1]), "r"(dd[2]), "r"(dd[3]), "r"(dd[4]), "r"(dd[5]), "r"(dd[6]), "r"(dd[7])
                 : "memory");
  }
  if (!ds_free) {   // the dQ product that read this dS^T buffer two half-iterations ago has retired
    mbar_wait(bar_dsfree, dsfree_par);
    ds_free = true;
  }
#pragma unroll
  for (int g = 0; g < W / 8; ++g) {   // 8 queries = 16 B; query qq of the tile sits in 64-query block qq / 64
    const uint32_t qq = q + g * 8;
    sts_u4(ds_row + (qq >> 6) * 16384u + ((((qq & 63u) >> 3) ^ swz) << 4), dd[g * 4], dd[g * 4 + 1], dd[g * 4 + 2],
           dd[g * 4 + 3]);
  }
}

// 32 accumulator rows of a warp (one TMEM lane = one row per thread, 64 fp32 columns in a0 | a1) -> bf16 (x f) -> 32 rows
// of 128 bytes in global memory, `ld` elements apart, rows >= n_valid skipped.  Written straight from the registers every
// store instruction would touch 32 different lines with 16 bytes each; the warp transposes through 2 KB of shared memory
// (one 64-byte half row at a time, chunks XOR-swizzled by the row: no bank conflicts either way) so that 4 lanes write a
// contiguous 64 bytes and an instruction covers 8 rows.
__device__ __forceinline__ void b2_store_rows(const uint32_t (&a0)[32], const uint32_t (&a1)[32], float f, bf16* row0, long ld,
                                              int n_valid, uint32_t so, int lane) {
  const int sub = lane >> 2, ch = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t* a = half ? a1 : a0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts_u4(so + (uint32_t)lane * 64u + (uint32_t)((c ^ ((lane >> 1) & 3)) << 4),
             pack_bf16(__uint_as_float(a[8 * c]) * f, __uint_as_float(a[8 * c + 1]) * f),
             pack_bf16(__uint_as_float(a[8 * c + 2]) * f, __uint_as_float(a[8 * c + 3]) * f),
             pack_bf16(__uint_as_float(a[8 * c + 4]) * f, __uint_as_float(a[8 * c + 5]) * f),
             pack_bf16(__uint_as_float(a[8 * c + 6]) * f, __uint_as_float(a[8 * c + 7]) * f));
    __syncwarp();
#pragma unroll
    for (int r8 = 0; r8 < 4; ++r8) {
      const int r = r8 * 8 + sub;
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "r"(so + (uint32_t)r * 64u + (uint32_t)((ch ^ ((r >> 1) & 3)) << 4)));
      if (r < n_valid) *reinterpret_cast<uint4*>(row0 + (long)r * ld + half * 32 + ch * 8) = v;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(B2_THREADS, 1)
attn_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                    const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tdo,
                    const bf16* __restrict__ Og, int ldo, const bf16* __restrict__ dOg, int lddo,
                    const float* __restrict__ lse, bf16* __restrict__ dQ, int lddq, bf16* __restrict__ dK, int lddk,
                    bf16* __restrict__ dV, int lddv, int heads, int nq, int nkv, int n_bh, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sdS = smem + B2_OFF_DS;
  float* sDelta = reinterpret_cast<float*>(smem + B2_OFF_DELTA);
  float* sLse2 = reinterpret_cast<float*>(smem + B2_OFF_LSE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B2_OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + B2_OFF_TMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  B2_TRACE_DECL;
  const int nkt = (nkv + 127) >> 7, nhq = (nq + 127) >> 7;

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&tq); tma_prefetch_desc(&tk); tma_prefetch_desc(&tv); tma_prefetch_desc(&tdo);
      for (int i = 0; i < 18; ++i) mbar_init(bars + i, 1);        // full, released
      mbar_init(bars + BB_S, 1); mbar_init(bars + BB_S + 1, 1);
      mbar_init(bars + BB_MATH, 4); mbar_init(bars + BB_MATH + 1, 4);
      mbar_init(bars + BB_DSFREE, 1); mbar_init(bars + BB_DSFREE + 1, 1);
      mbar_init(bars + BB_DKV, 1); mbar_init(bars + BB_DKVFREE, 4);
      for (int i = 0; i < 2; ++i) { mbar_init(bars + BB_DQ + i, 1); mbar_init(bars + BB_DQFREE + i, 4); }
      for (int i = 0; i < 4; ++i) { mbar_init(bars + BB_DFULL + i, 2); mbar_init(bars + BB_DEMPTY + i, 8); }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t C_DV = 256, C_DK = 320, C_DQ = 384;

  // Warp roles.  The hardware arbiter prefers the HIGHEST warp id of a scheduler partition (warp id % 4), so the UMMA issuer,
  // whose instruction stream is the critical path, is warp 15; drain warps (bursty, latency-tolerant) are warps 0-3.
  if (warp >= 12) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(B2_REGS_AUX));
  if (warp == 14) {
    // ------------------------------------------------------------------------------------------------ TMA producer
    // Tiles in load order; before a slot is refilled its previous tile must have been released (one phase per use).
    if (lane == 0) {
      uint32_t gkv = 0;           // key / value tiles loaded so far (ring of 4)
      int qs = 0;                 // query / dO ring position and use count (ring of 5)
      uint32_t qu = 0;
      auto load = [&](int slot, uint32_t use, const CUtensorMap* map, int c0, int tok, int b) {
        if (use > 0) mbar_wait(bars + BB_REL + slot, (use - 1) & 1);
        mbar_expect_tx(bars + BB_FULL + slot, B2_SLOT);
        tma_load_3d(smem + slot * B2_SLOT, map, bars + BB_FULL + slot, c0, tok, b);
      };
      auto load_q = [&](const CUtensorMap* map, int c0, int tok, int b) {
        load(B2_KV_SLOTS + qs, qu, map, c0, tok, b);
        if (++qs == B2_QO_SLOTS) { qs = 0; ++qu; }
      };
      for (int bh = blockIdx.x; bh < n_bh; bh += gridDim.x) {
        const int b = bh / heads, c0 = (bh - b * heads) * DH;
        // the order the issuer first touches them: K0 V0, Q0 dO0, Q1 dO1, ..., then K1 V1
        load((int)(gkv & 3), gkv >> 2, &tk, c0, 0, b); ++gkv;
        load((int)(gkv & 3), gkv >> 2, &tv, c0, 0, b); ++gkv;
        for (int hq = 0; hq < nhq; ++hq) {
          load_q(&tq, c0, hq * 128, b);
          load_q(&tdo, c0, hq * 128, b);
        }
        if (nkt > 1) {
          load((int)(gkv & 3), gkv >> 2, &tk, c0, 128, b); ++gkv;
          load((int)(gkv & 3), gkv >> 2, &tv, c0, 128, b); ++gkv;
        }
      }
    }
  } else if (warp == 15) {
    // ------------------------------------------------------------------------------------------------ UMMA issuer
    // All 32 lanes walk the chunk sequence (everything below is warp-uniform); one elected lane issues the tcgen05 ops.
    {
      const uint32_t idesc_g = umma_idesc_bf16(128, DH, 0, 1);   // dV, dK: A = P^T / dS^T in TMEM, B MN-major
      const uint32_t idesc_q = umma_idesc_bf16(128, DH, 1, 1);   // dQ   : A MN-major (dS^T as dS), B MN-major
      const uint32_t sa = smem_u32(sdS), s0 = smem_u32(smem);
      auto nk_tile = [&](int kt) { return min(128, ((nkv - kt * 128) + 15) & ~15); };

      // tiles a chunk touches first: K_kt / V_kt at the first chunk of a key tile, Q_hq / dO_hq at the first chunk of a
      // query tile in key tile 0
      auto scores_ready = [&](const B2Iter& c) {
        bool ok = true;
        if (c.first_in_kt()) {
          const B2Tile t = c.tile_k();
          ok = mbar_try_wait(bars + BB_FULL + t.slot, t.use & 1) && mbar_try_wait(bars + BB_FULL + t.slot + 1, t.use & 1);
        }
        if (ok && c.kt == 0 && c.part == 0) {
          const B2Tile tq_ = c.tile_q(), to_ = c.tile_do();
          ok = mbar_try_wait(bars + BB_FULL + tq_.slot, tq_.use & 1) && mbar_try_wait(bars + BB_FULL + to_.slot, to_.use & 1);
        }
        return ok;
      };

      auto issue_scores = [&](const B2Iter& c, uint32_t cc) {
        const B2Tile tk_ = c.tile_k(), tq_ = c.tile_q(), to_ = c.tile_do();
        if (c.first_in_kt()) {
          mbar_wait(bars + BB_FULL + tk_.slot, tk_.use & 1);
          mbar_wait(bars + BB_FULL + tk_.slot + 1, tk_.use & 1);
        }
        if (c.kt == 0 && c.part == 0) {
          mbar_wait(bars + BB_FULL + tq_.slot, tq_.use & 1);
          mbar_wait(bars + BB_FULL + to_.slot, to_.use & 1);
        }
        tc_fence_after();
        B2_TRACE(16);
        const uint32_t buf = tmem_base + (cc & 1) * 128;
        const uint32_t ka = s0 + (uint32_t)tk_.slot * B2_SLOT, va = ka + B2_SLOT;
        const uint32_t qa = s0 + (uint32_t)tq_.slot * B2_SLOT + (uint32_t)c.qo() * 128u;
        const uint32_t oa = s0 + (uint32_t)to_.slot * B2_SLOT + (uint32_t)c.qo() * 128u;
        const uint32_t idesc_s = umma_idesc_bf16(128, c.len(), 0, 0);
        const uint64_t dk = umma_desc_sw128(ka, 0, 1024), dq = umma_desc_sw128(qa, 0, 1024);
        const uint64_t dv = umma_desc_sw128(va, 0, 1024), dd = umma_desc_sw128(oa, 0, 1024);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16(buf, dk + 2 * kk, dq + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);   // +32 B
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_bf16(buf + B2_DP_OFF, dv + 2 * kk, dd + 2 * kk, idesc_s, kk > 0 ? 1u : 0u);
          umma_commit(bars + BB_S + (cc & 1));
        }
        __syncwarp();
        B2_TRACE(13);
      };

      uint32_t n_kt_done = 0;        // key-tile iterations whose dV / dK are complete (drain counter)
      uint32_t n_half = 0;           // half-iterations whose dQ product has been issued (dS^T buffer = n_half & 1)
      uint32_t n_dq[2] = {0, 0};     // completed uses of each dQ accumulator
      auto issue_grads = [&](const B2Iter& c, uint32_t cc) {
        const uint32_t buf = tmem_base + (cc & 1) * 128;
        const B2Tile tk_ = c.tile_k(), tq_ = c.tile_q(), to_ = c.tile_do();
        const int qb = c.hq & 1;
        B2_TRACE(10);
        mbar_wait(bars + BB_MATH + (cc & 1), (cc >> 1) & 1);
        B2_TRACE(11);
        if (c.first_in_kt() && n_kt_done > 0) { mbar_wait(bars + BB_DKVFREE, (n_kt_done - 1) & 1); B2_TRACE(14); }
        const bool dq_now = c.last_in_half();
        const uint32_t ndq = qb ? n_dq[1] : n_dq[0];
        if (dq_now && c.kt == 0 && ndq > 0)   // the accumulator's previous query tile has been read out?
          { mbar_wait(bars + BB_DQFREE + qb, (ndq - 1) & 1); B2_TRACE(15); }
        tc_fence_after();
        const int st0 = c.qo() >> 4;   // first 16-query step of the chunk inside its tile
        const int nst = c.len() >> 4;
        // MN-major B tiles advance 2048 B (= 128 in descriptor units) per 16-row step
        const uint64_t d_o = umma_desc_sw128(s0 + (uint32_t)to_.slot * B2_SLOT, 0, 1024) + (uint64_t)(st0 * 128);
        const uint64_t d_q = umma_desc_sw128(s0 + (uint32_t)tq_.slot * B2_SLOT, 0, 1024) + (uint64_t)(st0 * 128);
        const uint64_t d_sq = umma_desc_sw128(sa + (n_half & 1) * 2 * B2_SLOT, 16384, 1024);   // dS^T read MN-major: +2048 B per key step
        const uint64_t d_k = umma_desc_sw128(s0 + (uint32_t)tk_.slot * B2_SLOT, 0, 1024);
        const int nks = nk_tile(c.kt) >> 4;
        const uint32_t first = c.first_in_kt() ? 0u : 1u;
        const bool rel_q = c.dq_final(), rel_k = c.last_in_kt();
        if (elect_one()) {
          // straight-line issue (uniform predicates instead of loops): the issuing thread, not the tensor pipe, is what
          // bounds a run of small MMAs (tools/probes/mma_probe.cu: 36 - 53 cycles each at N = 64)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk < nst) {
              const uint32_t acc = kk > 0 ? 1u : first;
              umma_bf16_ts(tmem_base + C_DV, buf + kk * 8, d_o + (uint64_t)(kk * 128), idesc_g, acc);
              umma_bf16_ts(tmem_base + C_DK, buf + B2_DP_OFF + kk * 8, d_q + (uint64_t)(kk * 128), idesc_g, acc);
            }
          }
          if (dq_now) {
            const uint32_t dq_acc = c.kt > 0 ? 1u : 0u;
#pragma unroll
            for (int st = 0; st < 8; ++st)   // contraction over the keys of this tile
              if (st < nks)
                umma_bf16(tmem_base + C_DQ + qb * 64, d_sq + (uint64_t)(st * 128), d_k + (uint64_t)(st * 128), idesc_q,
                          st > 0 ? 1u : dq_acc);
            umma_commit(bars + BB_DSFREE + (n_half & 1));
            if (rel_q) {   // dQ of this query tile is final; last use of its Q / dO tiles
              umma_commit(bars + BB_REL + tq_.slot);
              umma_commit(bars + BB_REL + to_.slot);
              umma_commit(bars + BB_DQ + qb);
            }
          }
          if (rel_k) {
            umma_commit(bars + BB_REL + tk_.slot);
            umma_commit(bars + BB_REL + tk_.slot + 1);
            umma_commit(bars + BB_DKV);
          }
        }
        __syncwarp();
        B2_TRACE(12);
        if (dq_now) ++n_half;
        if (rel_q) { if (qb) ++n_dq[1]; else ++n_dq[0]; }
        if (rel_k) ++n_kt_done;
      };

      // Scores run up to two chunks ahead of the gradient products.  Running ahead must never BLOCK: the tiles a score
      // product waits for may sit in ring slots whose previous occupants are released only by gradient products this warp
      // has not issued yet.  So a chunk whose tiles have not landed is issued ahead only if a non-blocking probe says so;
      // the blocking wait happens when there is nothing else left to issue.
      B2Iter sc, gr;
      sc.init(blockIdx.x, gridDim.x, n_bh, nq, nkv);
      gr = sc;
      uint32_t cs = 0, cg = 0;
      while (gr.valid()) {
        while (cs - cg < 2 && sc.valid() && (cs == cg || scores_ready(sc))) { issue_scores(sc, cs++); sc.advance(); }
        issue_grads(gr, cg++);
        gr.advance();
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------------ delta / lse2
    // delta_i = sum_d dO[i,d] O[i,d] and lse_i log2(e) per QUERY TILE, straight from global memory, up to three tiles
    // ahead of the math warps (4 buffers)
    const int t2 = threadIdx.x - 12 * 32;   // 0..63
    uint32_t seq = 0;
    for (int bh = blockIdx.x; bh < n_bh; bh += gridDim.x) {
      const int b = bh / heads, h = bh - b * heads;
      for (int hq = 0; hq < nhq; ++hq, ++seq) {
        const int buf = (int)(seq & 3);
        mbar_wait(bars + BB_DEMPTY + buf, ((seq >> 2) & 1) ^ 1);
        for (int r = t2; r < 128; r += 64) {
          const int i = hq * 128 + r;
          float acc = 0.f, l2 = INFINITY;
          if (i < nq) {
            const uint4* o4 = reinterpret_cast<const uint4*>(Og + ((long)b * nq + i) * ldo + h * DH);
            const uint4* g4 = reinterpret_cast<const uint4*>(dOg + ((long)b * nq + i) * lddo + h * DH);
#pragma unroll
            for (int jj = 0; jj < 8; jj += 4) {
              uint4 ro[4], rg[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) { ro[j] = __ldg(o4 + jj + j); rg[j] = __ldg(g4 + jj + j); }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t wo[4] = {ro[j].x, ro[j].y, ro[j].z, ro[j].w}, wg[4] = {rg[j].x, rg[j].y, rg[j].z, rg[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 fo = unpack_bf16(wo[e]), fg = unpack_bf16(wg[e]);
                  acc = fmaf(fo.x, fg.x, acc);
                  acc = fmaf(fo.y, fg.y, acc);
                }
              }
            }
            l2 = lse[(long)bh * nq + i] * 1.4426950408889634f;
          }
          sts_f32(smem_u32(sDelta) + (uint32_t)(buf * 128 + r) * 4u, acc);   // padding queries: delta 0, lse2 +inf -> P = 0, dS = 0
          sts_f32(smem_u32(sLse2) + (uint32_t)(buf * 128 + r) * 4u, l2);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BB_DFULL + buf);
      }
    }
  }
  } else if (warp < 4) {
    // ------------------------------------------------------------------------------------------------ drain warps
    // dV / dK of every key tile and dQ of every query tile: TMEM -> bf16 -> global, one accumulator row (128 B) per thread,
    // in the order the issuer commits them.  Dedicated warps: the accumulators are handed back to the issuer a few hundred
    // cycles after they complete, and the math sets never stall on the tensor pipe.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(B2_REGS_DRAIN));
    const int quarter = warp & 3;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t so = smem_u32(smem + B2_OFF_OST) + (uint32_t)warp * 2048u;
    uint32_t n_kt = 0, n_dq0 = 0, n_dq1 = 0;
    B2Iter c;
    c.init(blockIdx.x, gridDim.x, n_bh, nq, nkv);
    for (; c.valid(); c.advance()) {
      const int b = c.bh / heads, h = c.bh - b * heads;
      if (c.dq_final()) {
        const int qb = c.hq & 1;
        const uint32_t n = qb ? n_dq1 : n_dq0;
        mbar_wait(bars + BB_DQ + qb, n & 1);
        tc_fence_after();
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(t_row + C_DQ + qb * 64, a0);
        tmem_ld_32x32(t_row + C_DQ + qb * 64 + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BB_DQFREE + qb);
        if (qb) ++n_dq1; else ++n_dq0;
        const int i0 = c.hq * 128 + quarter * 32;   // first query row of this warp
        b2_store_rows(a0, a1, scale, dQ + ((long)b * nq + i0) * lddq + h * DH, lddq, nq - i0, so, lane);
      }
      if (c.last_in_kt()) {
        mbar_wait(bars + BB_DKV, n_kt & 1);
        ++n_kt;
        tc_fence_after();
        const int j0 = c.kt * 128 + quarter * 32;   // first key row of this warp
        uint32_t a0[32], a1[32];
        tmem_ld_32x32(t_row + C_DV, a0);
        tmem_ld_32x32(t_row + C_DV + 32, a1);
        tmem_ld_wait();
        b2_store_rows(a0, a1, 1.f, dV + ((long)b * nkv + j0) * lddv + h * DH, lddv, nkv - j0, so, lane);
        tmem_ld_32x32(t_row + C_DK, a0);
        tmem_ld_32x32(t_row + C_DK + 32, a1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BB_DKVFREE);
        b2_store_rows(a0, a1, scale, dK + ((long)b * nkv + j0) * lddk + h * DH, lddk, nkv - j0, so, lane);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------------ math sets
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(B2_REGS_MATH));
    const int set = (warp - 4) >> 2;                // owns chunk buffer `set`
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;            // key row inside the tile == TMEM lane
    const uint32_t t_buf = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)set * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    const float sl2 = scale * 1.4426950408889634f;

    B2Iter c;
    c.init(blockIdx.x, gridDim.x, n_bh, nq, nkv);
    uint32_t cc = 0, n_half = 0;
    for (; c.valid(); c.advance(), ++cc) {
      const uint32_t seq = c.qseq();
      const int dbuf = (int)(seq & 3);
      if ((int)(cc & 1) == set) {
        mbar_wait(bars + BB_DFULL + dbuf, (seq >> 2) & 1);   // delta / lse2 of this query tile (re-waiting a done phase is free)
        B2_TRACE(20);
        mbar_wait(bars + BB_S + set, (cc >> 1) & 1);
        tc_fence_after();
        B2_TRACE(21);
        const int nk16 = min(128, ((nkv - c.kt * 128) + 15) & ~15);
        if (quarter * 32 < nk16) {   // warps whose 32 key rows are all padding only keep the barriers moving
          const int len = c.len(), qo = c.qo();
          const uint32_t l2a = smem_u32(sLse2) + (uint32_t)(dbuf * 128 + qo) * 4u;
          const uint32_t dla = smem_u32(sDelta) + (uint32_t)(dbuf * 128 + qo) * 4u;
          const uint32_t ds_row = smem_u32(sdS) + (n_half & 1) * 2 * B2_SLOT + (uint32_t)row * 128u;
          // dS^T buffer n_half & 1 was last read by the dQ product of half-iteration n_half - 2
          bool ds_free = (n_half < 2);
          uint64_t* bar_ds = bars + BB_DSFREE + (n_half & 1);
          const uint32_t ds_par = ((n_half >> 1) + 1) & 1;
          int c0 = 0;
          for (; c0 + 32 <= len; c0 += 32)
            b2_step<32>(t_buf + c0, t_buf + (c0 >> 1), l2a + c0 * 4, dla + c0 * 4, ds_row, (uint32_t)(qo + c0), swz, sl2, bar_ds,
                        ds_par, ds_free);
          if (c0 < len)
            b2_step<16>(t_buf + c0, t_buf + (c0 >> 1), l2a + c0 * 4, dla + c0 * 4, ds_row, (uint32_t)(qo + c0), swz, sl2, bar_ds,
                        ds_par, ds_free);
          B2_TRACE(24);
          tmem_st_wait();
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BB_MATH + set);
        B2_TRACE(22);
      }
      if (c.last_in_half()) ++n_half;
      if (c.dq_final()) {   // last use of this query tile's delta / lse2 (by either set)
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + BB_DEMPTY + dbuf);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool vtb_attn_tc_bwd_ok(const vtb_attn_params* p) {
  return g_attn_tc && p->mode == VTB_ATTN_GLOBAL && p->dh == DH && p->nkv <= 256 &&
         (p->nq <= 256 || (p->nkv <= 128 && g_attn_tc_bwd_version != 1)) &&
         !p->rel_bias && !p->mask && !p->dkv_f32 && p->lddq % 8 == 0 && p->lddk % 8 == 0 && p->lddv % 8 == 0 &&
         p->ldo % 8 == 0 && p->lddo % 8 == 0 &&
         ((((uintptr_t)p->dq) | ((uintptr_t)p->dk) | ((uintptr_t)p->dv) | ((uintptr_t)p->o) | ((uintptr_t)p->dout)) & 15) == 0;
}

int vtb_attn_tc_bwd(const vtb_attn_params* p, cudaStream_t stream) {
  if (int rc0 = ensure_encode()) return rc0;
  const uint64_t cols = (uint64_t)p->heads * DH;
  CUtensorMap tq, tk, tv, tdo, to;
  int rc;
  if ((rc = make_tmap3(&tq, p->q, cols, p->nq, p->batch, p->ldq, 128))) return rc;
  if ((rc = make_tmap3(&tk, p->k, cols, p->nkv, p->batch, p->ldk, 128))) return rc;
  if ((rc = make_tmap3(&tv, p->v, cols, p->nkv, p->batch, p->ldv, 128))) return rc;
  if ((rc = make_tmap3(&tdo, p->dout, cols, p->nq, p->batch, p->lddo, 128))) return rc;
  static bool attr = false;
  if (!attr) {
    VTB_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD));
    VTB_CUDA(cudaFuncSetAttribute(attn_tc_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_B2));
    attr = true;
  }
  const long blocks = (long)p->batch * p->heads;
  VTB_CHECK(blocks < (1L << 30), -1, "vtb_attention_bwd: grid too large");
  if (g_attn_tc_bwd_version == 1) {
    if ((rc = make_tmap3(&to, p->o, cols, p->nq, p->batch, p->ldo, 128))) return rc;
    attn_tc_bwd_kernel<<<(unsigned)blocks, BWD_THREADS, SMEM_BWD, stream>>>(
        tq, tk, tv, tdo, to, p->lse, reinterpret_cast<bf16*>(p->dq), p->lddq, reinterpret_cast<bf16*>(p->dk), p->lddk,
        reinterpret_cast<bf16*>(p->dv), p->lddv, p->heads, p->nq, p->nkv, p->scale);
    VTB_LAUNCH_CHECK();
    return 0;
  }
  const int grid = (int)(blocks < vtb_num_sms() ? blocks : vtb_num_sms());
  attn_tc_bwd2_kernel<<<grid, B2_THREADS, SMEM_B2, stream>>>(
      tq, tk, tv, tdo, reinterpret_cast<const bf16*>(p->o), p->ldo, reinterpret_cast<const bf16*>(p->dout), p->lddo,
      p->lse, reinterpret_cast<bf16*>(p->dq), p->lddq, reinterpret_cast<bf16*>(p->dk), p->lddk,
      reinterpret_cast<bf16*>(p->dv), p->lddv, p->heads, p->nq, p->nkv, (int)blocks, p->scale);
  VTB_LAUNCH_CHECK();
  return 0;
}

#ifdef VTB_ATTN_TRACE
extern "C" int vtb_debug_attn_trace(uint32_t* out) {   // uint32 [16 warps][1024][2]; event id 0 = unused entry
  VTB_CUDA(cudaMemcpyFromSymbol(out, g_b2_trace, sizeof(unsigned int) * 16 * 2 * 1024));
  static unsigned int zeros[16 * 2 * 1024];
  VTB_CUDA(cudaMemcpyToSymbol(g_b2_trace, zeros, sizeof(zeros)));
  return 0;
}
#endif

#ifdef VTB_MBAR_DEBUG
// debug builds only: {shared address of the barrier, parity, block, thread} of the first mbarrier wait that timed out
extern "C" int vtb_debug_mbar_timeout(uint32_t* out) {
  VTB_CUDA(cudaMemcpyFromSymbol(out, g_mbar_timeout, 16));
  const uint32_t zero[4] = {0, 0, 0, 0};
  VTB_CUDA(cudaMemcpyToSymbol(g_mbar_timeout, zero, 16));
  return 0;
}
#endif
