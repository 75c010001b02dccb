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

}  // namespace

void vtb_attn_tc_set(bool on) { g_attn_tc = on; }

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
    attr = true;
  }
  const int q_tiles = (p->nq + TQ - 1) / TQ;
  const long blocks = (long)p->batch * p->heads * q_tiles;
  VTB_CHECK(blocks < (1L << 31), -1, "vtb_attention_fwd: grid too large");
  attn_tc_fwd_kernel<<<(unsigned)blocks, TC_THREADS, SMEM_ATT, stream>>>(
      tq, tk, tv, reinterpret_cast<bf16*>(p->o), p->ldo, p->lse, p->heads, p->nq, p->nkv, q_tiles, p->scale);
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

bool vtb_attn_tc_bwd_ok(const vtb_attn_params* p) {
  return g_attn_tc && p->mode == VTB_ATTN_GLOBAL && p->dh == DH && p->nkv <= 256 && p->nq <= 256 &&
         !p->rel_bias && !p->mask && !p->dkv_f32 && p->lddq % 8 == 0 && p->lddk % 8 == 0 && p->lddv % 8 == 0 &&
         p->ldo % 8 == 0 && p->lddo % 8 == 0 &&
         ((((uintptr_t)p->dq) | ((uintptr_t)p->dk) | ((uintptr_t)p->dv) | ((uintptr_t)p->o) | ((uintptr_t)p->dout)) & 15) == 0;
}

int vtb_attn_tc_bwd(const vtb_attn_params* p, cudaStream_t stream) {
  if (int rc0 = ensure_encode()) return rc0;
  const uint64_t cols = (uint64_t)p->heads * DH;
  CUtensorMap tq, tk, tv, tdo, to;
  int rc;
  if ((rc = make_tmap3(&to, p->o, cols, p->nq, p->batch, p->ldo, 128))) return rc;
  if ((rc = make_tmap3(&tq, p->q, cols, p->nq, p->batch, p->ldq, 128))) return rc;
  if ((rc = make_tmap3(&tk, p->k, cols, p->nkv, p->batch, p->ldk, 128))) return rc;
  if ((rc = make_tmap3(&tv, p->v, cols, p->nkv, p->batch, p->ldv, 128))) return rc;
  if ((rc = make_tmap3(&tdo, p->dout, cols, p->nq, p->batch, p->lddo, 128))) return rc;
  static bool attr = false;
  if (!attr) {
    VTB_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BWD));
    attr = true;
  }
  const long blocks = (long)p->batch * p->heads;
  VTB_CHECK(blocks < (1L << 31), -1, "vtb_attention_bwd: grid too large");
  attn_tc_bwd_kernel<<<(unsigned)blocks, BWD_THREADS, SMEM_BWD, stream>>>(
      tq, tk, tv, tdo, to, p->lse, reinterpret_cast<bf16*>(p->dq), p->lddq, reinterpret_cast<bf16*>(p->dk), p->lddk,
      reinterpret_cast<bf16*>(p->dv), p->lddv, p->heads, p->nq, p->nkv, p->scale);
  VTB_LAUNCH_CHECK();
  return 0;
}
