// tcgen05 GEMM for sm_100a: persistent, warp-specialised.
//   warp 0      : TMA producer (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      : UMMA issuer  (tcgen05.mma cta_group::1, M=128, N=BN, K=16; fp32 accumulators in TMEM)
//   warp 2      : TMEM allocator
//   warp 3      : epilogue TMA helper: lanes 0-3 issue the TMA stores / aux prefetches of the four TMEM lane quarters
//                 (a bulk-tensor instruction costs its issuing thread 250-800 cycles, measured: off the math warps)
//   warps 4..11 : epilogue (two warps per TMEM lane quarter, each taking half the columns): tcgen05.ld -> registers -> fused epilogue -> 128B-swizzled smem staging ->
//                 TMA store (cp.async.bulk.tensor; cp.reduce.async.bulk .add for split-K); residual /
//                 pre-activation tiles are TMA-prefetched into smem three sub-tiles ahead.
//                 (Unaligned outputs fall back to direct per-thread global stores.)
// Two TMEM accumulator stages let the epilogue of tile i overlap the mainloop of tile i+1.
// Operands may be K-major or MN-major (wgrad / dgrad read the same buffers the forward wrote,
// no transposes are materialised).  See include/vtb200.h for the contract.
#include "common.cuh"
#include "../../include/vtb200.h"
#include <stdlib.h>
#include <string.h>
#include <type_traits>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int EPI_BUF_BYTES = BM * 128;     // one staged sub-tile: 128 rows x 128 B (64 bf16 or 32 f32 columns)
constexpr int N_AUX = 3;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
#ifdef VTB_GEMM_TRACE
__device__ unsigned long long g_trace[32];   // [0,16): epilogue phases per sub-tile; [16,32): tile-level waits (see tools/trace_gemm.py)
#define TRACE_T(i) const long long tr##i = clock64()
#define TRACE_ADD(slot, expr) do { if (blockIdx.x == 7) atomicAdd(&g_trace[slot], (unsigned long long)(expr)); } while (0)
#else
#define TRACE_T(i)
#define TRACE_ADD(slot, expr)
#endif
#ifdef VTB_GEMM_DBG
int g_gemm_dbg = 0;         // knock-out probes of the staged epilogue (debug build only): 1 no bias loads, 2 no staging stores,
                              // 4 no TMA stores, 8 no epilogue math at all, 16 no tcgen05.ld of the next sub-tile,
                              // 32 / 64 no B / A operand loads on CTA-pair tiles (stale smem: timing only)
#define GDBG(e, bit) (((e).dbg & (bit)) != 0)
#else
#define GDBG(e, bit) false
#endif
int g_bn_waste_pct = 120;   // tile-N choice: a wider tile may waste this much more of the MMA columns than the next narrower one
int g_helpers = 2;          // warps issuing the staged epilogue's TMA stores (vtb_set_option("gemm_helpers", 1 | 2))
int g_colsum_pair = 1;      // a_colsum launches may use CTA pairs (vtb_set_option("gemm_colsum_pair", 0): 1-CTA tiles as in round 1)
int g_use_clusters = 1;     // CTA-pair (cta_group::2) tiles wherever legal; vtb_set_option("gemm_cluster", 0) / VTB_GEMM_CLUSTER=0 forces
                              // 1-CTA tiles, 2 = 1 (kept for the tests' parametrisation), 3 = pairs only when pairs x splits fill the slots (round 1)

struct EpiParams {
  int M, N;
  void* out;
  int ldo;
  int out_f32;
  bf16* out2;
  const float* bias;
  const float* resid;
  int ldr;
  const float* row_scale;
  int rows_per_scale;
  const bf16* aux;
  int ldaux;
  int epilogue;
  int accumulate;
  float alpha;
  int vec;  // all epilogue pointers / leading dims allow 16-byte vector access (direct path)
  int tma;  // 1: staged TMA-store epilogue (tensor maps valid)
  int dbg;          // VTB_GEMM_DBG builds only
  int helpers;      // 1 | 2: warps that issue the epilogue's bulk-tensor instructions (see the helper role)
  float* a_colsum;  // MN-major A only: += column sums of the A operand (bias gradient riding on a wgrad)
};

template <int BN, int CL>
struct Cfg {
  // CL = 2: CTA pair (cta_group::2), each CTA stages its 128 rows of A and HALF of the B tile
  static constexpr int B_STAGE_BYTES = (BN / CL) * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int BAR_BYTES = 1536;  // mbarriers (< 1 KB) + 512 B scratch for the a_colsum reduction
  static constexpr int BUDGET = 227 * 1024 - 1024 /*align*/ - BAR_BYTES;
  static constexpr int clamp6(int v) { return v > 6 ? 6 : v; }
  // epilogue staging: 2 (or 3) output buffers + 3 aux buffers (second output, or prefetched residual / pre-activation);
  // the third output buffer is taken only where it does not cost a ring stage
  static constexpr int STAGES = clamp6((BUDGET - (2 + 3) * EPI_BUF_BYTES) / STAGE_BYTES);  // 3/4/6 (CL=1), 4/6/6 (CL=2)
  static constexpr int N_OUT = (clamp6((BUDGET - (3 + 3) * EPI_BUF_BYTES) / STAGE_BYTES) == STAGES) ? 3 : 2;
  static constexpr int STAGING_BYTES = (N_OUT + 3) * EPI_BUF_BYTES;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // 512 / 256 / 128
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + BAR_BYTES;
};

// One 32-column chunk of one accumulator row -> global memory with the fused epilogue.
__device__ __forceinline__ void epilogue_chunk(const EpiParams& e, const uint32_t (&acc)[32],
                                               int m, int n0) {
  if (m >= e.M) return;
  const long m_out = m;
  const float rs = e.row_scale ? __ldg(e.row_scale + m / e.rows_per_scale) : 1.f;
  const bool full = e.vec && (n0 + 32 <= e.N);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * e.alpha;
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (full || n0 + j < e.N) v[j] += __ldg(e.bias + n0 + j);
  }
  if (e.epilogue == VTB_EPI_SILU_DUAL) {
    // out <- bf16(u); v <- silu(float(bf16(u))) goes to out2  (layer.py:191-193 under autocast)
    bf16* o1 = reinterpret_cast<bf16*>(e.out) + m_out * e.ldo + n0;
    bf16* o2 = e.out2 + m_out * e.ldo + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t pu[4], ph[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          pu[q] = pack_bf16(v[j + 2 * q], v[j + 2 * q + 1]);
          float2 ur = unpack_bf16(pu[q]);
          ph[q] = pack_bf16(silu_f(ur.x), silu_f(ur.y));
        }
        *reinterpret_cast<uint4*>(o1 + j) = make_uint4(pu[0], pu[1], pu[2], pu[3]);
        *reinterpret_cast<uint4*>(o2 + j) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n0 + j < e.N) {
        bf16 u = __float2bfloat16(v[j]);
        o1[j] = u;
        o2[j] = __float2bfloat16(silu_f(__bfloat162float(u)));
      }
    }
    return;
  }
  if (e.epilogue == VTB_EPI_SILU_GRAD) {
    const bf16* a = e.aux + (long)m * e.ldaux + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 raw = *reinterpret_cast<const uint4*>(a + j);
        uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 u = unpack_bf16(w[q]);
          v[j + 2 * q] *= silu_grad_f(u.x);
          v[j + 2 * q + 1] *= silu_grad_f(u.y);
        }
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n0 + j < e.N) v[j] *= silu_grad_f(__bfloat162float(a[j]));
    }
  }
  if (e.row_scale) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rs;
  }
  if (e.resid) {
    const float* pr = e.resid + m_out * e.ldr + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 t = *reinterpret_cast<const float4*>(pr + j);
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n0 + j < e.N) v[j] += pr[j];
    }
  }
  if (e.out_f32) {
    float* o = reinterpret_cast<float*>(e.out) + m_out * e.ldo + n0;
    if (e.accumulate) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (full || n0 + j < e.N) atomicAdd(o + j, v[j]);
    } else if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n0 + j < e.N) o[j] = v[j];
    }
  } else {
    bf16* o = reinterpret_cast<bf16*>(e.out) + m_out * e.ldo + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        *reinterpret_cast<uint4*>(o + j) =
            make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                       pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
      }
    } else {
      _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n0 + j < e.N) o[j] = __float2bfloat16(v[j]);
    }
  }
}

// Staged epilogue of one warp's share (CW columns) of a sub-tile row: TMEM registers -> fused math ->
// swizzled smem staging (chunks cb..cb+3 of the 128-byte row).  ob/ab point at this thread's row.
// Explicit shared-space accesses for the staging slabs: `ob` / `ab` are carved out of the dynamic shared block after an
// integer round-up, which makes them GENERIC pointers to the compiler — their loads / stores were LD.E / ST.E through the
// L1TEX path (long-scoreboard latency) instead of LDS / STS.
__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f4(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// bias_vec: this warp's CW bias values as 16-byte read-only loads at a warp-uniform address (L1 broadcast; the line was
// prefetched a sub-tile ahead) instead of CW shuffles — or nullptr (ragged N / unaligned bias): lane j of the warp then holds
// the bias of column j in bias_lane.  All element-wise math runs on packed fp32 pairs (FFMA2 / FADD2 / FMUL2).
template <int CW>
__device__ __forceinline__ void staged_row(const EpiParams& e, const uint32_t (&acc)[CW], float bias_lane,
                                           const float* __restrict__ bias_vec, float rs,
                                           uint32_t ob, uint32_t ab, uint32_t cb, uint32_t swz, bool dual,
                                           bool f32out) {
  uint64_t v[CW / 2];
#pragma unroll
  for (int j = 0; j < CW / 2; ++j) v[j] = f2_pack(__uint_as_float(acc[2 * j]), __uint_as_float(acc[2 * j + 1]));
  if (e.alpha != 1.f) {
    const uint64_t al = f2_pack(e.alpha, e.alpha);
#pragma unroll
    for (int j = 0; j < CW / 2; ++j) v[j] = f2_mul(v[j], al);
  }
  if (e.bias && !GDBG(e, 1)) {
    if (bias_vec) {
#pragma unroll
      for (int j = 0; j < CW; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_vec + j));
        v[j / 2] = f2_add(v[j / 2], f2_pack(b4.x, b4.y));
        v[j / 2 + 1] = f2_add(v[j / 2 + 1], f2_pack(b4.z, b4.w));
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; j += 2)
        v[j / 2] = f2_add(v[j / 2], f2_pack(__shfl_sync(0xffffffffu, bias_lane, j), __shfl_sync(0xffffffffu, bias_lane, j + 1)));
    }
  }
  if (dual) {
    // out <- bf16(u) ; out2 <- bf16(silu(float(bf16(u))))     (layer.py:191-193 under autocast)
#pragma unroll
    for (int j = 0; j < CW; j += 8) {
      uint32_t pu[4], ph[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float u0, u1, h0, h1;
        f2_unpack(v[j / 2 + t], u0, u1);
        pu[t] = pack_bf16(u0, u1);
        silu_pair(bf16lo_f(pu[t]), bf16hi_f(pu[t]), h0, h1);
        ph[t] = pack_bf16(h0, h1);
      }
      const uint32_t off = ((cb + j / 8) ^ swz) << 4;
      if (!GDBG(e, 2)) {
        sts_u4(ob + off, pu[0], pu[1], pu[2], pu[3]);
        sts_u4(ab + off, ph[0], ph[1], ph[2], ph[3]);
      }
    }
    return;
  }
  if (e.epilogue == VTB_EPI_SILU_GRAD) {
#pragma unroll
    for (int j = 0; j < CW; j += 8) {
      const uint4 raw = lds_u4(ab + (((cb + j / 8) ^ swz) << 4));
      const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float g0, g1;
        silu_grad_pair(bf16lo_f(w[t]), bf16hi_f(w[t]), g0, g1);
        v[j / 2 + t] = f2_mul(v[j / 2 + t], f2_pack(g0, g1));
      }
    }
  }
  if (e.row_scale) {
    const uint64_t r2 = f2_pack(rs, rs);
#pragma unroll
    for (int j = 0; j < CW / 2; ++j) v[j] = f2_mul(v[j], r2);
  }
  if (f32out) {
    if (e.resid) {  // f32 residual sub-tile prefetched by TMA
#pragma unroll
      for (int j = 0; j < CW; j += 4) {
        const float4 t = lds_f4(ab + (((cb + j / 4) ^ swz) << 4));
        v[j / 2] = f2_add(v[j / 2], f2_pack(t.x, t.y));
        v[j / 2 + 1] = f2_add(v[j / 2 + 1], f2_pack(t.z, t.w));
      }
    }
#pragma unroll
    for (int j = 0; j < CW; j += 4) {
      float a0, a1, a2, a3;
      f2_unpack(v[j / 2], a0, a1);
      f2_unpack(v[j / 2 + 1], a2, a3);
      if (!GDBG(e, 2)) sts_f4(ob + (((cb + j / 4) ^ swz) << 4), a0, a1, a2, a3);
    }
  } else {
#pragma unroll
    for (int j = 0; j < CW; j += 8) {
      uint32_t pk[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float a0, a1;
        f2_unpack(v[j / 2 + t], a0, a1);
        pk[t] = pack_bf16(a0, a1);
      }
      if (!GDBG(e, 2)) sts_u4(ob + (((cb + j / 8) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
      else if (pk[0] == 0x12345678u && pk[3] == 0x9abcdef0u) sts_u4(ob, pk[0], pk[1], pk[2], pk[3]);  // keep the math alive
    }
  }
}

// CL = CTAs per cluster along M (1 or 2).  CL = 2 is the CTA-pair mode of the 5th-gen tensor cores: ONE
// tcgen05.mma.cta_group::2 (M = 256) issued by the leader CTA drives the tensor cores of both SMs on a 256 x BN output
// tile; each CTA stages its own 128 rows of A and HALF of the B tile (the halves are exchanged by the hardware), so
// per-SM operand traffic (L2 -> smem writes and smem -> tensor-core reads) drops by a third against two independent
// 128 x BN tiles and the ring gets deeper.  Both producers credit the LEADER's full barrier; the leader's
// tcgen05.commit multicasts stage release / accumulator-ready to both CTAs; the peer's epilogue warps hand the TMEM
// stage back with remote mbarrier arrives.
template <int BN, bool A_MN, bool B_MN, int CL>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_out2,
               const __grid_constant__ CUtensorMap tma_aux, int m_tiles, int n_tiles, int k_blocks,
               int splits, EpiParams epi) {
  using C = Cfg<BN, CL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * A_STAGE_BYTES;
  uint8_t* sOut = smem + C::STAGES * C::STAGE_BYTES;          // [2][EPI_BUF_BYTES]
  uint8_t* sAux = sOut + C::N_OUT * EPI_BUF_BYTES;            // [N_AUX][EPI_BUF_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tmem_full = bars + 2 * C::STAGES;
  uint64_t* tmem_empty = bars + 2 * C::STAGES + 2;
  uint64_t* aux_full = bars + 2 * C::STAGES + 4;              // [4 lane quarters][N_AUX]
  constexpr int MAXOB = 8;                                    // output staging buffers per quarter (<= N_OUT + N_AUX)
  uint64_t* staged_bar = aux_full + 4 * N_AUX;                // [4 quarters][MAXOB]: sub-tile written + fenced (2 warps)
  uint64_t* free_bar = staged_bar + 4 * MAXOB;                // [4 quarters][MAXOB]: TMA has read the staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(free_bar + 4 * MAXOB);
  uint64_t* mma_done = free_bar + 4 * MAXOB + 1;              // [STAGES], pair-mode a_colsum only (ends below the scratch at +1024)
  // Pair-mode bias gradient (`a_colsum` on cta_group::2 tiles): the operand bytes of both CTAs are credited to the
  // LEADER's full barrier, so the peer's readers have no local "tile landed" event.  The issuer's commit therefore goes to
  // `mma_done` (multicast to both CTAs), the readers add up the stage AFTER the tensor cores are through with it and
  // release it themselves (`empty` = 8 reader warps).  Six ring stages hide the later release; the k-blocks are dealt
  // round-robin over the n-tiles of a row of tiles, so every CTA reads 1 / n_tiles of its stages instead of one CTA in
  // n_tiles reading all of them.  (Tried and not kept: warp 2 as the only reader, so that the epilogue warps can drain
  // tile i while the ring serves tile i+1 — one warp has too few shared-memory loads in flight next to the tensor
  // cores' operand reads: ViT-B fc1 wgrad 237 us against 190 us with the 8 warps, profiles/r02_cabi_gemm_colsum_*.log.)
  const bool cs_post = (CL > 1) && A_MN && (epi.a_colsum != nullptr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], (A_MN && epi.a_colsum) ? (CL > 1 ? 8 : 9) : 1);  // + the 8 epilogue warps that read the A tiles
      if (CL > 1) mbar_init(&mma_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8 * CL);  // every epilogue warp of the pair arrives on the leader's barrier
    }
    for (int i = 0; i < 4 * N_AUX; ++i) mbar_init(&aux_full[i], 1);
    for (int i = 0; i < 4 * 8; ++i) { mbar_init(&staged_bar[i], 2); mbar_init(&free_bar[i], 1); }
    mbar_fence_init();
  }
  if (warp == 2) {
    if (CL == 1) tmem_alloc(tmem_slot, C::TMEM_COLS); else tmem_alloc_pair(tmem_slot, C::TMEM_COLS);
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // barrier inits visible cluster-wide before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work items are (m-group of CL tiles, n tile, k split); this CTA takes row `rank` of its cluster's group
  const int rank = (CL > 1) ? (int)(blockIdx.x % CL) : 0;
  const int cta = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;
  const int total_tiles = ((m_tiles + CL - 1) / CL) * n_tiles * splits;
  const int kb_per_split = (k_blocks + splits - 1) / splits;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cta; tile < total_tiles; tile += ncl) {
        const int ks = tile % splits;
        const int mn = tile / splits;
        const int n_blk = mn % n_tiles;
        const int m_blk = (mn / n_tiles) * CL + rank;
        const int kb0 = ks * kb_per_split;
        const int kb1 = min(k_blocks, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = sA + stage * A_STAGE_BYTES;
          uint8_t* b_dst = sB + stage * C::B_STAGE_BYTES;
          if (CL == 1) {
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d(a_dst, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_2d(a_dst + i * (BK * 128), &tma_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d(b_dst, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_2d(b_dst + i * (BK * 128), &tma_b, &full_bar[stage], n_blk * BN + i * 64, kb * BK);
            }
          } else {
            // the leader's barrier collects the bytes of BOTH CTAs (its expect_tx may race with the peer's
            // complete_tx: the phase cannot complete before the leader's own arrive)
            const bool no_b = GDBG(epi, 32), no_a = GDBG(epi, 64);   // operand-traffic probes (debug build)
            if (rank == 0)
              mbar_expect_tx(&full_bar[stage], ((no_a ? 0 : A_STAGE_BYTES) + (no_b ? 0 : C::B_STAGE_BYTES)) * CL);
            if (no_a) {
            } else if (!A_MN) {
              tma_load_2d_pair(a_dst, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_2d_pair(a_dst + i * (BK * 128), &tma_a, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
            }
            // this CTA's half of the B tile's N range
            if (no_b) {
            } else if (!B_MN) {
              tma_load_2d_pair(b_dst, &tma_b, &full_bar[stage], kb * BK, n_blk * BN + rank * (BN / CL));
            } else {
#pragma unroll
              for (int i = 0; i < BN / 64 / CL; ++i)
                tma_load_2d_pair(b_dst + i * (BK * 128), &tma_b, &full_bar[stage],
                                 n_blk * BN + (rank * (BN / 64 / CL) + i) * 64, kb * BK);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer (the leader CTA only in pair mode)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CL, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = cta; tile < total_tiles; tile += ncl) {
        const int ks = tile % splits;
        const int kb0 = ks * kb_per_split;
        const int kb1 = min(k_blocks, kb0 + kb_per_split);
#ifdef VTB_GEMM_TRACE
        const long long ti0 = clock64();
#endif
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        TRACE_ADD(20, clock64() - ti0);   // MMA issuer waiting for a free accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
#ifdef VTB_GEMM_TRACE
          const long long ti1 = clock64();
#endif
          mbar_wait(&full_bar[stage], phase);
          TRACE_ADD(19, clock64() - ti1);  // ... for operands
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * A_STAGE_BYTES);
          const uint32_t b_base = smem_u32(sB + stage * C::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major : atoms of 8 rows x 128 B, SBO = 1024 B, advance 32 B per UMMA_K inside the row.
            // MN-major: atoms of 8 k-rows x 64 mn (128 B), SBO = 1024 B between k-groups,
            //           LBO = BK*128 B between 64-wide mn atoms, advance 16 k-rows = 2048 B.
            const uint64_t adesc = A_MN ? umma_desc_sw128(a_base + k * (UMMA_K * 128), BK * 128, 1024)
                                        : umma_desc_sw128(a_base + k * (UMMA_K * 2), 0, 1024);
            const uint64_t bdesc = B_MN ? umma_desc_sw128(b_base + k * (UMMA_K * 128), BK * 128, 1024)
                                        : umma_desc_sw128(b_base + k * (UMMA_K * 2), 0, 1024);
            if (CL == 1) umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16_pair(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (CL == 1) umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          else umma_commit_pair(cs_post ? &mma_done[stage] : &empty_bar[stage], (uint16_t)0x3);  // ... in both CTAs of the pair
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs in pair mode)
        if (CL == 1) umma_commit(&tmem_full[as]); else umma_commit_pair(&tmem_full[as], (uint16_t)0x3);
        TRACE_ADD(21, clock64() - ti0);   // issuer: whole tile
        TRACE_ADD(22, 1);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp == 3 || (warp == 2 && epi.helpers == 2)) {
    // ------------------------------------------------------------ epilogue TMA helper (staged path only)
    // helpers == 1: lanes 0-3 of warp 3 serve the four quarters (ONE instruction stream: the lanes' waits and their
    // 250-800 cycle bulk-tensor issues serialise); helpers == 2: lanes 0-1 of warps 3 and 2 (warp 2 is idle after the
    // TMEM allocation), two quarters per instruction stream
    if (epi.tma && lane < (epi.helpers == 2 ? 2 : 4)) {
      const int k = (epi.helpers == 2 && warp == 2) ? lane + 2 : lane;   // TMEM lane quarter served by this lane
      const bool f32out = epi.out_f32 != 0;
      const int SUBN = f32out ? 32 : 64;
      const int n_sub = BN / SUBN;
      const bool dual = epi.epilogue == VTB_EPI_SILU_DUAL;
      const bool aux_in = (epi.resid != nullptr) || (epi.epilogue == VTB_EPI_SILU_GRAD);
      const uint32_t qoff = (uint32_t)k * 4096u;
      uint64_t* aux_q = aux_full + k * N_AUX;
      const int my_tiles = (total_tiles - cta + ncl - 1) / ncl;
      const long total_q = (long)my_tiles * n_sub;
      auto q_coords = [&](long q, int& m0, int& n0) {
        const int tl = (int)(q / n_sub), sidx = (int)(q - (long)tl * n_sub);
        const int tile = cta + tl * ncl;
        const int mn = tile / splits;
        m0 = ((mn / n_tiles) * CL + rank) * BM + k * 32;
        n0 = (mn % n_tiles) * BN + sidx * SUBN;
      };
      auto issue_aux = [&](long q) {
        int m0, n0;
        q_coords(q, m0, n0);
        uint64_t* bar = &aux_q[q % N_AUX];
        mbar_expect_tx(bar, EPI_BUF_BYTES / 4);
        tma_load_2d(sAux + (q % N_AUX) * EPI_BUF_BYTES + qoff, &tma_aux, bar, n0, m0);
      };
      tma_prefetch_desc(&tma_out);
      if (dual) tma_prefetch_desc(&tma_out2);
      if (aux_in) {
        tma_prefetch_desc(&tma_aux);
        for (long q = 0; q < N_AUX && q < total_q; ++q) issue_aux(q);
      }
      // output staging ring of this launch: the aux buffers join it when the epilogue neither prefetches through
      // them nor stages a second output (sOut and sAux are contiguous)
      const int nob = (aux_in || dual) ? C::N_OUT : C::N_OUT + N_AUX;
      long qb = 0;       // q % nob, q / nob kept incrementally
      uint32_t qph = 0;
      for (long q = 0; q < total_q; ++q) {
        mbar_wait(&staged_bar[k * MAXOB + qb], qph);
        if (nob == 2 && q >= 1) {
          // two buffers: the math warps want the buffer of sub-tile q-1 back for q+1 — hand it over BEFORE the slow
          // bulk-tensor issue below (its store was issued a whole sub-tile ago and has been read out by now)
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(&free_bar[k * MAXOB + (int)((q - 1) & 1)]);
        }
        int m0, n0;
        q_coords(q, m0, n0);
        if (n0 < epi.N && !GDBG(epi, 4)) {
          const uint32_t so = smem_u32(sOut + qb * EPI_BUF_BYTES) + qoff;
          if (epi.accumulate) {
            asm volatile(
                "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tma_out),
                "r"(so), "r"(n0), "r"(m0)
                : "memory");
          } else {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tma_out),
                         "r"(so), "r"(n0), "r"(m0)
                         : "memory");
            if (dual)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tma_out2),
                           "r"(smem_u32(sAux + (q % N_AUX) * EPI_BUF_BYTES) + qoff), "r"(n0), "r"(m0)
                           : "memory");
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // both warps of the quarter are past their reads of aux buffer q % N_AUX: refill it for sub-tile q + N_AUX
        if (aux_in && q + N_AUX < total_q) issue_aux(q + N_AUX);
        if (nob > 2 && q >= 1) {
          // three or more buffers: sub-tile q-1's buffer is not needed again before q+2, so its read-out may finish
          // behind the issue of store q
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          mbar_arrive(&free_bar[k * MAXOB + (qb == 0 ? nob - 1 : qb - 1)]);
        }
        if (++qb == nob) { qb = 0; qph ^= 1; }
      }
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int ew = warp & 3;  // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    // Bias gradient riding on a weight gradient (A = dy, MN-major): while the tensor cores work through a tile's
    // k-blocks the epilogue warps have nothing to do, so they add up the columns of the A tiles sitting in the ring
    // (A stage = two boxes of 64 k-rows x 64 m, 128B-swizzled: thread = one 16-byte chunk column of 4 k-rows) and hand
    // each stage back themselves (its `empty` barrier counts 1 commit + 8 warps).  Only CTAs on the first n-tile add.
    // The service of the ring is a small state machine (tile, k-block, stage, phase, partial sums) instead of a phase of
    // the tile loop: a tile's remaining k-blocks are served blocking BEFORE its drain, and during the drain every sub-tile
    // step polls the barriers and serves whatever k-blocks of the NEXT tile have become ready — a CTA with several tiles
    // (ViT-B QKV weight gradient: 3 per CTA) no longer stalls its ring for the length of a drain.
    const bool cs_on = A_MN && (epi.a_colsum != nullptr);
    int rd_stage = 0;
    uint32_t rd_phase = 0;
    int r_tile = cta, r_kb = 0, r_kb1 = 0, r_nmod = 0;
    float cacc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cacc[j] = 0.f;
    const int ct = threadIdx.x - 128;                  // 0..255
    const int ccc = ct & 7, cbox = (ct >> 3) & 1, crg = ct >> 4;
    const uint32_t coff = (uint32_t)cbox * (BK * 128) + (uint32_t)crg * 128u + (uint32_t)((ccc ^ (crg & 7)) << 4);
    auto ring_begin = [&]() {
      if (r_tile < total_tiles) {
        const int ks = r_tile % splits;
        r_nmod = (r_tile / splits) % n_tiles;
        r_kb = ks * kb_per_split;
        r_kb1 = min(k_blocks, r_kb + kb_per_split);
      } else {
        r_kb = r_kb1 = 0;
      }
    };
    if (cs_on) ring_begin();
    auto serve_one = [&](bool blocking) -> bool {
      if (r_kb >= r_kb1) return false;
      uint64_t* bar = cs_post ? &mma_done[rd_stage] : &full_bar[rd_stage];
      if (blocking) mbar_wait(bar, rd_phase);
      else if (!__any_sync(0xffffffffu, mbar_try_wait(bar, rd_phase))) return false;  // a completed phase stays completed
      if (cs_post ? (r_kb % n_tiles == r_nmod) : (r_nmod == 0)) {
        const uint32_t a = smem_u32(sA) + (uint32_t)(rd_stage * A_STAGE_BYTES) + coff;
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          const uint4 v = lds_u4(a + r4 * (16 * 128));
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = unpack_bf16(w[q]);
            cacc[2 * q] += f.x; cacc[2 * q + 1] += f.y;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[rd_stage]);
      if (++rd_stage == C::STAGES) { rd_stage = 0; rd_phase ^= 1; }
      ++r_kb;
      return true;
    };
    auto ring_poll = [&]() {
      if (cs_on) while (serve_one(false)) {}
    };
    auto a_colsum_phase = [&](int tile) {   // top of the tile loop, all 8 epilogue warps: finish `tile`'s k-blocks, flush its sums
      if (!cs_on) return;
      while (serve_one(true)) {}
      const int mn = tile / splits;
      const int m_blk = (mn / n_tiles) * CL + rank;
      if (cs_post || (mn % n_tiles) == 0) {  // 16 row groups -> one value per column of the tile (scratch behind the mbarriers) -> global
        const uint32_t s_col = smem_u32(bars) + 1024u;   // float [BM]
        if (ct < BM) sts_f32(s_col + ct * 4, 0.f);
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) {   // lanes l and l ^ 16 hold the same columns (row groups 2 w and 2 w + 1)
          const float v = cacc[j] + __shfl_xor_sync(0xffffffffu, cacc[j], 16);
          if (lane < 16)
            asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(s_col + (uint32_t)(cbox * 64 + ccc * 8 + j) * 4u), "f"(v) : "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (ct < BM && m_blk * BM + ct < epi.M) {
          float sv;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sv) : "r"(s_col + ct * 4));
          atomicAdd(epi.a_colsum + m_blk * BM + ct, sv);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) cacc[j] = 0.f;
      r_tile += ncl;
      ring_begin();
    };
    if (!epi.tma) {
      // direct path (unaligned outputs): per-thread row stores; the two warps of a lane quarter alternate chunks
      const int ehalf = (warp - 4) >> 2;
      for (int tile = cta; tile < total_tiles; tile += ncl) {
        a_colsum_phase(tile);
        const int mn = tile / splits;
        const int n_blk = mn % n_tiles;
        const int m_blk = (mn / n_tiles) * CL + rank;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        const int m = m_blk * BM + ew * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + as * BN;
#pragma unroll 1
        for (int c = ehalf; c < BN / 32; c += 2) {
          const int n0 = n_blk * BN + c * 32;
          if (n0 >= epi.N) break;  // warp-uniform
          ring_poll();
          uint32_t acc[32];
          tmem_ld_32x32(t_row + c * 32, acc);
          tmem_ld_wait();
          epilogue_chunk(epi, acc, m, n0);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CL == 1) mbar_arrive(&tmem_empty[as]); else mbar_arrive_cluster(&tmem_empty[as], 0); }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    } else {
      // staged path: sub-tiles of SUBN columns (one 128-byte row each) -> swizzled smem -> TMA store.
      // The four TMEM lane quarters run INDEPENDENTLY: the two warps of a quarter share a 32-row slab of every staging
      // buffer; a finished slab is handed (mbarrier) to the quarter's lane of the helper warp, which issues the TMA
      // store / aux prefetch (32-row boxes) and hands the buffer back once TMA has read it.  The math warps execute no
      // bulk-tensor instruction, no wait_group and no named barrier.
      const bool f32out = epi.out_f32 != 0;
      const int SUBN = f32out ? 32 : 64;
      const int n_sub = BN / SUBN;
      const bool dual = epi.epilogue == VTB_EPI_SILU_DUAL;
      const bool aux_in = (epi.resid != nullptr) || (epi.epilogue == VTB_EPI_SILU_GRAD);
      uint64_t* aux_q = aux_full + ew * N_AUX;
      uint64_t* staged_q = staged_bar + ew * MAXOB;
      uint64_t* free_q = free_bar + ew * MAXOB;
      const int nob = (aux_in || dual) ? C::N_OUT : C::N_OUT + N_AUX;  // output staging ring (see the helper warp)
      int qb = 0;        // q % nob, (q / nob) & 1 kept incrementally
      uint32_t qph = 0;
      const int ehalf = (warp - 4) >> 2;           // which half of the sub-tile's columns this warp owns
      const int row = ew * 32 + lane;              // row inside the tile == TMEM lane
      const uint32_t swz = (uint32_t)(row & 7);
      const uint32_t cb = (uint32_t)(ehalf * 4);   // first 16-byte chunk of this warp's 64-byte share
      // One sub-tile step.  The TMEM load of the NEXT sub-tile and the bias values of the next sub-tile are issued
      // before this sub-tile's math, so their latencies hide behind it (accumulator registers ping-pong).
      long q = 0;
      auto run = [&](auto cw_tag) {
        constexpr int CW = decltype(cw_tag)::value;   // columns per warp per sub-tile
        constexpr int SUBC = 2 * CW;
        constexpr int NSUB = BN / SUBC;
        for (int tile = cta; tile < total_tiles; tile += ncl) {
          a_colsum_phase(tile);
          const int mn = tile / splits;
          const int n_blk = mn % n_tiles;
          const int m_blk = (mn / n_tiles) * CL + rank;
          const int m = m_blk * BM + row;
#ifdef VTB_GEMM_TRACE
          const long long te0 = clock64();
#endif
          mbar_wait(&tmem_full[as], aphase);
          if (warp == 4 && lane == 0) { TRACE_ADD(16, clock64() - te0); TRACE_ADD(18, 1); }  // epilogue waiting for the accumulator
          tc_fence_after();
          const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + as * BN + ehalf * CW;
          const float rs = (epi.row_scale && m < epi.M) ? __ldg(epi.row_scale + m / epi.rows_per_scale) : 1.f;
          // vector bias path: N a multiple of 4 and a 16-byte aligned bias -> every in-range group of 4 columns is a legal float4
          const bool bias_v = epi.bias && (epi.N & 3) == 0 && (reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0;
          auto bias_at = [&](int sidx) -> float {
            const int col = n_blk * BN + sidx * SUBC + ehalf * CW + lane;
            return (epi.bias && lane < CW && col < epi.N) ? __ldg(epi.bias + col) : 0.f;   // also warms L1 for the vector path
          };
          uint32_t accA[CW], accB[CW];
          tmem_ld_cols<CW>(t_row, accA);
          float b_cur = bias_at(0);
          auto step = [&](const uint32_t (&cur)[CW], uint32_t (&nxt)[CW], int sidx) {
            const int n0 = n_blk * BN + sidx * SUBC;
            const bool live = n0 < epi.N;            // CTA-uniform
            const bool last = (sidx == NSUB - 1);
            const uint32_t ob = smem_u32(sOut) + (uint32_t)(qb * EPI_BUF_BYTES + row * 128);
            const uint32_t ab = smem_u32(sAux) + (uint32_t)((q % N_AUX) * EPI_BUF_BYTES + row * 128);
            const float b_nxt = last ? 0.f : bias_at(sidx + 1);
            ring_poll();   // a_colsum launches: serve the next tile's ready k-blocks (no-op otherwise)
            TRACE_T(0);
            TRACE_T(1);
            TRACE_T(2);
            tmem_ld_wait();
            TRACE_T(3);
            if (!last) {
              if (!GDBG(epi, 16)) tmem_ld_cols<CW>(t_row + (sidx + 1) * SUBC, nxt);
            } else {  // accumulator drained into registers: hand the TMEM stage back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) { if (CL == 1) mbar_arrive(&tmem_empty[as]); else mbar_arrive_cluster(&tmem_empty[as], 0); }
            }
            if (aux_in) mbar_wait(&aux_q[q % N_AUX], (uint32_t)((q / N_AUX) & 1));
            // staging buffer qb was handed to TMA `nob` sub-tiles ago: wait until it has been read out
            mbar_wait(&free_q[qb], qph ^ 1);
            TRACE_T(4);
            const int ncol0 = n0 + ehalf * CW;   // this warp's first column; the vector path needs all CW of them inside N
            const float* bvec = (bias_v && ncol0 + CW <= epi.N) ? epi.bias + ncol0 : nullptr;
            if (live && !GDBG(epi, 8)) staged_row<CW>(epi, cur, b_cur, bvec, rs, ob, ab, cb, swz, dual, CW == 16);
            b_cur = b_nxt;
            TRACE_T(5);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // smem writes -> visible to TMA
            TRACE_T(6);
            __syncwarp();
            if (lane == 0) mbar_arrive(&staged_q[qb]);
            if (++qb == nob) { qb = 0; qph ^= 1; }
            TRACE_T(7);
#ifdef VTB_GEMM_TRACE
            if (lane == 0 && blockIdx.x == 7) {
              const long long tr8 = clock64();
              const long long d[8] = {tr1 - tr0, tr2 - tr1, tr3 - tr2, tr4 - tr3, tr5 - tr4, tr6 - tr5, tr7 - tr6, tr8 - tr7};
              for (int i = 0; i < 8; ++i) atomicAdd(&g_trace[i + (ehalf ? 8 : 0)], (unsigned long long)d[i]);
            }
#endif
            ++q;
          };
#pragma unroll 1
          for (int sidx = 0; sidx < NSUB; sidx += 2) {
            step(accA, accB, sidx);
            if (NSUB > 1) step(accB, accA, sidx + 1);  // NSUB is 1 (BN 64, bf16 out) or even
          }
          if (warp == 4 && lane == 0) TRACE_ADD(17, clock64() - te0);   // epilogue: whole tile
          if (++as == 2) { as = 0; aphase ^= 1; }
        }
      };
      if (f32out) run(std::integral_constant<int, 16>{}); else run(std::integral_constant<int, 32>{});
    }
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // peers may still multicast into / arrive on our smem
  if (warp == 2) {
    tc_fence_after();
    if (CL == 1) tmem_dealloc(tmem_base, C::TMEM_COLS); else tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

int make_tmap(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
              uint32_t box_inner, uint32_t box_outer, bool f32 = false) {
  const uint64_t esz = f32 ? 4 : 2;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    vtb_set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu ld=%llu box=%ux%u f32=%d",
                  (int)r, base, (unsigned long long)inner, (unsigned long long)outer,
                  (unsigned long long)ld_elems, box_inner, box_outer, (int)f32);
    return -3;
  }
  return 0;
}

template <int BN, bool A_MN, bool B_MN, int CL>
int launch(const vtb_gemm_params* p, const EpiParams& epi, int splits, cudaStream_t stream) {
  using C = Cfg<BN, CL>;
  CUtensorMap ta, tb;
  int rc;
  if (!A_MN) rc = make_tmap(&ta, p->A, p->K, p->M, p->lda, BK, BM);
  else       rc = make_tmap(&ta, p->A, p->M, p->K, p->lda, 64, BK);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap(&tb, p->B, p->K, p->N, p->ldb, BK, BN / CL);  // each CTA of a cluster loads BN/CL rows
  else       rc = make_tmap(&tb, p->B, p->N, p->K, p->ldb, 64, BK);
  if (rc) return rc;
  CUtensorMap to = ta, to2 = ta, tx = ta;  // placeholders when the staged epilogue is off
  if (epi.tma) {
    const bool f32 = p->out_f32 != 0;
    const uint32_t subn = f32 ? 32 : 64;
    rc = make_tmap(&to, p->out, p->N, p->M, p->ldo, subn, 32, f32);  // one box per TMEM lane quarter
    if (rc) return rc;
    if (p->out2) {
      rc = make_tmap(&to2, p->out2, p->N, p->M, p->ldo, 64, 32, false);
      if (rc) return rc;
    }
    if (p->resid) {
      rc = make_tmap(&tx, p->resid, p->N, p->M, p->ldr, 32, 32, true);
      if (rc) return rc;
    } else if (p->epilogue == VTB_EPI_SILU_GRAD) {
      rc = make_tmap(&tx, p->aux, p->N, p->M, p->ldaux, 64, 32, false);
      if (rc) return rc;
    }
  }
  const int m_tiles = (p->M + BM - 1) / BM;
  const int n_tiles = (p->N + BN - 1) / BN;
  const int k_blocks = (p->K + BK - 1) / BK;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, CL>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    VTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int total = ((m_tiles + CL - 1) / CL) * n_tiles * splits;  // work items per cluster
  int nclusters = g_num_sms / CL;
  if (total < nclusters) nclusters = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CL);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VTB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, to2, tx, m_tiles, n_tiles, k_blocks, splits, epi));
  VTB_LAUNCH_CHECK();
  return 0;
}

template <int BN, int CL>
int dispatch_major(const vtb_gemm_params* p, const EpiParams& epi, int splits, cudaStream_t s) {
  if (!p->a_mn_major && !p->b_mn_major) return launch<BN, false, false, CL>(p, epi, splits, s);
  if (!p->a_mn_major && p->b_mn_major) return launch<BN, false, true, CL>(p, epi, splits, s);
  if (p->a_mn_major && p->b_mn_major) return launch<BN, true, true, CL>(p, epi, splits, s);
  return launch<BN, true, false, CL>(p, epi, splits, s);
}

}  // namespace

int vtb_gemm_init() {
  if (g_encode) return 0;
  int dev = 0;
  VTB_CUDA(cudaGetDevice(&dev));
  VTB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  VTB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  VTB_CHECK(fn != nullptr && q == cudaDriverEntryPointSuccess, -2,
            "cuTensorMapEncodeTiled not available from the driver");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  if (const char* e = getenv("VTB_GEMM_CLUSTER")) g_use_clusters = atoi(e);
  return 0;
}

int vtb_num_sms() { return g_num_sms; }

void vtb_attn_tc_set(bool on);
void vtb_attn_tc_version_set(int fwd, int bwd);
void vtb_attn_wp_set(bool on);
void vtb_attn_wt_set(bool on);
void vtb_attn_ht_set(bool on);
void vtb_attn_ht_dbg_set(int bits);
void vtb_ln_stream_set(bool on);
void vtb_input_variant_set(int v);

extern "C" int vtb_set_option(const char* name, int32_t value) {
  VTB_CHECK(name != nullptr, -1, "vtb_set_option: null name");
  if (strcmp(name, "gemm_cluster") == 0) { g_use_clusters = value; return 0; }
  if (strcmp(name, "gemm_colsum_pair") == 0) { g_colsum_pair = value; return 0; }
  if (strcmp(name, "gemm_bn_waste_pct") == 0) { g_bn_waste_pct = value; return 0; }
  if (strcmp(name, "gemm_helpers") == 0) { g_helpers = value == 2 ? 2 : 1; return 0; }
#ifdef VTB_GEMM_DBG
  if (strcmp(name, "gemm_dbg") == 0) { g_gemm_dbg = value; return 0; }
#endif
  if (strcmp(name, "attn_tc") == 0) { vtb_attn_tc_set(value != 0); return 0; }
  if (strcmp(name, "attn_tc_fwd_version") == 0) { vtb_attn_tc_version_set(value, 0); return 0; }
  if (strcmp(name, "attn_tc_bwd_version") == 0) { vtb_attn_tc_version_set(0, value); return 0; }
  if (strcmp(name, "input_variant") == 0) { vtb_input_variant_set(value); return 0; }
  if (strcmp(name, "attn_wp") == 0) { vtb_attn_wp_set(value != 0); return 0; }
  if (strcmp(name, "attn_wt") == 0) { vtb_attn_wt_set(value != 0); return 0; }
  if (strcmp(name, "attn_ht") == 0) { vtb_attn_ht_set(value != 0); return 0; }
  if (strcmp(name, "attn_ht_dbg") == 0) { vtb_attn_ht_dbg_set(value); return 0; }
  if (strcmp(name, "ln_stream") == 0) { vtb_ln_stream_set(value != 0); return 0; }
  vtb_set_error("vtb_set_option: unknown option '%s'", name);
  return -1;
}

extern "C" int vtb_gemm_bf16(const vtb_gemm_params* p, vtb_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VTB_CHECK(p != nullptr, -1, "vtb_gemm_bf16: null params");
  VTB_CHECK(g_encode != nullptr, -2, "vtb_gemm_bf16: call vtb_init() first");
  VTB_CHECK(p->M > 0 && p->N > 0 && p->K > 0, -1, "vtb_gemm_bf16: bad shape M=%d N=%d K=%d", p->M,
            p->N, p->K);
  VTB_CHECK(p->A && p->B && p->out, -1, "vtb_gemm_bf16: null operand");
  VTB_CHECK(((uintptr_t)p->A & 15) == 0 && ((uintptr_t)p->B & 15) == 0, -1,
            "vtb_gemm_bf16: operands must be 16-byte aligned");
  VTB_CHECK(p->lda % 8 == 0 && p->ldb % 8 == 0, -1,
            "vtb_gemm_bf16: lda/ldb must be multiples of 8 (TMA 16-byte strides), got %d %d",
            p->lda, p->ldb);
  VTB_CHECK(!p->accumulate || p->out_f32, -1, "vtb_gemm_bf16: accumulate needs an f32 output");
  VTB_CHECK(p->splits <= 1 || p->accumulate, -1, "vtb_gemm_bf16: split-K needs accumulate=1");
  VTB_CHECK(p->epilogue != VTB_EPI_SILU_DUAL || (p->out2 && !p->out_f32), -1,
            "vtb_gemm_bf16: SILU_DUAL needs bf16 out and out2");
  VTB_CHECK(p->epilogue != VTB_EPI_SILU_GRAD || p->aux, -1, "vtb_gemm_bf16: SILU_GRAD needs aux");
  VTB_CHECK(!(p->bias && p->splits > 1), -1, "vtb_gemm_bf16: bias cannot be combined with split-K");
  VTB_CHECK(!(p->resid && p->epilogue != VTB_EPI_NONE), -1, "vtb_gemm_bf16: resid only with EPI_NONE");
  VTB_CHECK(!p->resid || p->out_f32, -1, "vtb_gemm_bf16: resid needs an f32 output");
  VTB_CHECK(p->epilogue != VTB_EPI_SILU_GRAD || !p->out_f32, -1, "vtb_gemm_bf16: SILU_GRAD writes bf16");
  VTB_CHECK(!p->row_scale || p->rows_per_scale > 0, -1, "vtb_gemm_bf16: rows_per_scale");
  VTB_CHECK(!p->a_colsum || (p->a_mn_major && p->epilogue == VTB_EPI_NONE && !p->resid && !p->out2), -1,
            "vtb_gemm_bf16: a_colsum needs an MN-major A operand and a plain epilogue");

  EpiParams e;
  e.M = p->M; e.N = p->N;
  e.out = p->out; e.ldo = p->ldo; e.out_f32 = p->out_f32;
  e.out2 = reinterpret_cast<bf16*>(p->out2);
  e.bias = p->bias;
  e.resid = p->resid; e.ldr = p->ldr;
  e.row_scale = p->row_scale; e.rows_per_scale = p->rows_per_scale;
  e.aux = reinterpret_cast<const bf16*>(p->aux); e.ldaux = p->ldaux;
  e.epilogue = p->epilogue;
  e.accumulate = p->accumulate;
  e.alpha = p->alpha;
  e.a_colsum = p->a_colsum;
  e.helpers = g_helpers;
#ifdef VTB_GEMM_DBG
  e.dbg = g_gemm_dbg;
#else
  e.dbg = 0;
#endif
  {  // 16-byte vector epilogue only when every touched row start is 16-byte aligned; scalar path otherwise
    const int oalign = p->out_f32 ? 4 : 8;
    bool v = (p->ldo % oalign == 0) && (((uintptr_t)p->out & 15) == 0);
    if (p->out2) v = v && (((uintptr_t)p->out2 & 15) == 0);
    if (p->resid) v = v && (p->ldr % 4 == 0) && (((uintptr_t)p->resid & 15) == 0);
    if (p->aux) v = v && (p->ldaux % 8 == 0) && (((uintptr_t)p->aux & 15) == 0);
    e.vec = v ? 1 : 0;
    // staged TMA epilogue needs the same alignment plus a 16-byte aligned bias (float4 loads)
    e.tma = (v && (!p->bias || ((uintptr_t)p->bias & 15) == 0)) ? 1 : 0;
  }

  // Tile-N choice: widest tile that does not waste more than ~25% of the MMA columns.
  int bn = 256;
  if (p->N <= 64) bn = 64;
  else if (p->N <= 128) bn = 128;
  else {
    auto waste = [&](int b) { return (double)(((p->N + b - 1) / b) * b) / p->N; };
    const double tol = g_bn_waste_pct / 100.0;
    if (waste(256) > tol * waste(128)) bn = 128;
    if (bn == 128 && waste(128) > tol * waste(64)) bn = 64;
  }
  const int m_tiles = (p->M + BM - 1) / BM;
  const int n_tiles = (p->N + bn - 1) / bn;
  const int k_blocks = (p->K + BK - 1) / BK;
  // CTA pairs (256 x bn tiles, cta_group::2) whenever there are at least two row tiles and the pairs can fill the
  // machine; an MN-major B tile of 64 columns is one TMA box and cannot be halved
  bool pair = g_use_clusters && m_tiles >= 2 && !(p->b_mn_major && bn < 128);
  if (p->a_colsum && !g_colsum_pair) pair = false;  // A/B switch: 1-CTA tiles, the readers wait on their own CTA's `full` barrier
  const long units = (long)(pair ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;  // work items before split-K
  const int slots = pair ? g_num_sms / 2 : g_num_sms;                        // concurrently resident work items
  int splits = p->splits;
  if (splits <= 0) {
    splits = 1;
    if (p->accumulate && !p->bias) {
      // split-K so that the work items fill whole waves of the persistent grid: among the split factors that
      // leave >= 8 k-blocks per slice pick the one with the best wave efficiency (fewest slices on ties)
      int maxs = k_blocks / 8;
      if (maxs < 1) maxs = 1;
      if (maxs > 64) maxs = 64;
      double best = -1.0;
      for (int sp = 1; sp <= maxs; ++sp) {
        const long t = units * sp;
        const long waves = (t + slots - 1) / slots;
        double eff = (double)t / (double)(waves * slots);
        if (t < slots) eff = (double)t / slots;
        eff -= 0.002 * sp;  // reduce-add traffic grows with the split factor
        if (eff > best + 1e-9) { best = eff; splits = sp; }
      }
    }
  }
  if (splits > k_blocks) splits = k_blocks;
  {  // no empty split
    int per = (k_blocks + splits - 1) / splits;
    splits = (k_blocks + per - 1) / per;
  }
  // Round 1 dropped the pairs when pairs x splits did not fill the pair slots ("independent 128-row tiles spread wider").  They do not:
  // the split factor is kept, so 1-CTA tiles occupy exactly the SMs the pairs would, with a third more operand traffic per SM
  // and a shallower ring — the ViT-B fc1 / fc2 weight gradients (36 pair tiles x 2 splits = 72 of 74 slots) ran 16-20 % slower for it
  // (profiles/r02_cabi_gemm_wgrad_splits_pairs.log; whole steps: ViT-B 34.13 -> 33.44 ms, Swin-S 35.25 -> 34.88, r02_ab_pairs_rule.log).
  // gemm_cluster = 3 keeps the old rule for A/B.
  if (pair && g_use_clusters == 3 && units * splits < slots) pair = false;
  switch (bn) {
    case 256: return pair ? dispatch_major<256, 2>(p, e, splits, stream) : dispatch_major<256, 1>(p, e, splits, stream);
    case 128: return pair ? dispatch_major<128, 2>(p, e, splits, stream) : dispatch_major<128, 1>(p, e, splits, stream);
    default:  return pair ? dispatch_major<64, 2>(p, e, splits, stream) : dispatch_major<64, 1>(p, e, splits, stream);
  }
}

#ifdef VTB_GEMM_TRACE
extern "C" int vtb_debug_gemm_trace(unsigned long long* out, int reset) {
  if (out) VTB_CUDA(cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 32));
  if (reset) { unsigned long long z[32] = {0}; VTB_CUDA(cudaMemcpyToSymbol(g_trace, z, sizeof(z))); }
  return 0;
}
#endif
