// tcgen05 / TMEM shifted-window attention (Swin / Twins-LSA: W^2 <= 64 tokens per window, dh = 32), forward and
// backward.  swin_transformer.py:103-160, twins.py:109-152.
//
// A window is a 49 x 49 x 32 problem — far below one UMMA tile — so TWO windows share every 128-row tile and the
// two problems are kept apart by stacking them along the CONTRACTION dimension:
//     A rows (w, t) hold window w's 32 features in columns [32 w, 32 w + 32) of a 64-wide K axis, zeros elsewhere
//     B rows  t     hold [window 0 features | window 1 features]
//   =>  D[(w, t), u] = A_w[t] . B_w[u]: the two diagonal blocks of the 128 x 128 product in 64 TMEM columns, no
//       off-diagonal work and no garbage.  The same tiles serve the other products as MN-major operands (the padded
//       tile as B kills the cross-window terms), so nothing is transposed or copied.
// One persistent CTA per SM, fixed head per CTA (its bias tile lives in shared memory for the whole kernel):
//   warps 0-3  : loaders, cp.async 16-byte gathers straight out of the fused qkv buffer (window partition and
//                cyclic shift are address arithmetic); the copies themselves complete the stage barriers
//   warps 4-11 : math, one tile row (TMEM lane) per thread, scores read straight from TMEM
//   warp 12    : tcgen05.mma issuer (one thread), TMEM allocation
// Registers are re-balanced with setmaxnreg: loader / issuer warpgroups give theirs to the two math warpgroups.
// Relative-position bias and shift mask are an fp32 tile (pre-multiplied by log2 e, -inf on key padding) and one
// 64-bit word per row; the bias gradient is accumulated in registers over every window a thread sees and reduced
// once per CTA.
#include "attn_win_common.cuh"
#include "../../include/vtb200.h"
#include <math.h>

namespace {

struct WtGeom {
  int heads, nq, Hs, Ws, window, shift, nwx, nw, groups;
  float inv_nw, inv_nwx;
};

// token of in-window position t of group grp, -1 for padding
__device__ __forceinline__ int wt_token(const WtGeom& g, int grp, int t) {
  if (t >= g.nq || grp >= g.groups) return -1;
  const int b = grp / g.nw, wi = grp - b * g.nw;
  const int wy = wi / g.nwx, wx = wi - wy * g.nwx;
  const int ty = t / g.window, tx = t - ty * g.window;
  int y = wy * g.window + ty + g.shift, x = wx * g.window + tx + g.shift;
  if (y >= g.Hs) y -= g.Hs;
  if (x >= g.Ws) x -= g.Ws;
  return (b * g.Hs + y) * g.Ws + x;
}
// group part / row part of wt_token (the loaders resolve several rows of the same window)
struct WtOrigin { int y0, x0, img, wi; };
__device__ __forceinline__ WtOrigin wt_origin(const WtGeom& g, int grp) {
  WtOrigin o;
  if (grp >= g.groups) { o.img = -1; o.y0 = o.x0 = o.wi = 0; return o; }
  const int b = wt_div(grp, g.nw, g.inv_nw), wi = grp - b * g.nw;
  const int wy = wt_div(wi, g.nwx, g.inv_nwx), wx = wi - wy * g.nwx;
  o.img = b; o.wi = wi; o.y0 = wy * g.window + g.shift; o.x0 = wx * g.window + g.shift;
  return o;
}
__device__ __forceinline__ int wt_row_token(const WtGeom& g, const WtOrigin& o, int t) {
  if (t >= g.nq || o.img < 0) return -1;
  const int ty = t / g.window, tx = t - ty * g.window;
  int y = o.y0 + ty, x = o.x0 + tx;
  if (y >= g.Hs) y -= g.Hs;
  if (x >= g.Ws) x -= g.Ws;
  return (o.img * g.Hs + y) * g.Ws + x;
}
// per-thread constants of the four token rows (it * 16 + r4) a loader thread copies in every window
struct WtRows {
  int ty[4], tx[4];
  bool valid[4];
};
__device__ __forceinline__ WtRows wt_rows(const WtGeom& g, int r4) {
  WtRows r;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 16 + r4;
    r.valid[it] = row < g.nq;
    r.ty[it] = row / g.window;
    r.tx[it] = row - r.ty[it] * g.window;
  }
  return r;
}
__device__ __forceinline__ int wt_tok(const WtGeom& g, const WtOrigin& o, int ty, int tx) {
  int y = o.y0 + ty, x = o.x0 + tx;
  if (y >= g.Hs) y -= g.Hs;
  if (x >= g.Ws) x -= g.Ws;
  return (o.img * g.Hs + y) * g.Ws + x;
}

// bias tile shared by forward (rows = queries, cols = keys) and backward (rows = keys, cols = queries):
// tile[row][col] (fp32, 16-byte chunks XOR-swizzled by row & 7) = rel_bias[pos[i, j], h] * log2(e); -inf where the KEY
// index is padding (>= nq); 0 where only the query index is padding.
__device__ __forceinline__ void wt_build_bias(const vtb_attn_params& p, int nq, int heads, int h, float* tab,
                                              uint8_t* tile, bool rows_are_keys) {
  if (p.rel_bias)
    for (int t = threadIdx.x; t < p.n_pos; t += blockDim.x) tab[t] = __ldg(p.rel_bias + (long)t * heads + h) * WT_L2E;
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
    const int row = e >> 6, col = e & 63;
    const int i = rows_are_keys ? col : row, j = rows_are_keys ? row : col;
    float v = 0.f;
    if (j >= nq) v = -INFINITY;
    else if (i < nq && p.rel_bias) v = tab[__ldg(p.pos + i * nq + j)];
    *reinterpret_cast<float*>(tile + row * 256 + ((((col >> 2) ^ (row & 7))) << 4) + (col & 3) * 4) = v;
  }
  __syncthreads();
}

// =====================================================================================================
// forward
//   S[(w,i), j] = Qpad . Kcat^T (M 128, N 64, K 64)   -> softmax per row (one thread per row) -> P (bf16) back
//   into TMEM over S -> O[(w,i), (w',d)] = P . Vcat (A from TMEM, B MN-major); columns w' = w are the output.
//   6 smem stages (a stage's mbarrier is completed by the cp.async engine itself: the loaders never block on a load)
//   / 4 TMEM buffers; the two math groups take alternate tiles.
// =====================================================================================================
constexpr int F_STAGES = 6;
constexpr int F_STAGE_BYTES = 16384 + 8192 + 8192;
constexpr int F_SIDE_BYTES = 128 * 8;  // mask bits
constexpr int F_SMEM = F_STAGES * F_STAGE_BYTES + 16384 + F_STAGES * F_SIDE_BYTES + 256 + 1024;

__global__ void __launch_bounds__(WT_THREADS, 1)
attn_wt_fwd_kernel(vtb_attn_params p, WtGeom g, int ntiles, int nchunks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sStage = smem;
  uint8_t* sBias = smem + F_STAGES * F_STAGE_BYTES;
  uint8_t* sSide = sBias + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSide + F_STAGES * F_SIDE_BYTES);
  uint64_t* full = bars;                 // [6] 128 cp.async arrivals (loader threads) -> issuer / math
  uint64_t* empty = bars + 6;            // [6] PV retired -> loaders
  uint64_t* s_full = bars + 12;          // [4] S complete
  uint64_t* p_full = bars + 16;          // [4] P written (4 warps)
  uint64_t* o_full = bars + 20;          // [4] O complete
  uint64_t* o_free = bars + 24;          // [4] O drained (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % g.heads;
  const int chunk = blockIdx.x / g.heads;
  const int my_tiles = (ntiles - chunk + nchunks - 1) / nchunks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < F_STAGES; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  // zero the operand ring once (the zero halves of Qpad never change; padding rows are re-zeroed by the loaders)
  for (int e = threadIdx.x; e < F_STAGES * F_STAGE_BYTES / 16; e += WT_THREADS)
    reinterpret_cast<uint4*>(sStage)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  wt_build_bias(p, g.nq, g.heads, h, reinterpret_cast<float*>(sStage), sBias, false);
  // the table scratch lived in stage 0: clear it again
  for (int e = threadIdx.x; e < 2048 / 16; e += WT_THREADS) reinterpret_cast<uint4*>(sStage)[e] = make_uint4(0, 0, 0, 0);
  wt_proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders (128 threads)
    // thread = (window w, chunk cc of rows r4 + 16 it): four lanes cover the four 16-byte chunks of one token row
    // (8 rows x 64 B per warp instruction); thread tt also fetches the mask word of tile row tt.
    wt_reg_dec<80>();
    const int tt = threadIdx.x, w = tt >> 6, tl = tt & 63, cc = tl & 3, r4 = tl >> 2;
    const WtRows rows = wt_rows(g, r4);
    const uint32_t ch = (uint32_t)(((w * 4 + cc) ^ (r4 & 7)) << 4);
    const uint32_t st0 = smem_u32(sStage) + (uint32_t)r4 * 128u + ch;
    const char* qb = reinterpret_cast<const char*>(p.q) + (h * 32 + cc * 8) * 2;
    const char* kb = reinterpret_cast<const char*>(p.k) + (h * 32 + cc * 8) * 2;
    const char* vb = reinterpret_cast<const char*>(p.v) + (h * 32 + cc * 8) * 2;
    const long ldq2 = (long)p.ldq * 2, ldk2 = (long)p.ldk * 2, ldv2 = (long)p.ldv * 2;
    for (int n = 0; n < my_tiles; ++n) {
      const int tile = chunk + n * nchunks;
      const int stage = n % F_STAGES;
      mbar_wait(&empty[stage], ((n / F_STAGES) & 1) ^ 1);
      const uint32_t st = st0 + stage * F_STAGE_BYTES;
      unsigned long long* s_mb = reinterpret_cast<unsigned long long*>(sSide + stage * F_SIDE_BYTES);
      const int grp = tile * 2 + w;
      const WtOrigin org = wt_origin(g, grp);
      const bool ok = org.img >= 0;
      if (p.mask_bits && ok && tl < g.nq)
        cp_async8(smem_u32(&s_mb[tt]), reinterpret_cast<const unsigned long long*>(p.mask_bits) +
                                           (long)(p.n_mask == g.nw ? org.wi : grp % p.n_mask) * 128 + tl);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (rows.valid[it]) {  // rows >= nq stay zero from the prologue
          const long gr = ok ? (long)wt_tok(g, org, rows.ty[it], rows.tx[it]) : 0;
          const uint32_t ro = (uint32_t)(it * 16) * 128u;
          cp_async16(st + (uint32_t)(w * 64) * 128u + ro, qb + gr * ldq2, ok);
          cp_async16(st + 16384 + ro, kb + gr * ldk2, ok);
          cp_async16(st + 24576 + ro, vb + gr * ldv2, ok);
        }
      }
      cp_async_arrive_noinc(&full[stage]);
    }
    cp_async_wait<0>();  // nothing may still be in flight towards this CTA's shared memory at exit
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ UMMA issuer (warp 12, one thread)
    wt_reg_dec<40>();
    if (warp == 12 && lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      // event-driven: whichever of "next score tile" / "next P.V" has its inputs ready is issued, so a late load
      // never delays a P.V product the math warps are waiting for
      int ns = 0, np = 0;
      WtWatchdog dog;
      dog.reset();
      while (np < my_tiles) {
        bool did = false;
        if (ns < my_tiles && ns < np + 4) {
          const int b = ns & 3, stage = ns % F_STAGES;
          if (mbar_test(&full[stage], (ns / F_STAGES) & 1) && mbar_test(&o_free[b], ((ns >> 2) & 1) ^ 1)) {
            wt_proxy_fence();  // cp.async (generic proxy) writes -> visible to the tensor core's async-proxy reads
            tc_fence_after();
            const uint32_t qa = smem_u32(sStage + stage * F_STAGE_BYTES), ka = qa + 16384;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + b * 128, umma_desc_sw128(qa + k * 32, 0, 1024),
                        umma_desc_sw128(ka + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&s_full[b]);
            ++ns;
            did = true;
          }
        }
        if (np < ns) {
          const int b = np & 3, stage = np % F_STAGES;
          if (mbar_test(&p_full[b], (np >> 2) & 1)) {
            tc_fence_after();
            const uint32_t va = smem_u32(sStage + stage * F_STAGE_BYTES) + 24576;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              wt_umma_ts(tmem_base + b * 128 + 64, tmem_base + b * 128 + k * 8,
                         umma_desc_sw128(va + k * 2048, 0, 1024), idesc_o, k > 0 ? 1u : 0u);
            umma_commit(&o_full[b]);
            umma_commit(&empty[stage]);
            ++np;
            did = true;
          }
        }
        if (did) dog.reset(); else dog.idle();
      }
    }
  } else {
    // ------------------------------------------------------------------ math: thread = query row (w, i)
    wt_reg_inc<192>();
    const int grpi = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane, w = r >> 6, i = r & 63;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * WT_L2E;
    // explicit shared-space addresses: pointers carved out of the dynamic block are generic to the compiler and their
    // loads would go through the L1TEX path (LD.E) instead of LDS
    const uint32_t brow = smem_u32(sBias) + (uint32_t)i * 256u;
    const int sx = i & 7;
    bf16* O = reinterpret_cast<bf16*>(p.o);
    for (int n = grpi; n < my_tiles; n += 2) {
      const int tile = chunk + n * nchunks;
      const int b = n & 3, stage = n % F_STAGES;
      const int tok = wt_row_token(g, wt_origin(g, tile * 2 + w), i);
      mbar_wait(&full[stage], (n / F_STAGES) & 1);
      const unsigned long long mb =
          (p.mask_bits && tok >= 0) ? wt_lds_u64(smem_u32(sSide) + (uint32_t)(stage * F_SIDE_BYTES + r * 8)) : 0ull;
      mbar_wait(&s_full[b], (n >> 2) & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(t_lane + b * 128, s0);
      tmem_ld_32x32(t_lane + b * 128 + 32, s1);
      tmem_ld_wait();
      float x[64];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float4 bb = lds_f4(brow + (uint32_t)((c ^ sx) << 4));
        const uint32_t* src = (c < 8) ? &s0[c * 4] : &s1[(c - 8) * 4];
        x[c * 4 + 0] = fmaf(__uint_as_float(src[0]), sl2, bb.x);
        x[c * 4 + 1] = fmaf(__uint_as_float(src[1]), sl2, bb.y);
        x[c * 4 + 2] = fmaf(__uint_as_float(src[2]), sl2, bb.z);
        x[c * 4 + 3] = fmaf(__uint_as_float(src[3]), sl2, bb.w);
      }
      if (__any_sync(0xffffffffu, mb != 0ull)) {
        const uint32_t lo = (uint32_t)mb, hi = (uint32_t)(mb >> 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (lo & (1u << j)) x[j] = -INFINITY;
          if (hi & (1u << j)) x[32 + j] = -INFINITY;
        }
      }
#pragma unroll
      for (int j = 0; j < 64; ++j) mx = fmaxf(mx, x[j]);
      const float m_use = (mx == -INFINITY) ? 0.f : mx;
      float sum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        const float p0 = wt_ex2(x[j] - m_use), p1 = wt_ex2(x[j + 1] - m_use);
        sum += p0 + p1;
        pk[j >> 1] = pack_bf16(p0, p1);
      }
      wt_tmem_st16(t_lane + b * 128, pk);
      wt_tmem_st16(t_lane + b * 128 + 16, pk + 16);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      wt_warp_arrive(&p_full[b], lane);
      // epilogue: O / l -> bf16 -> global; lse = ln 2 * (max + log2(sum))
      mbar_wait(&o_full[b], (n >> 2) & 1);
      tc_fence_after();
      uint32_t a[32];
      tmem_ld_32x32(t_lane + b * 128 + 64 + w * 32, a);
      tmem_ld_wait();
      tc_fence_before();
      wt_warp_arrive(&o_free[b], lane);
      if (tok >= 0) {
        wt_store_row32(O + (long)tok * p.ldo + h * 32, a, 1.f / sum);
        if (p.lse) p.lse[((long)(tile * 2 + w) * g.heads + h) * g.nq + i] = (mx + log2f(sum)) * WT_LN2;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================
// backward (keys on the tile rows, as in the global tcgen05 backward):
//   S^T[(w,j), i] = Kpad . Qcat^T      dP^T[(w,j), i] = Vpad . dOcat^T             (M 128, N 64, K 64)
//   P^T = exp2(S^T sl2 + bias - lse2[i]),  dS^T = P^T (dP^T - delta[i])   -> bf16 tiles in shared memory
//   dV[(w,j), (w',d)] = P^T . dOcat     dK = dS^T . Qcat    (A K-major, B MN-major; columns w' = w are the result)
//   dQ[i, (w,d)]      = dS . Kpad       (A = the dS^T tile read MN-major, B MN-major; the padded tile separates
//                                        the windows; rows 64-127 of the accumulator are a don't-care second M atom)
// TMEM: S^T 2 x 64 | dP^T 2 x 64 | dV 64 | dK 64 | dQ 64.   3 operand stages, P^T / dS^T single-buffered.
// =====================================================================================================
constexpr int B_STAGES = 3;
constexpr int B_OFF_V = 16384, B_OFF_Q = 32768, B_OFF_DO = 40960, B_OFF_O = 49152;
constexpr int B_STAGE_BYTES = 57344;  // Kpad 16K | Vpad 16K | Qcat 8K | dOcat 8K | O rows 8K (plain, for delta)
constexpr int B_SIDE_BYTES = 128 * 4 * 3 + 128 * 8;  // tok, lse2, delta, mask bits
constexpr int B_SMEM = B_STAGES * B_STAGE_BYTES + 2 * 16384 + 16384 + B_STAGES * B_SIDE_BYTES + 256 + 1024;
constexpr uint32_t C_ST = 0, C_DP = 128, C_DV = 256, C_DK = 320, C_DQ = 384;

__global__ void __launch_bounds__(WT_THREADS, 1)
attn_wt_bwd_kernel(vtb_attn_params p, WtGeom g, int ntiles, int nchunks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sStage = smem;
  uint8_t* sdS = smem + B_STAGES * B_STAGE_BYTES;   // [128 key rows][64 queries] bf16, 128B-swizzled
  uint8_t* sP = sdS + 16384;                        // directly behind dS^T: the dQ product's second M atom lands here
  uint8_t* sBias = sP + 16384;
  uint8_t* sSide = sBias + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSide + B_STAGES * B_SIDE_BYTES);
  uint64_t* full = bars;            // [3] count 4 (loader warps, after delta / lse of the landed tile are in place)
  uint64_t* land = bars + 16;       // [3] 128 cp.async arrivals: the tile's rows have landed
  uint64_t* empty = bars + 3;       // [3] gradient MMAs of the tile retired
  uint64_t* s_full = bars + 6;      // [2] S^T / dP^T complete
  uint64_t* s_free = bars + 8;      // [2] read out of TMEM (8 warps)
  uint64_t* pds_full = bars + 10;   // P^T / dS^T tiles written (8 warps)
  uint64_t* pds_free = bars + 11;   // ... and consumed by the gradient MMAs
  uint64_t* g_full = bars + 12;     // dV / dK / dQ complete
  uint64_t* g_free = bars + 13;     // ... and drained (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % g.heads;
  const int chunk = blockIdx.x / g.heads;
  const int my_tiles = (ntiles - chunk + nchunks - 1) / nchunks;
  const bool has_tab = p.rel_bias != nullptr && p.drel_bias != nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&full[i], 4); mbar_init(&empty[i], 1); mbar_init(&land[i], 128); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 8); }
    mbar_init(pds_full, 8); mbar_init(pds_free, 1); mbar_init(g_full, 1); mbar_init(g_free, 8);
    mbar_fence_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  for (int e = threadIdx.x; e < (B_STAGES * B_STAGE_BYTES + 2 * 16384) / 16; e += WT_THREADS)
    reinterpret_cast<uint4*>(sStage)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  wt_build_bias(p, g.nq, g.heads, h, reinterpret_cast<float*>(sP), sBias, true);
  for (int e = threadIdx.x; e < 2048 / 16; e += WT_THREADS) reinterpret_cast<uint4*>(sP)[e] = make_uint4(0, 0, 0, 0);
  wt_proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders (128 threads)
    // data: thread = (window w, chunk cc of rows r4 + 16 it), four lanes per token row; side info (token, mask word,
    // lse, delta): thread tt owns tile row tt.  Everything a tile needs arrives by cp.async (no blocking loads).
    wt_reg_dec<80>();
    const int tt = threadIdx.x, w = tt >> 6, tl = tt & 63, cc = tl & 3, r4 = tl >> 2;
    // a tile's rows have landed: delta = dO . O and lse * log2(e) from shared memory, then release
    auto finish = [&](int stage) {
      uint8_t* st = sStage + stage * B_STAGE_BYTES;
      float* side = reinterpret_cast<float*>(sSide + stage * B_SIDE_BYTES);
      const int* s_tok = reinterpret_cast<const int*>(side);
      float acc = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint4 ra = *reinterpret_cast<const uint4*>(st + B_OFF_DO + sw128(tl, w * 4 + c4));
        const uint4 rb = *reinterpret_cast<const uint4*>(st + B_OFF_O + tt * 64 + c4 * 16);
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fa = unpack_bf16(wa[q]), fb = unpack_bf16(wb[q]);
          acc += fa.x * fb.x + fa.y * fb.y;
        }
      }
      side[128 + tt] = (s_tok[tt] >= 0) ? side[128 + tt] * WT_L2E : INFINITY;
      side[256 + tt] = acc;
      wt_proxy_fence();
      wt_warp_arrive(&full[stage], lane);
    };
    const WtRows rows = wt_rows(g, r4);
    const uint32_t ch = (uint32_t)(((w * 4 + cc) ^ (r4 & 7)) << 4);
    const uint32_t st0 = smem_u32(sStage) + (uint32_t)r4 * 128u + ch;
    const uint32_t so0 = smem_u32(sStage) + B_OFF_O + (uint32_t)(w * 64 + r4) * 64u + cc * 16;
    const int colb = (h * 32 + cc * 8) * 2;
    const char* qb = reinterpret_cast<const char*>(p.q) + colb;
    const char* kb = reinterpret_cast<const char*>(p.k) + colb;
    const char* vb = reinterpret_cast<const char*>(p.v) + colb;
    const char* dob = reinterpret_cast<const char*>(p.dout) + colb;
    const char* ob = reinterpret_cast<const char*>(p.o) + colb;
    const long ldq2 = (long)p.ldq * 2, ldk2 = (long)p.ldk * 2, ldv2 = (long)p.ldv * 2, lddo2 = (long)p.lddo * 2,
               ldo2 = (long)p.ldo * 2;
    auto issue = [&](int n) {
      const int tile = chunk + n * nchunks;
      const int stage = n % 3;
      const uint32_t st = st0 + stage * B_STAGE_BYTES, so = so0 + stage * B_STAGE_BYTES;
      float* side = reinterpret_cast<float*>(sSide + stage * B_SIDE_BYTES);
      int* s_tok = reinterpret_cast<int*>(side);
      unsigned long long* s_mb = reinterpret_cast<unsigned long long*>(sSide + stage * B_SIDE_BYTES + 1536);
      const int grp = tile * 2 + w;
      const WtOrigin org = wt_origin(g, grp);
      const bool ok = org.img >= 0;
      const int tok_side = wt_row_token(g, org, tl);
      s_tok[tt] = tok_side;
      if (p.mask_bits && tok_side >= 0)
        cp_async8(smem_u32(&s_mb[tt]), reinterpret_cast<const unsigned long long*>(p.mask_bits) +
                                           (long)(p.n_mask == g.nw ? org.wi : grp % p.n_mask) * 128 + 64 + tl);
      else
        s_mb[tt] = 0ull;
      if (tok_side >= 0) cp_async4(smem_u32(&side[128 + tt]), p.lse + ((long)grp * g.heads + h) * g.nq + tl);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (rows.valid[it]) {  // rows >= nq stay zero from the prologue
          const long gr = ok ? (long)wt_tok(g, org, rows.ty[it], rows.tx[it]) : 0;
          const uint32_t ro = (uint32_t)(it * 16) * 128u;
          cp_async16(st + (uint32_t)(w * 64) * 128u + ro, kb + gr * ldk2, ok);
          cp_async16(st + B_OFF_V + (uint32_t)(w * 64) * 128u + ro, vb + gr * ldv2, ok);
          cp_async16(st + B_OFF_Q + ro, qb + gr * ldq2, ok);
          cp_async16(st + B_OFF_DO + ro, dob + gr * lddo2, ok);
          cp_async16(so + (uint32_t)(it * 16) * 64u, ob + gr * ldo2, ok);
        }
      }
      cp_async_arrive_noinc(&land[stage]);
    };
    // event-driven (warp-uniform votes): fetch the next tile as soon as its stage is free, publish a tile as soon
    // as the copy engine reports it landed — the loaders never block on a load
    int ni = 0, nf = 0;
    WtWatchdog dog;
    dog.reset();
    while (nf < my_tiles) {
      bool did = false;
      if (ni < my_tiles && __all_sync(0xffffffffu, mbar_test(&empty[ni % 3], ((ni / 3) & 1) ^ 1))) {
        issue(ni);
        ++ni;
        did = true;
      }
      if (nf < ni && __all_sync(0xffffffffu, mbar_test(&land[nf % 3], (nf / 3) & 1))) {
        finish(nf % 3);
        ++nf;
        did = true;
      }
      if (did) dog.reset(); else dog.idle();
    }
    cp_async_wait<0>();
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ UMMA issuer (warp 12, one thread)
    wt_reg_dec<40>();
    if (warp == 12 && lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, 64, 0, 1);
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, 1, 1);
      const uint32_t pa = smem_u32(sP), sa = smem_u32(sdS);
      int ns = 0, ng = 0;  // next score tile / next gradient tile (event-driven, see the forward issuer)
      WtWatchdog dog;
      dog.reset();
      while (ng < my_tiles) {
        bool did = false;
        if (ns < my_tiles && ns < ng + 2) {
          const int stage = ns % 3, b = ns & 1;
          if (mbar_test(&full[stage], (ns / 3) & 1) && mbar_test(&s_free[b], ((ns >> 1) & 1) ^ 1)) {
            wt_proxy_fence();  // cp.async (generic proxy) writes -> visible to the tensor core's async-proxy reads
            tc_fence_after();
            const uint32_t ka = smem_u32(sStage + stage * B_STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + C_ST + b * 64, umma_desc_sw128(ka + k * 32, 0, 1024),
                        umma_desc_sw128(ka + B_OFF_Q + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + C_DP + b * 64, umma_desc_sw128(ka + B_OFF_V + k * 32, 0, 1024),
                        umma_desc_sw128(ka + B_OFF_DO + k * 32, 0, 1024), idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&s_full[b]);
            ++ns;
            did = true;
          }
        }
        if (ng < ns) {
          if (mbar_test(pds_full, ng & 1) && mbar_test(g_free, (ng & 1) ^ 1)) {
            tc_fence_after();
            const int stage = ng % 3;
            const uint32_t ka = smem_u32(sStage + stage * B_STAGE_BYTES);
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2)
              umma_bf16(tmem_base + C_DV, umma_desc_sw128(pa + s2 * 32, 0, 1024),
                        umma_desc_sw128(ka + B_OFF_DO + s2 * 2048, 0, 1024), idesc_g, s2 > 0 ? 1u : 0u);
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2)
              umma_bf16(tmem_base + C_DK, umma_desc_sw128(sa + s2 * 32, 0, 1024),
                        umma_desc_sw128(ka + B_OFF_Q + s2 * 2048, 0, 1024), idesc_g, s2 > 0 ? 1u : 0u);
#pragma unroll
            for (int s2 = 0; s2 < 8; ++s2)
              umma_bf16(tmem_base + C_DQ, umma_desc_sw128(sa + s2 * 2048, 16384, 1024),
                        umma_desc_sw128(ka + s2 * 2048, 0, 1024), idesc_q, s2 > 0 ? 1u : 0u);
            umma_commit(g_full);
            umma_commit(pds_free);
            umma_commit(&empty[stage]);
            ++ng;
            did = true;
          }
        }
        if (did) dog.reset(); else dog.idle();
      }
    }
  } else {
    // ------------------------------------------------------------------ math: thread = key row (w, j), 32 of the 64 queries
    wt_reg_inc<192>();
    const int half = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane, w = r >> 6, j = r & 63;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * WT_L2E;
    const uint32_t brow = smem_u32(sBias) + (uint32_t)j * 256u;   // shared-space addresses (see the forward kernel)
    const int sx = j & 7;
    bf16* dQ = reinterpret_cast<bf16*>(p.dq);
    bf16* dK = reinterpret_cast<bf16*>(p.dk);
    bf16* dV = reinterpret_cast<bf16*>(p.dv);
    float acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.f;
    int ptok_r = -1, ptok_q = -1;

    auto epilogue = [&](int n, int tok_r, int tok_q) {
      mbar_wait(g_full, n & 1);
      tc_fence_after();
      uint32_t a[32], bq[32];
      tmem_ld_32x32(t_lane + (half ? C_DK : C_DV) + w * 32, a);
      if (quarter < 2) tmem_ld_32x32(t_lane + C_DQ + half * 32, bq);
      tmem_ld_wait();
      tc_fence_before();
      wt_warp_arrive(g_free, lane);
      if (tok_r >= 0) {
        if (half) wt_store_row32(dK + (long)tok_r * p.lddk + h * 32, a, p.scale);
        else      wt_store_row32(dV + (long)tok_r * p.lddv + h * 32, a, 1.f);
      }
      if (quarter < 2 && tok_q >= 0) wt_store_row32(dQ + (long)tok_q * p.lddq + h * 32, bq, p.scale);
    };

    for (int n = 0; n < my_tiles; ++n) {
      const int stage = n % 3, b = n & 1;
      mbar_wait(&full[stage], (n / 3) & 1);
      const uint32_t side = smem_u32(sSide) + (uint32_t)(stage * B_SIDE_BYTES);
      const int tok_r = (int)lds_u32(side + r * 4);
      const int tok_q = (int)lds_u32(side + (half * 64 + j) * 4);  // dQ row (query j of window `half`)
      const uint32_t lrow = side + 512 + (uint32_t)(w * 64 + half * 32) * 4u;
      const uint32_t drow = side + 1024 + (uint32_t)(w * 64 + half * 32) * 4u;
      const unsigned long long mb64 = wt_lds_u64(side + 1536 + r * 8);
      const uint32_t mb = half ? (uint32_t)(mb64 >> 32) : (uint32_t)mb64;
      mbar_wait(&s_full[b], (n >> 1) & 1);
      tc_fence_after();
      uint32_t st[32], dp[32];
      tmem_ld_32x32(t_lane + C_ST + b * 64 + half * 32, st);
      tmem_ld_32x32(t_lane + C_DP + b * 64 + half * 32, dp);
      tmem_ld_wait();
      tc_fence_before();
      wt_warp_arrive(&s_free[b], lane);
      const bool any_mask = __any_sync(0xffffffffu, mb != 0u);
      uint32_t pp[16], dd[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 bb = lds_f4(brow + (uint32_t)(((half * 8 + c) ^ sx) << 4));
        const float4 ll = lds_f4(lrow + c * 16);
        const float4 dl = lds_f4(drow + c * 16);
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, lv[4] = {ll.x, ll.y, ll.z, ll.w}, dv[4] = {dl.x, dl.y, dl.z, dl.w};
        float pv[4], sv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = fmaf(__uint_as_float(st[c * 4 + e]), sl2, bv[e]) - lv[e];
          if (any_mask && (mb & (1u << (c * 4 + e)))) x = -INFINITY;
          pv[e] = wt_ex2(x);
          sv[e] = pv[e] * (__uint_as_float(dp[c * 4 + e]) - dv[e]);
          acc[c * 4 + e] += sv[e];
        }
        pp[c * 2] = pack_bf16(pv[0], pv[1]); pp[c * 2 + 1] = pack_bf16(pv[2], pv[3]);
        dd[c * 2] = pack_bf16(sv[0], sv[1]); dd[c * 2 + 1] = pack_bf16(sv[2], sv[3]);
      }
      mbar_wait(pds_free, (n & 1) ^ 1);  // the previous tile's gradient MMAs no longer read the tiles
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t off = sw128(r, half * 4 + q4);
        sts_u4(smem_u32(sP) + off, pp[q4 * 4], pp[q4 * 4 + 1], pp[q4 * 4 + 2], pp[q4 * 4 + 3]);
        sts_u4(smem_u32(sdS) + off, dd[q4 * 4], dd[q4 * 4 + 1], dd[q4 * 4 + 2], dd[q4 * 4 + 3]);
      }
      wt_proxy_fence();
      wt_warp_arrive(pds_full, lane);
      if (n > 0) epilogue(n - 1, ptok_r, ptok_q);
      ptok_r = tok_r;
      ptok_q = tok_q;
    }
    epilogue(my_tiles - 1, ptok_r, ptok_q);

    // bias gradient: registers -> per-CTA table in shared memory (over the retired P^T tile) -> global atomics
    if (has_tab) {
      float* dtab = reinterpret_cast<float*>(sP);
      const int mt = threadIdx.x - 128;  // 0..255
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int e = mt; e < p.n_pos; e += 256) dtab[e] = 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (j < g.nq) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int i = half * 32 + c;
          if (i < g.nq && acc[c] != 0.f) atomicAdd(&dtab[__ldg(p.pos + i * g.nq + j)], acc[c]);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int e = mt; e < p.n_pos; e += 256) {
        const float v = dtab[e];
        if (v != 0.f) atomicAdd(p.drel_bias + (long)e * g.heads + h, v);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool g_attn_wt = true;

int wt_geom(const vtb_attn_params* p, WtGeom* g) {
  g->heads = p->heads; g->nq = p->nq; g->Hs = p->Hs; g->Ws = p->Ws; g->window = p->window; g->shift = p->shift;
  g->nwx = p->Ws / p->window;
  g->nw = (p->Hs / p->window) * g->nwx;
  const long groups = (long)p->batch * g->nw;
  VTB_CHECK(groups < (1L << 30) && (long)p->batch * p->Hs * p->Ws < (1L << 31), -1,
            "vtb_attention(window tcgen05): problem too large for 32-bit token indices");
  g->groups = (int)groups;
  g->inv_nw = 1.f / (float)g->nw;
  g->inv_nwx = 1.f / (float)g->nwx;
  VTB_CHECK(groups < (1L << 22), -1, "vtb_attention(window tcgen05): too many windows");
  return 0;
}

template <typename K>
int wt_launch(K kern, size_t smem, const vtb_attn_params* p, cudaStream_t stream, bool* attr_set) {
  WtGeom g;
  if (int rc = wt_geom(p, &g)) return rc;
  if (!*attr_set) {
    VTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    *attr_set = true;
  }
  const int ntiles = (g.groups + 1) / 2;
  int nchunks = vtb_num_sms() / p->heads;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > ntiles) nchunks = ntiles;
  kern<<<(unsigned)(nchunks * p->heads), WT_THREADS, smem, stream>>>(*p, g, ntiles, nchunks);
  VTB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

void vtb_attn_wt_set(bool on) { g_attn_wt = on; }

bool vtb_attn_wt_ok(const vtb_attn_params* p, bool bwd) {
  if (!g_attn_wt || p->mode != VTB_ATTN_WINDOW || p->dh != 32 || p->nq > 64 || p->nkv != p->nq) return false;
  if (p->mask && !p->mask_bits) return false;
  if (p->rel_bias && p->n_pos > 512) return false;
  if (p->heads > 148) return false;
  uintptr_t al = (uintptr_t)p->q | (uintptr_t)p->k | (uintptr_t)p->v | (uintptr_t)p->o;
  int ld = p->ldq | p->ldk | p->ldv | p->ldo;
  if (bwd) {
    if (p->dkv_f32) return false;
    al |= (uintptr_t)p->dout | (uintptr_t)p->dq | (uintptr_t)p->dk | (uintptr_t)p->dv;
    ld |= p->lddo | p->lddq | p->lddk | p->lddv;
  }
  return (al & 15) == 0 && (ld & 7) == 0;
}

int vtb_attn_wt_fwd(const vtb_attn_params* p, cudaStream_t stream) {
  static bool attr = false;
  return wt_launch(attn_wt_fwd_kernel, (size_t)F_SMEM, p, stream, &attr);
}
int vtb_attn_wt_bwd(const vtb_attn_params* p, cudaStream_t stream) {
  static bool attr = false;
  return wt_launch(attn_wt_bwd_kernel, (size_t)B_SMEM, p, stream, &attr);
}
