// tcgen05 / TMEM blocked local attention with a zero-padded halo (HaloNet: W^2 <= 64 queries per block, (W + 2 halo)^2
// <= 176 key slots, dh = 32), forward and backward.  halo_transformer.py:57-114 (F.unfold gather, rel_pos bias, softmax).
//
// Same tile algebra as the window kernels (attention_win_tc.cu): TWO blocks share every 128-row tile and are kept apart by
// stacking them along the contraction axis (A rows of block w hold their 32 features in K columns [32 w, 32 w + 32), B rows
// hold [block 0 | block 1]), so the products come out block-diagonal with nothing to mask.  What is different:
//   * the keys / values of a block are the (W + 2 halo)^2 slots of its halo window: the loaders resolve slot -> token
//     (or "outside the image" -> cp.async zero fill: the reference pads with zeros, the slot still takes part in the softmax
//     with score = bias) — the unfold never exists in memory;
//   * forward: S is 128 x 176 in TMEM (N = 176 in one UMMA), the softmax makes two passes over TMEM (row max, then
//     probabilities written back over S as bf16 pairs), O = P V takes P straight from TMEM (11 K-steps);
//   * backward: keys sit on the tile rows as in the window kernel, so a block pair is cut into 64-slot key chunks; every chunk
//     is one window-style tile (S^T, dP^T, dV, dK complete per chunk; dQ accumulates in TMEM over the chunks of a pair).
//     Key / value tokens belong to up to four halos, so dK / dV leave as PER-BLOCK partial rows (bf16, plain coalesced
//     stores, workspace [2][block][head][slot][32]) and a second small kernel sums, for every token, the <= 4 partial rows that
//     cover it in a fixed order (fp32) — no atomics, run-to-run deterministic.
//     The bias gradient stays in registers (one accumulator set per key chunk) until the CTA's last tile.
// One persistent CTA per SM with a fixed head; 4 loader warps (16-byte cp.async gathers whose completion fires the stage
// barriers), one UMMA issuer warp, 8 math warps (thread = TMEM lane = tile row); registers re-balanced with setmaxnreg.
#include "attn_win_common.cuh"
#include "../../include/vtb200.h"
#include <math.h>

namespace {

constexpr int HT_MAXK = 176;  // key slots per block, padded to a multiple of 16 ((7 + 2 * 3)^2 = 169)

struct HtGeom {
  int heads, nq, nkv, Hs, Ws, window, halo, kw, nbx, nb, groups;
  int nkc;  // 16-slot steps covering nkv
  int nch;  // 64-slot chunks covering nkv (backward tiles per block pair)
  float inv_nb, inv_nbx, inv_kw, inv_w;
};
struct HtOrigin { int img, y0, x0; };  // top-left token of the block (without halo); img < 0: no such block

__device__ __forceinline__ HtOrigin ht_origin(const HtGeom& g, int grp) {
  HtOrigin o;
  if (grp >= g.groups) { o.img = -1; o.y0 = o.x0 = 0; return o; }
  const int b = wt_div(grp, g.nb, g.inv_nb), bi = grp - b * g.nb;
  const int by = wt_div(bi, g.nbx, g.inv_nbx), bx = bi - by * g.nbx;
  o.img = b; o.y0 = by * g.window; o.x0 = bx * g.window;
  return o;
}
// token of query t of the block, -1 for padding
__device__ __forceinline__ int ht_q_token(const HtGeom& g, const HtOrigin& o, int t) {
  if (t >= g.nq || o.img < 0) return -1;
  const int ty = wt_div(t, g.window, g.inv_w), tx = t - ty * g.window;
  return (o.img * g.Hs + o.y0 + ty) * g.Ws + o.x0 + tx;
}
// token of key slot j of the block's halo window, -1 where the slot is zero padding (halo_transformer.py:74-80)
__device__ __forceinline__ int ht_k_token(const HtGeom& g, const HtOrigin& o, int j) {
  if (j >= g.nkv || o.img < 0) return -1;
  const int ky = wt_div(j, g.kw, g.inv_kw), kx = j - ky * g.kw;
  const int y = o.y0 - g.halo + ky, x = o.x0 - g.halo + kx;
  if (y < 0 || y >= g.Hs || x < 0 || x >= g.Ws) return -1;
  return (o.img * g.Hs + y) * g.Ws + x;
}

// packed fp32 pairs (FFMA2 / FADD2 / FMUL2: one issue slot for two lanes of work — the math warps are issue-bound)
__device__ __forceinline__ uint64_t ht_pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t ht_pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void ht_upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ht_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t ht_add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t ht_mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// Tile-invariant part of ht_k_token / ht_q_token for a loader thread: slot -> (dy, dx) relative to the block origin, packed
// as two signed 16-bit fields (0x7fff7fff: no such slot).  Per tile only the bounds test and one multiply-add remain.
__device__ __forceinline__ int ht_slot_pack(const HtGeom& g, int j) {
  if (j >= g.nkv) return 0x7fff7fff;
  const int ky = j / g.kw, kx = j - ky * g.kw;
  return ((ky - g.halo) & 0xffff) | ((kx - g.halo) << 16);
}
__device__ __forceinline__ int ht_query_pack(const HtGeom& g, int t) {
  if (t >= g.nq) return 0x7fff7fff;
  const int ty = t / g.window, tx = t - ty * g.window;
  return ty | (tx << 16);
}
// token of a packed slot for the block at `o` (base = token of the block origin), -1 outside the image / no slot / no block
__device__ __forceinline__ int ht_pack_token(const HtGeom& g, const HtOrigin& o, int base, int pk) {
  const int dy = (int)(short)(pk & 0xffff), dx = pk >> 16;
  const int y = o.y0 + dy, x = o.x0 + dx;
  const bool ok = o.img >= 0 && (unsigned)y < (unsigned)g.Hs && (unsigned)x < (unsigned)g.Ws;
  return ok ? base + dy * g.Ws + dx : -1;
}

// ---------------------------------------------------------------------------------------------------------------
// forward
//   S[(w,i), j] = Qpad . Kcat^T  (M 128, N = 16 nkc <= 176, K 64)  -> two-pass softmax per row -> P (bf16) over S in TMEM
//   O[(w,i), (w',d)] = P . Vcat  (A from TMEM, B MN-major, nkc K-steps); columns w' = w are the output.
//   3 operand stages (Qpad 16 KB | Kcat 22 KB | Vcat 22 KB), 2 TMEM buffers (S 176 | O 64), the two math groups take
//   alternate tiles.  Bias tile: fp32 [64 queries][16 nkc slots] x log2 e, -inf on slot padding, 16-byte chunks XOR-swizzled.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HF_STAGES = 3;
constexpr int HF_OFF_K = 16384, HF_OFF_V = 16384 + HT_MAXK * 128;
constexpr int HF_STAGE_BYTES = 16384 + 2 * HT_MAXK * 128;  // 61440
constexpr int HF_BIAS_BYTES = 64 * HT_MAXK * 4;            // 45056
constexpr int HF_SMEM = HF_STAGES * HF_STAGE_BYTES + HF_BIAS_BYTES + 256 + 1024;
constexpr uint32_t HF_TBUF = 256, HF_TO = 192;
static_assert(HF_SMEM <= 232448, "forward shared memory");
static_assert(HF_OFF_V % 1024 == 0 && HF_STAGE_BYTES % 1024 == 0, "operand tiles must be 1024-byte aligned");

// physical 16-byte chunk of logical chunk ch in a bias row of `nchunk` chunks: XOR with the row inside whole groups of 8,
// with the row's low two bits inside a trailing group of 4 (conflict-free float4 reads down a column for both row pitches)
__device__ __forceinline__ int ht_bias_chunk(int ch, int full, int s7, int s3) {
  return ch < full ? (ch ^ s7) : full + ((ch - full) ^ s3);
}

template <int NKC>
__global__ void __launch_bounds__(WT_THREADS, 1)
attn_ht_fwd_kernel(vtb_attn_params p, HtGeom g, int ntiles, int nchunks) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sStage = smem;
  uint8_t* sBias = smem + HF_STAGES * HF_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + HF_BIAS_BYTES);
  uint64_t* full = bars;         // [3] 128 cp.async arrivals (loader threads) -> issuer
  uint64_t* empty = bars + 3;    // [3] PV retired -> loaders
  uint64_t* s_full = bars + 6;   // [2] S complete
  uint64_t* p_full = bars + 8;   // [2] P written (4 warps)
  uint64_t* o_full = bars + 10;  // [2] O complete
  uint64_t* o_free = bars + 12;  // [2] O drained (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % g.heads;
  const int chunk = blockIdx.x / g.heads;
  const int my_tiles = (ntiles - chunk + nchunks - 1) / nchunks;
  constexpr int nkc = NKC;              // 16-slot steps (compile time: the unit loops unroll, bias addresses become immediates)
  constexpr int pitch = nkc * 64;        // bias row pitch in bytes
  constexpr int full8 = (nkc * 4) & ~7;  // chunks in whole groups of 8

  if (threadIdx.x == 0) {
    for (int i = 0; i < HF_STAGES; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  // zero the operand ring once: the zero halves of Qpad, the query rows >= nq and the slot rows >= nkv never change
  for (int e = threadIdx.x; e < HF_STAGES * HF_STAGE_BYTES / 16; e += WT_THREADS)
    reinterpret_cast<uint4*>(sStage)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    float* tab = reinterpret_cast<float*>(sStage);  // scratch: the head's column of rel_pos.weight
    if (p.rel_bias)
      for (int t = threadIdx.x; t < p.n_pos; t += WT_THREADS) tab[t] = __ldg(p.rel_bias + (long)t * g.heads + h) * WT_L2E;
    __syncthreads();
    const int ncol = nkc * 16;
    for (int e = threadIdx.x; e < 64 * ncol; e += WT_THREADS) {
      const int row = e / ncol, col = e - row * ncol;
      float v = 0.f;
      if (col >= g.nkv) v = -INFINITY;
      else if (row < g.nq && p.rel_bias) v = tab[__ldg(p.pos + row * g.nkv + col)];
      *reinterpret_cast<float*>(sBias + row * pitch + ht_bias_chunk(col >> 2, full8, row & 7, row & 3) * 16 + (col & 3) * 4) = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2048 / 16; e += WT_THREADS) reinterpret_cast<uint4*>(sStage)[e] = make_uint4(0, 0, 0, 0);
  }
  wt_proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders (128 threads)
    // thread = (block w of the tile, 16-byte chunk cc, row r4 + 16 it): four lanes cover one 64-byte token row
    wt_reg_dec<80>();
    const int tt = threadIdx.x, w = tt >> 6, tl = tt & 63, cc = tl & 3, r4 = tl >> 2;
    const uint32_t ch = (uint32_t)(((w * 4 + cc) ^ (r4 & 7)) << 4);
    const uint32_t st0 = smem_u32(sStage) + (uint32_t)r4 * 128u + ch;
    const char* qb = reinterpret_cast<const char*>(p.q) + (h * 32 + cc * 8) * 2;
    const char* kb = reinterpret_cast<const char*>(p.k) + (h * 32 + cc * 8) * 2;
    const char* vb = reinterpret_cast<const char*>(p.v) + (h * 32 + cc * 8) * 2;
    const long ldq2 = (long)p.ldq * 2, ldk2 = (long)p.ldk * 2, ldv2 = (long)p.ldv * 2;
    int qpk[4], kpk[nkc];
#pragma unroll
    for (int it = 0; it < 4; ++it) qpk[it] = ht_query_pack(g, it * 16 + r4);
#pragma unroll
    for (int it = 0; it < nkc; ++it) kpk[it] = ht_slot_pack(g, it * 16 + r4);
    for (int n = 0; n < my_tiles; ++n) {
      const int tile = chunk + n * nchunks;
      const int stage = n % HF_STAGES;
      mbar_wait(&empty[stage], ((n / HF_STAGES) & 1) ^ 1);
      const uint32_t st = st0 + stage * HF_STAGE_BYTES;
      const HtOrigin org = ht_origin(g, tile * 2 + w);
      const int base = (org.img * g.Hs + org.y0) * g.Ws + org.x0;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (qpk[it] != 0x7fff7fff) {  // rows >= nq stay zero from the prologue
          const int tok = ht_pack_token(g, org, base, qpk[it]);
          cp_async16(st + (uint32_t)(w * 64 + it * 16) * 128u, qb + (long)(tok < 0 ? 0 : tok) * ldq2, tok >= 0);
        }
      }
#pragma unroll
      for (int it = 0; it < nkc; ++it) {
        if (kpk[it] != 0x7fff7fff) {  // slots >= nkv stay zero from the prologue
          const int tok = ht_pack_token(g, org, base, kpk[it]);
          const long gr = tok < 0 ? 0 : tok;
          const uint32_t ro = (uint32_t)(it * 16) * 128u;
          cp_async16(st + HF_OFF_K + ro, kb + gr * ldk2, tok >= 0);
          cp_async16(st + HF_OFF_V + ro, vb + gr * ldv2, tok >= 0);
        }
      }
      cp_async_arrive_noinc(&full[stage]);
    }
    cp_async_wait<0>();  // nothing may still be in flight towards this CTA's shared memory at exit
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ UMMA issuer (warp 12)
    // All 32 lanes run the warp-uniform loop; only tcgen05.mma / commit sit under elect.sync, so descriptors live in uniform
    // registers and a K-step is "previous descriptor + constant" (a single-thread issuer spends ~15 instructions per MMA on
    // descriptor arithmetic and vector -> uniform moves: the issuer's instruction stream is the critical path of a tile).
    wt_reg_dec<40>();
    if (warp == 12) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, nkc * 16, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      const uint64_t dq0 = umma_desc_sw128(smem_u32(sStage), 0, 1024);  // + (byte offset >> 4) addresses any other tile
      int ns = 0, np = 0;  // event-driven: whichever of "next score tile" / "next P.V" has its inputs ready is issued
      WtWatchdog dog;
      dog.reset();
      while (np < my_tiles) {
        bool did = false;
        if (ns < my_tiles && ns < np + 2) {
          const int b = ns & 1, stage = ns % HF_STAGES;
          // S(ns) overwrites the S / P columns of tile ns - 2: free once that tile's P.V has completed (its O may still drain)
          if (mbar_test(&full[stage], (ns / HF_STAGES) & 1) && mbar_test(&o_full[b], ((ns >> 1) & 1) ^ 1)) {
            wt_proxy_fence();  // cp.async (generic proxy) writes -> visible to the tensor core's async-proxy reads
            tc_fence_after();
            const uint64_t dq = dq0 + (uint64_t)((stage * HF_STAGE_BYTES) >> 4), dk = dq + (uint64_t)(HF_OFF_K >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + b * HF_TBUF, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
              umma_commit(&s_full[b]);
            }
            __syncwarp();
            ++ns;
            did = true;
          }
        }
        if (np < ns) {
          const int b = np & 1, stage = np % HF_STAGES;
          if (mbar_test(&p_full[b], (np >> 1) & 1) && mbar_test(&o_free[b], ((np >> 1) & 1) ^ 1)) {
            tc_fence_after();
            const uint64_t dv = dq0 + (uint64_t)((stage * HF_STAGE_BYTES + HF_OFF_V) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < nkc; ++k)
                wt_umma_ts(tmem_base + b * HF_TBUF + HF_TO, tmem_base + b * HF_TBUF + k * 8, dv + (uint64_t)(k * 128), idesc_o,
                           k > 0 ? 1u : 0u);
              umma_commit(&o_full[b]);
              umma_commit(&empty[stage]);
            }
            __syncwarp();
            ++np;
            did = true;
          }
        }
        if (did) dog.reset(); else { dog.idle(); __nanosleep(20); }
      }
    }
  } else {
    // ------------------------------------------------------------------ math: thread = query row (w, i)
    wt_reg_inc<192>();
    const int grpi = (warp - 4) >> 2;  // math group = TMEM buffer: tiles n = grpi, grpi + 2, ...
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane, w = r >> 6, i = r & 63;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + grpi * HF_TBUF;
    const float sl2 = p.scale * WT_L2E;
    // bias row of this thread: the address of logical 16-byte chunk c is bx[c & 7] + (c & ~7) * 16 inside whole groups of 8
    // and bt[c - full8] in a trailing group of 4 (the XOR swizzle folded into per-thread bases, the rest is an immediate)
    uint32_t bx[8], bt[4];
#pragma unroll
    for (int k = 0; k < 8; ++k) bx[k] = smem_u32(sBias) + (uint32_t)(i * pitch) + (uint32_t)((k ^ (i & 7)) << 4);
#pragma unroll
    for (int k = 0; k < 4; ++k) bt[k] = smem_u32(sBias) + (uint32_t)(i * pitch) + (uint32_t)((full8 + (k ^ (i & 3))) << 4);
    const uint64_t sl22 = ht_pk2(sl2, sl2);
    bf16* O = reinterpret_cast<bf16*>(p.o);
    for (int n = grpi; n < my_tiles; n += 2) {
      const int tile = chunk + n * nchunks;
      const int grp = tile * 2 + w;
      const int tok = ht_q_token(g, ht_origin(g, grp), i);
      mbar_wait(&s_full[grpi], (n >> 1) & 1);
      tc_fence_after();
      uint32_t sa[16], sb[16];
      // pass 1: row maximum of x = s * sl2 + bias (one TMEM load kept in flight)
      float mx0 = -INFINITY, mx1 = -INFINITY;
      tmem_ld_32x16(t_lane, sa);
#pragma unroll
      for (int u = 0; u < nkc; ++u) {
        uint32_t (&cur)[16] = (u & 1) ? sb : sa;
        uint32_t (&nxt)[16] = (u & 1) ? sa : sb;
        tmem_ld_wait();
        if (u + 1 < nkc) tmem_ld_32x16(t_lane + (u + 1) * 16, nxt);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = u * 4 + q;
          const float4 bb = lds_f4(c < full8 ? bx[c & 7] + (uint32_t)((c & ~7) << 4) : bt[(c - full8) & 3]);
          float x0, x1, x2, x3;
          ht_upk2(ht_fma2(ht_pk2u(cur[q * 4 + 0], cur[q * 4 + 1]), sl22, ht_pk2(bb.x, bb.y)), x0, x1);
          ht_upk2(ht_fma2(ht_pk2u(cur[q * 4 + 2], cur[q * 4 + 3]), sl22, ht_pk2(bb.z, bb.w)), x2, x3);
          mx0 = fmaxf(mx0, fmaxf(x0, x1));
          mx1 = fmaxf(mx1, fmaxf(x2, x3));
        }
      }
      const float mx = fmaxf(mx0, mx1);
      const float m_use = (mx == -INFINITY) ? 0.f : mx;
      const uint64_t nm2 = ht_pk2(-m_use, -m_use);
      // pass 2: probabilities; P unit u lands in 32-bit columns [8 u, 8 u + 8): below every S column still to be read
      uint64_t sum2a = 0ull, sum2b = 0ull;
      tmem_ld_32x16(t_lane, sa);
#pragma unroll
      for (int u = 0; u < nkc; ++u) {
        uint32_t (&cur)[16] = (u & 1) ? sb : sa;
        uint32_t (&nxt)[16] = (u & 1) ? sa : sb;
        tmem_ld_wait();
        if (u + 1 < nkc) tmem_ld_32x16(t_lane + (u + 1) * 16, nxt);
        uint32_t pk[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = u * 4 + q;
          const float4 bb = lds_f4(c < full8 ? bx[c & 7] + (uint32_t)((c & ~7) << 4) : bt[(c - full8) & 3]);
          float x0, x1, x2, x3;
          ht_upk2(ht_fma2(ht_pk2u(cur[q * 4 + 0], cur[q * 4 + 1]), sl22, ht_add2(ht_pk2(bb.x, bb.y), nm2)), x0, x1);
          ht_upk2(ht_fma2(ht_pk2u(cur[q * 4 + 2], cur[q * 4 + 3]), sl22, ht_add2(ht_pk2(bb.z, bb.w), nm2)), x2, x3);
          const float p0 = wt_ex2(x0), p1 = wt_ex2(x1), p2 = wt_ex2(x2), p3 = wt_ex2(x3);
          sum2a = ht_add2(sum2a, ht_pk2(p0, p1));
          sum2b = ht_add2(sum2b, ht_pk2(p2, p3));
          pk[q * 2] = pack_bf16(p0, p1);
          pk[q * 2 + 1] = pack_bf16(p2, p3);
        }
        wt_tmem_st8(t_lane + u * 8, pk);
      }
      float sum;
      {
        float a0, a1;
        ht_upk2(ht_add2(sum2a, sum2b), a0, a1);
        sum = a0 + a1;
      }
      wt_tmem_st_wait();
      tc_fence_before();
      wt_warp_arrive(&p_full[grpi], lane);
      // epilogue: O / l -> bf16 -> global; lse = ln 2 * (max + log2(sum))
      mbar_wait(&o_full[grpi], (n >> 1) & 1);
      tc_fence_after();
      uint32_t a[32];
      tmem_ld_32x32(t_lane + HF_TO + w * 32, a);
      tmem_ld_wait();
      tc_fence_before();
      wt_warp_arrive(&o_free[grpi], lane);
      if (tok >= 0) {
        wt_store_row32(O + (long)tok * p.ldo + h * 32, a, 1.f / sum);
        if (p.lse) p.lse[((long)grp * g.heads + h) * g.nq + i] = (mx + log2f(sum)) * WT_LN2;
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

// ---------------------------------------------------------------------------------------------------------------
// backward: tile = (block pair, 64-slot key chunk c); keys on the tile rows
//   S^T[(w,jj), i] = Kpad . Qcat^T      dP^T[(w,jj), i] = Vpad . dOcat^T             (M 128, N 64, K 64)
//   P^T = exp2(S^T sl2 + bias - lse2[i]),  dS^T = P^T (dP^T - delta[i])   -> bf16 tiles in shared memory
//   dV[(w,jj), (w',d)] = P^T . dOcat    dK = dS^T . Qcat   (complete per chunk -> per-block partial rows)
//   dQ[i, (w,d)]      += dS . Kpad      (accumulates over the chunks of the pair, drained after the last one)
//   P^T and dS^T go back into TMEM (bf16 pairs over S^T / dP^T) and feed dV / dK as TMEM A operands; dS^T is also written
//   to one of two shared-memory buffers for dQ — so the math of tile n + 1 never waits for the gradient MMAs of tile n.
// TMEM: S^T 2 x 64 | dP^T 2 x 64 | dV 64 | dK 64 | dQ 64.   Rings: 3 key/value chunk stages (Kpad 16 KB | Vpad 16 KB),
// 2 query stages (Qcat 8 KB | dOcat 8 KB | O rows 8 KB for delta) loaded once per pair.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HB_KV_STAGES = 3, HB_KV_BYTES = 32768;
constexpr int HB_Q_BYTES = 24576, HB_OFF_DO = 8192, HB_OFF_O = 16384;
constexpr int HB_BIAS_ROWS = HT_MAXK + 8;  // slots, plus a guard row that is -inf whatever nkv is
constexpr int HB_SIDE_BYTES = 128 * 4 * 3;  // query token, lse * log2 e, delta per tile row (w, i)
constexpr int HB_OFF_Q = HB_KV_STAGES * HB_KV_BYTES;
constexpr int HB_OFF_DS = HB_OFF_Q + 2 * HB_Q_BYTES;
constexpr int HB_OFF_P = HB_OFF_DS + 16384;
constexpr int HB_OFF_BIAS = HB_OFF_P + 16384;
constexpr int HB_OFF_SIDE = HB_OFF_BIAS + HB_BIAS_ROWS * 256;
constexpr int HB_OFF_BARS = HB_OFF_SIDE + 2 * HB_SIDE_BYTES;
constexpr int HB_SMEM = HB_OFF_BARS + 256 + 1024;
static_assert(HB_SMEM <= 232448, "backward shared memory");
constexpr uint32_t HC_ST = 0, HC_DP = 128, HC_DV = 256, HC_DK = 320, HC_DQ = 384;

__global__ void __launch_bounds__(WT_THREADS, 1)
attn_ht_bwd_kernel(vtb_attn_params p, HtGeom g, int npairs, int nchunks, bf16* part_k, bf16* part_v, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sKV = smem;
  uint8_t* sQ = smem + HB_OFF_Q;
  uint8_t* sdS = smem + HB_OFF_DS;   // [2][128 key rows][64 queries] bf16, 128B-swizzled: dS^T of tile n in buffer n & 1 (the
                                     // dQ product reads it MN-major; its don't-care second M atom lands 16 KB further on)
  uint8_t* sP = smem + HB_OFF_P;     // = buffer 1 of sdS; scratch for the bias tables before / after the tile loop
  uint8_t* sBias = smem + HB_OFF_BIAS;
  uint8_t* sSide = smem + HB_OFF_SIDE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + HB_OFF_BARS);
  uint64_t* kv_land = bars;         // [3] 128 cp.async arrivals
  uint64_t* kv_empty = bars + 3;    // [3] gradient MMAs of the chunk retired
  uint64_t* q_land = bars + 6;      // [2] 128 cp.async arrivals
  uint64_t* q_full = bars + 8;      // [2] count 4: delta / lse of the landed pair in place
  uint64_t* q_empty = bars + 10;    // [2] gradient MMAs of the pair's last chunk retired
  uint64_t* s_full = bars + 12;     // [2] S^T / dP^T complete
  uint64_t* s_free = bars + 14;     // [2] the gradient MMAs of the tile have consumed P^T / dS^T (TMEM columns + sdS buffer)
  uint64_t* pds_full = bars + 16;   // [2] P^T / dS^T of the tile written (8 warps)
  uint64_t* g_full = bars + 18;     // dV / dK (/ dQ) complete
  uint64_t* g_free = bars + 19;     // ... and drained (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % g.heads;
  const int chunk = blockIdx.x / g.heads;
  const int my_pairs = (npairs - chunk + nchunks - 1) / nchunks;
  const int nch = g.nch;
  const int my_tiles = my_pairs * nch;
  const bool has_tab = p.rel_bias != nullptr && p.drel_bias != nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&kv_land[i], 128); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_land[i], 128); mbar_init(&q_full[i], 4); mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 1); mbar_init(&pds_full[i], 8);
    }
    mbar_init(g_full, 1); mbar_init(g_free, 8);
    mbar_fence_init();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  for (int e = threadIdx.x; e < HB_OFF_BIAS / 16; e += WT_THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    // bias tile, keys on the rows: tile[j][i] = rel_bias[pos[i, j], h] * log2 e; -inf rows for j >= nkv (incl. the guard row)
    float* tab = reinterpret_cast<float*>(sP);
    if (p.rel_bias)
      for (int t = threadIdx.x; t < p.n_pos; t += WT_THREADS) tab[t] = __ldg(p.rel_bias + (long)t * g.heads + h) * WT_L2E;
    __syncthreads();
    for (int e = threadIdx.x; e < HB_BIAS_ROWS * 64; e += WT_THREADS) {
      const int row = e >> 6, col = e & 63;
      float v = 0.f;
      if (row >= g.nkv) v = -INFINITY;
      else if (col < g.nq && p.rel_bias) v = tab[__ldg(p.pos + col * g.nkv + row)];
      *reinterpret_cast<float*>(sBias + row * 256 + ((((col >> 2) ^ (row & 7))) << 4) + (col & 3) * 4) = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2048 / 16; e += WT_THREADS) reinterpret_cast<uint4*>(sP)[e] = make_uint4(0, 0, 0, 0);
  }
  wt_proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders (128 threads)
    wt_reg_dec<72>();
    const int tt = threadIdx.x, w = tt >> 6, tl = tt & 63, cc = tl & 3, r4 = tl >> 2;
    const uint32_t ch = (uint32_t)(((w * 4 + cc) ^ (r4 & 7)) << 4);
    const uint32_t kv0 = smem_u32(sKV) + (uint32_t)(w * 64 + r4) * 128u + ch;  // Kpad row (w, r4 + 16 it)
    const uint32_t q0 = smem_u32(sQ) + (uint32_t)r4 * 128u + ch;               // Qcat / dOcat row r4 + 16 it
    const uint32_t o0 = smem_u32(sQ) + HB_OFF_O + (uint32_t)(w * 64 + r4) * 64u + cc * 16;
    const int colb = (h * 32 + cc * 8) * 2;
    const char* qb = reinterpret_cast<const char*>(p.q) + colb;
    const char* kb = reinterpret_cast<const char*>(p.k) + colb;
    const char* vb = reinterpret_cast<const char*>(p.v) + colb;
    const char* dob = reinterpret_cast<const char*>(p.dout) + colb;
    const char* ob = reinterpret_cast<const char*>(p.o) + colb;
    const long ldq2 = (long)p.ldq * 2, ldk2 = (long)p.ldk * 2, ldv2 = (long)p.ldv * 2, lddo2 = (long)p.lddo * 2,
               ldo2 = (long)p.ldo * 2;
    int qpk[4], kpk[3][4];
    const int qpk_side = ht_query_pack(g, tl);
#pragma unroll
    for (int it = 0; it < 4; ++it) qpk[it] = ht_query_pack(g, it * 16 + r4);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int it = 0; it < 4; ++it) kpk[c][it] = ht_slot_pack(g, c * 64 + it * 16 + r4);
    // queries of pair pi: rows, lse, token ids
    auto issue_q = [&](int pi) {
      const int qs = pi & 1;
      const int grp = (chunk + pi * nchunks) * 2 + w;
      const HtOrigin org = ht_origin(g, grp);
      const int base = (org.img * g.Hs + org.y0) * g.Ws + org.x0;
      const uint32_t side = smem_u32(sSide) + (uint32_t)(qs * HB_SIDE_BYTES);
      const int tok_side = qpk_side != 0x7fff7fff ? ht_pack_token(g, org, base, qpk_side) : -1;
      sts_u32(side + tt * 4, (uint32_t)tok_side);
      if (tok_side >= 0) cp_async4(side + 512 + tt * 4, p.lse + ((long)grp * g.heads + h) * g.nq + tl);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (qpk[it] != 0x7fff7fff) {  // rows >= nq stay zero from the prologue
          const int tok = ht_pack_token(g, org, base, qpk[it]);
          const long gr = tok < 0 ? 0 : tok;
          const uint32_t ro = (uint32_t)(qs * HB_Q_BYTES) + (uint32_t)(it * 16) * 128u;
          cp_async16(q0 + ro, qb + gr * ldq2, tok >= 0);
          cp_async16(q0 + HB_OFF_DO + ro, dob + gr * lddo2, tok >= 0);
          cp_async16(o0 + (uint32_t)(qs * HB_Q_BYTES) + (uint32_t)(it * 16) * 64u, ob + gr * ldo2, tok >= 0);
        }
      }
      cp_async_arrive_noinc(&q_land[qs]);
    };
    // the pair's rows have landed: delta = dO . O and lse * log2 e, then release
    auto finish_q = [&](int qs) {
      const uint32_t st = smem_u32(sQ) + (uint32_t)(qs * HB_Q_BYTES);
      const uint32_t side = smem_u32(sSide) + (uint32_t)(qs * HB_SIDE_BYTES);
      float acc = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint4 ra = wt_lds_u4(st + HB_OFF_DO + sw128(tl, w * 4 + c4));
        const uint4 rb = wt_lds_u4(st + HB_OFF_O + tt * 64 + c4 * 16);
        const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fa = unpack_bf16(wa[q]), fb = unpack_bf16(wb[q]);
          acc += fa.x * fb.x + fa.y * fb.y;
        }
      }
      const int tok = (int)lds_u32(side + tt * 4);
      const float l = __uint_as_float(lds_u32(side + 512 + tt * 4));
      sts_f32(side + 512 + tt * 4, tok >= 0 ? -l * WT_L2E : -INFINITY);  // stored negated: the math warps add
      sts_f32(side + 1024 + tt * 4, -acc);
      wt_warp_arrive(&q_full[qs], lane);
    };
    // key chunk of tile n: slots [64 c, 64 c + 64) of both blocks; slots outside the image / past nkv are zero-filled
    auto issue_kv = [&](int pi, int c, int stage) {
      const HtOrigin org = ht_origin(g, (chunk + pi * nchunks) * 2 + w);
      const int base = (org.img * g.Hs + org.y0) * g.Ws + org.x0;
      const uint32_t st = kv0 + stage * HB_KV_BYTES;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int pk = c == 0 ? kpk[0][it] : (c == 1 ? kpk[1][it] : kpk[2][it]);
        const int tok = pk != 0x7fff7fff ? ht_pack_token(g, org, base, pk) : -1;
        const long gr = tok < 0 ? 0 : tok;
        const uint32_t ro = (uint32_t)(it * 16) * 128u;
        if (dbg & 2) continue;
        cp_async16(st + ro, kb + gr * ldk2, tok >= 0);
        cp_async16(st + 16384 + ro, vb + gr * ldv2, tok >= 0);
      }
      cp_async_arrive_noinc(&kv_land[stage]);
    };
    // event-driven (warp-uniform votes): the loaders never block on a load
    int nk = 0, nk_pi = 0, nk_c = 0, nqi = 0, nqf = 0;
    WtWatchdog dog;
    dog.reset();
    while (nk < my_tiles || nqf < my_pairs) {
      bool did = false;
      if (nqi < my_pairs && __all_sync(0xffffffffu, mbar_test(&q_empty[nqi & 1], ((nqi >> 1) & 1) ^ 1))) {
        issue_q(nqi);
        ++nqi;
        did = true;
      }
      if (nqf < nqi && __all_sync(0xffffffffu, mbar_test(&q_land[nqf & 1], (nqf >> 1) & 1))) {
        finish_q(nqf & 1);
        ++nqf;
        did = true;
      }
      if (nk < my_tiles && __all_sync(0xffffffffu, mbar_test(&kv_empty[nk % 3], ((nk / 3) & 1) ^ 1))) {
        issue_kv(nk_pi, nk_c, nk % 3);
        ++nk;
        if (++nk_c == nch) { nk_c = 0; ++nk_pi; }
        did = true;
      }
      if (did) dog.reset(); else { dog.idle(); __nanosleep(32); }
    }
    cp_async_wait<0>();
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ UMMA issuer (warp 12; warp-uniform, see the forward)
    wt_reg_dec<56>();
    if (warp == 12) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint32_t idesc_g = umma_idesc_bf16(128, 64, 0, 1);
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, 1, 1);
      const uint64_t d0 = umma_desc_sw128(smem_u32(smem), 0, 1024);            // + (byte offset >> 4): any tile of the block
      const uint64_t dsq0 = umma_desc_sw128(smem_u32(sdS), 16384, 1024);       // dS^T read MN-major as the dQ product's A
      int ns = 0, ng = 0;          // next score tile / next gradient tile
      int ns_pi = 0, ns_c = 0;     // (pair, chunk) of ns
      int ng_pi = 0, ng_c = 0;
      WtWatchdog dog;
      dog.reset();
      while (ng < my_tiles) {
        bool did = false;
        if (ns < my_tiles && ns < ng + 2) {
          const int stage = ns % 3, b = ns & 1, qs = ns_pi & 1;
          if (mbar_test(&kv_land[stage], (ns / 3) & 1) && mbar_test(&q_full[qs], (ns_pi >> 1) & 1) &&
              mbar_test(&s_free[b], ((ns >> 1) & 1) ^ 1)) {
            wt_proxy_fence();  // cp.async / st.shared (generic proxy) writes -> visible to the tensor core's async-proxy reads
            tc_fence_after();
            const uint64_t dk = d0 + (uint64_t)((stage * HB_KV_BYTES) >> 4), dv = dk + (uint64_t)(16384 >> 4);
            const uint64_t dq = d0 + (uint64_t)((HB_OFF_Q + qs * HB_Q_BYTES) >> 4), ddo = dq + (uint64_t)(HB_OFF_DO >> 4);
            if (elect_one()) {
              if (!(dbg & 16))
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + HC_ST + b * 64, dk + (uint64_t)(k * 2), dq + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
              if (!(dbg & 16))
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + HC_DP + b * 64, dv + (uint64_t)(k * 2), ddo + (uint64_t)(k * 2), idesc_s, k > 0 ? 1u : 0u);
              umma_commit(&s_full[b]);
            }
            __syncwarp();
            ++ns;
            if (++ns_c == nch) { ns_c = 0; ++ns_pi; }
            did = true;
          }
        }
        if (ng < ns) {
          const int b = ng & 1;
          if (mbar_test(&pds_full[b], (ng >> 1) & 1) && mbar_test(g_free, (ng & 1) ^ 1)) {
            tc_fence_after();
            const int stage = ng % 3, qs = ng_pi & 1;
            const uint64_t dk = d0 + (uint64_t)((stage * HB_KV_BYTES) >> 4);
            const uint64_t dq = d0 + (uint64_t)((HB_OFF_Q + qs * HB_Q_BYTES) >> 4), ddo = dq + (uint64_t)(HB_OFF_DO >> 4);
            const bool last = ng_c == nch - 1;
            if (elect_one()) {
              if (!(dbg & 8)) {
#pragma unroll
              // dV = P^T dO, dK = dS^T Q: A straight from TMEM (bf16 pairs over the S^T / dP^T columns each math thread has
              // read: queries [32 half + 16 hh, +16) sit in columns 32 half + 8 hh), dQ += dS K from the shared-memory copy
              for (int s2 = 0; s2 < 4; ++s2)
                wt_umma_ts(tmem_base + HC_DV, tmem_base + HC_ST + b * 64 + (s2 >> 1) * 32 + (s2 & 1) * 8,
                           ddo + (uint64_t)(s2 * 128), idesc_g, s2 > 0 ? 1u : 0u);
#pragma unroll
              for (int s2 = 0; s2 < 4; ++s2)
                wt_umma_ts(tmem_base + HC_DK, tmem_base + HC_DP + b * 64 + (s2 >> 1) * 32 + (s2 & 1) * 8,
                           dq + (uint64_t)(s2 * 128), idesc_g, s2 > 0 ? 1u : 0u);
#pragma unroll
              for (int s2 = 0; s2 < 8; ++s2)
                umma_bf16(tmem_base + HC_DQ, dsq0 + (uint64_t)(b * 1024 + s2 * 128), dk + (uint64_t)(s2 * 128), idesc_q,
                          (ng_c > 0 || s2 > 0) ? 1u : 0u);
              }
              umma_commit(g_full);
              umma_commit(&s_free[b]);
              umma_commit(&kv_empty[stage]);
              if (last) umma_commit(&q_empty[qs]);
            }
            __syncwarp();
            ++ng;
            if (++ng_c == nch) { ng_c = 0; ++ng_pi; }
            did = true;
          }
        }
        if (did) dog.reset(); else { dog.idle(); __nanosleep(20); }
      }
    }
  } else {
    // ------------------------------------------------------------------ math: thread = key row (w, jj), 32 of the 64 queries
    wt_reg_inc<192>();
    const int half = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane, w = r >> 6, jj = r & 63;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * WT_L2E;
    bf16* dQ = reinterpret_cast<bf16*>(p.dq);
    bf16* part = half ? part_k : part_v;
    const uint64_t sl22 = ht_pk2(sl2, sl2);
    uint64_t acc[3][16];  // bias gradient, packed pairs: [key chunk][query pair of this half]
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[c][e] = 0ull;
    long prow = -1;     // partial row of the previous tile (dV / dK drain is deferred by one tile), -1: nothing to store
    int ptok_q = -1;    // dQ row of the previous tile if it closed its pair
    int n = 0;

    auto epilogue = [&](int np, long row, int tok_q, bool last) {
      mbar_wait(g_full, np & 1);
      tc_fence_after();
      uint32_t a[32], bq[32];
      tmem_ld_32x32(t_lane + (half ? HC_DK : HC_DV) + w * 32, a);
      if (last && quarter < 2) tmem_ld_32x32(t_lane + HC_DQ + half * 32, bq);
      tmem_ld_wait();
      tc_fence_before();
      wt_warp_arrive(g_free, lane);
      if (dbg & 1) return;
      if (row >= 0) wt_store_row32_raw(part + row * 32, a);  // dK partials leave unscaled: the per-token sum applies the softmax scale
      if (last && quarter < 2 && tok_q >= 0) wt_store_row32(dQ + (long)tok_q * p.lddq + h * 32, bq, p.scale);
    };

    for (int pi = 0; pi < my_pairs; ++pi) {
      const int qs = pi & 1;
      const int grp = (chunk + pi * nchunks) * 2 + w;
      const uint32_t side = smem_u32(sSide) + (uint32_t)(qs * HB_SIDE_BYTES);
      const uint32_t lrow = side + 512 + (uint32_t)(w * 64 + half * 32) * 4u;
      const uint32_t drow = side + 1024 + (uint32_t)(w * 64 + half * 32) * 4u;
      mbar_wait(&q_full[qs], (pi >> 1) & 1);
      const int tok_q = (int)lds_u32(side + (half * 64 + jj) * 4);  // dQ row (query jj of block `half`)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c < nch) {
          const int b = n & 1;
          const int j = c * 64 + jj;
          const int jb = j < HB_BIAS_ROWS - 1 ? j : HB_BIAS_ROWS - 1;   // rows past the tile read the -inf guard row
          const uint32_t brow = smem_u32(sBias) + (uint32_t)jb * 256u;
          const int sx = jb & 7;
          mbar_wait(&s_full[b], (n >> 1) & 1);
          tc_fence_after();
          // two passes of 16 queries (the three bias-gradient accumulator sets leave no room for 32-wide register tiles)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t st[16], dp[16];
            tmem_ld_32x16(t_lane + HC_ST + b * 64 + half * 32 + hh * 16, st);
            tmem_ld_32x16(t_lane + HC_DP + b * 64 + half * 32 + hh * 16, dp);
            tmem_ld_wait();
            uint32_t pp[8], dd[8];
            if (dbg & 4) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { pp[e] = st[e]; dd[e] = dp[e]; }
            } else
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int c8 = hh * 4 + c4;
              const float4 bb = lds_f4(brow + (uint32_t)(((half * 8 + c8) ^ sx) << 4));
              const float4 ll = lds_f4(lrow + c8 * 16);   // -lse * log2 e of the four queries
              const float4 dl = lds_f4(drow + c8 * 16);   // -delta
#pragma unroll
              for (int e2 = 0; e2 < 2; ++e2) {
                const uint64_t b2 = e2 ? ht_pk2(bb.z, bb.w) : ht_pk2(bb.x, bb.y);
                const uint64_t l2 = e2 ? ht_pk2(ll.z, ll.w) : ht_pk2(ll.x, ll.y);
                const uint64_t d2 = e2 ? ht_pk2(dl.z, dl.w) : ht_pk2(dl.x, dl.y);
                float x0, x1;
                ht_upk2(ht_fma2(ht_pk2u(st[c4 * 4 + e2 * 2], st[c4 * 4 + e2 * 2 + 1]), sl22, ht_add2(b2, l2)), x0, x1);
                const float p0 = wt_ex2(x0), p1 = wt_ex2(x1);
                const uint64_t sv = ht_mul2(ht_pk2(p0, p1), ht_add2(ht_pk2u(dp[c4 * 4 + e2 * 2], dp[c4 * 4 + e2 * 2 + 1]), d2));
                acc[c][c8 * 2 + e2] = ht_add2(acc[c][c8 * 2 + e2], sv);
                float s0, s1;
                ht_upk2(sv, s0, s1);
                pp[c4 * 2 + e2] = pack_bf16(p0, p1);
                dd[c4 * 2 + e2] = pack_bf16(s0, s1);
              }
            }
            // P^T / dS^T (bf16 pairs) back into TMEM over the columns just read; dS^T also into this tile's shared buffer.
            // Both were last read by the gradient MMAs of tile n - 2, which completed before scores(n) were issued.
            wt_tmem_st8(t_lane + HC_ST + b * 64 + half * 32 + hh * 8, pp);
            wt_tmem_st8(t_lane + HC_DP + b * 64 + half * 32 + hh * 8, dd);
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
              const uint32_t off = (uint32_t)(b * 16384) + sw128(r, half * 4 + hh * 2 + q2);
              sts_u4(smem_u32(sdS) + off, dd[q2 * 4], dd[q2 * 4 + 1], dd[q2 * 4 + 2], dd[q2 * 4 + 3]);
            }
          }
          wt_tmem_st_wait();
          tc_fence_before();
          wt_proxy_fence();
          wt_warp_arrive(&pds_full[b], lane);
          if (n > 0) epilogue(n - 1, prow, ptok_q, ptok_q != -2);
          prow = (j < g.nkv && grp < g.groups) ? (((long)grp * g.heads + h) * g.nkv + j) : -1;
          ptok_q = (c == nch - 1) ? tok_q : -2;  // -2: the pair is not finished, dQ keeps accumulating
          ++n;
        }
      }
    }
    if (n > 0) epilogue(n - 1, prow, ptok_q, ptok_q != -2);

    // bias gradient: registers -> per-CTA table in shared memory (over the retired P^T tile) -> global atomics
    if (has_tab) {
      float* dtab = reinterpret_cast<float*>(sP);
      const int mt = threadIdx.x - 128;  // 0..255
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int e = mt; e < p.n_pos; e += 256) dtab[e] = 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int j = c * 64 + jj;
        if (c < nch && j < g.nkv) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int i = half * 32 + e;
            float a0, a1;
            ht_upk2(acc[c][e >> 1], a0, a1);
            const float a = (e & 1) ? a1 : a0;
            if (i < g.nq && a != 0.f) atomicAdd(&dtab[__ldg(p.pos + i * g.nkv + j)], a);
          }
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

// dK / dV of a token = sum of the partial rows of the blocks whose halo covers it, in block order (fp32, rounded once)
__global__ void __launch_bounds__(256)
attn_ht_dkv_reduce_kernel(vtb_attn_params p, HtGeom g, const bf16* __restrict__ part_k, const bf16* __restrict__ part_v,
                          long total) {
  const int W = g.window, HL = g.halo, nby = g.Hs / W;
  const long T = (long)p.batch * g.Hs * g.Ws;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int cc = (int)(e & 3);
    long t2 = e >> 2;
    const int h = (int)(t2 % g.heads);
    t2 /= g.heads;
    const long tok = t2 % T;
    const int kv = (int)(t2 / T);
    const int b = (int)(tok / (g.Hs * g.Ws));
    const int rem = (int)(tok - (long)b * g.Hs * g.Ws);
    const int y = rem / g.Ws, x = rem - y * g.Ws;
    const int ay = y - W - HL + 1, ax = x - W - HL + 1;
    const int by_lo = ay > 0 ? (ay + W - 1) / W : 0, bx_lo = ax > 0 ? (ax + W - 1) / W : 0;
    int by_hi = (y + HL) / W, bx_hi = (x + HL) / W;
    if (by_hi > nby - 1) by_hi = nby - 1;
    if (bx_hi > g.nbx - 1) bx_hi = g.nbx - 1;
    const bf16* part = kv ? part_v : part_k;
    float s[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = 0.f;
    auto row_of = [&](int by, int bx) {
      const long grp = (long)b * g.nb + by * g.nbx + bx;
      const int j = (y - by * W + HL) * g.kw + (x - bx * W + HL);
      return reinterpret_cast<const uint4*>(part + ((grp * g.heads + h) * g.nkv + j) * 32 + cc * 8);
    };
    auto add = [&](const uint4& v) {
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16(wv[q]);
        s[q * 2] += f.x;
        s[q * 2 + 1] += f.y;
      }
    };
    if (by_hi - by_lo <= 1 && bx_hi - bx_lo <= 1) {
      // the usual case (halo <= window): at most 2 x 2 covering blocks — all loads in flight before the first add
      const bool y2 = by_hi > by_lo, x2 = bx_hi > bx_lo;
      const uint4 z = make_uint4(0, 0, 0, 0);
      const uint4 v00 = __ldg(row_of(by_lo, bx_lo));
      const uint4 v01 = x2 ? __ldg(row_of(by_lo, bx_hi)) : z;
      const uint4 v10 = y2 ? __ldg(row_of(by_hi, bx_lo)) : z;
      const uint4 v11 = (y2 && x2) ? __ldg(row_of(by_hi, bx_hi)) : z;
      add(v00); add(v01); add(v10); add(v11);
    } else {
      for (int by = by_lo; by <= by_hi; ++by)
        for (int bx = bx_lo; bx <= bx_hi; ++bx) add(__ldg(row_of(by, bx)));
    }
    if (!kv) {
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] *= p.scale;
    }
    bf16* dst = kv ? reinterpret_cast<bf16*>(p.dv) + tok * p.lddv : reinterpret_cast<bf16*>(p.dk) + tok * p.lddk;
    *reinterpret_cast<uint4*>(dst + h * 32 + cc * 8) =
        make_uint4(pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]), pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
  }
}

bool g_attn_ht = true;
int g_ht_dbg = 0;  // vtb_set_option("attn_ht_dbg", bits): timing experiments on the backward kernel (results are garbage):
                  // 1 no global stores, 2 no K/V copies, 4 no softmax math, 8 no gradient MMAs, 16 no score MMAs, 32 no reduce kernel

int ht_geom(const vtb_attn_params* p, HtGeom* g) {
  g->heads = p->heads; g->nq = p->nq; g->nkv = p->nkv; g->Hs = p->Hs; g->Ws = p->Ws;
  g->window = p->window; g->halo = p->halo; g->kw = p->window + 2 * p->halo;
  g->nbx = p->Ws / p->window;
  g->nb = (p->Hs / p->window) * g->nbx;
  const long groups = (long)p->batch * g->nb;
  VTB_CHECK(groups < (1L << 22) && (long)p->batch * p->Hs * p->Ws < (1L << 31), -1,
            "vtb_attention(halo tcgen05): problem too large for 32-bit token indices");
  g->groups = (int)groups;
  g->nkc = (p->nkv + 15) / 16;
  g->nch = (p->nkv + 63) / 64;
  g->inv_nb = 1.f / (float)g->nb; g->inv_nbx = 1.f / (float)g->nbx;
  g->inv_kw = 1.f / (float)g->kw; g->inv_w = 1.f / (float)g->window;
  return 0;
}

}  // namespace

void vtb_attn_ht_set(bool on) { g_attn_ht = on; }
void vtb_attn_ht_dbg_set(int bits) { g_ht_dbg = bits; }

size_t vtb_attn_ht_ws_bytes(const vtb_attn_params* p) {
  const long groups = (long)p->batch * (p->Hs / p->window) * (p->Ws / p->window);
  return (size_t)2 * groups * p->heads * p->nkv * 32 * sizeof(bf16);
}

bool vtb_attn_ht_ok(const vtb_attn_params* p, bool bwd) {
  if (!g_attn_ht || p->mode != VTB_ATTN_HALO || p->dh != 32 || p->nq > 64 || p->nkv > HT_MAXK || p->mask) return false;
  if (p->rel_bias && p->n_pos > 512) return false;
  if (p->heads > 148) return false;
  uintptr_t al = (uintptr_t)p->q | (uintptr_t)p->k | (uintptr_t)p->v | (uintptr_t)p->o;
  int ld = p->ldq | p->ldk | p->ldv | p->ldo;
  if (bwd) {
    if (p->dkv_f32 || !p->ws || p->ws_bytes < (int64_t)vtb_attn_ht_ws_bytes(p)) return false;
    al |= (uintptr_t)p->dout | (uintptr_t)p->dq | (uintptr_t)p->dk | (uintptr_t)p->dv | (uintptr_t)p->ws;
    ld |= p->lddo | p->lddq | p->lddk | p->lddv;
  }
  return (al & 15) == 0 && (ld & 7) == 0;
}

int vtb_attn_ht_fwd(const vtb_attn_params* p, cudaStream_t stream) {
  static bool attr = false;
  HtGeom g;
  if (int rc = ht_geom(p, &g)) return rc;
  using Kern = void (*)(vtb_attn_params, HtGeom, int, int);
  static const Kern kerns[11] = {attn_ht_fwd_kernel<1>, attn_ht_fwd_kernel<2>,  attn_ht_fwd_kernel<3>, attn_ht_fwd_kernel<4>,
                                 attn_ht_fwd_kernel<5>, attn_ht_fwd_kernel<6>,  attn_ht_fwd_kernel<7>, attn_ht_fwd_kernel<8>,
                                 attn_ht_fwd_kernel<9>, attn_ht_fwd_kernel<10>, attn_ht_fwd_kernel<11>};
  if (!attr) {
    for (int i = 0; i < 11; ++i)
      VTB_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, HF_SMEM));
    attr = true;
  }
  const int ntiles = (g.groups + 1) / 2;
  int nchunks = vtb_num_sms() / p->heads;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > ntiles) nchunks = ntiles;
  kerns[g.nkc - 1]<<<(unsigned)(nchunks * p->heads), WT_THREADS, HF_SMEM, stream>>>(*p, g, ntiles, nchunks);
  VTB_LAUNCH_CHECK();
  return 0;
}

int vtb_attn_ht_bwd(const vtb_attn_params* p, cudaStream_t stream) {
  static bool attr = false;
  HtGeom g;
  if (int rc = ht_geom(p, &g)) return rc;
  if (!attr) {
    VTB_CUDA(cudaFuncSetAttribute(attn_ht_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HB_SMEM));
    attr = true;
  }
  const int npairs = (g.groups + 1) / 2;
  int nchunks = vtb_num_sms() / p->heads;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > npairs) nchunks = npairs;
  bf16* part_k = reinterpret_cast<bf16*>(p->ws);
  bf16* part_v = part_k + (size_t)g.groups * p->heads * p->nkv * 32;
  attn_ht_bwd_kernel<<<(unsigned)(nchunks * p->heads), WT_THREADS, HB_SMEM, stream>>>(*p, g, npairs, nchunks, part_k, part_v, g_ht_dbg);
  VTB_LAUNCH_CHECK();
  if (g_ht_dbg & 32) return 0;
  const long total = 2L * p->batch * p->Hs * p->Ws * p->heads * 4;
  long blocks = (total + 255) / 256;
  const long cap = (long)vtb_num_sms() * 32;
  if (blocks > cap) blocks = cap;
  attn_ht_dkv_reduce_kernel<<<(unsigned)blocks, 256, 0, stream>>>(*p, g, part_k, part_v, total);
  VTB_LAUNCH_CHECK();
  return 0;
}
