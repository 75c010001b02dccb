/* vtb200.h — C-ABI of libvtb200.so: the sm_100a kernels behind the transformer-block hot path of
 * rosinality/vision-transformers-pytorch (models/{vit,swin_transformer,pvt,halo_transformer,twins}.py).
 *
 * The reference has NO native interface: every op below replaces an ATen call issued from the
 * reference's nn.Module.forward (and the autograd backward of it).  Each entry point cites the
 * reference call site it stands in for (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory owned by the caller (the library never
 *     allocates or frees device memory); all work is enqueued on `stream`, nothing synchronises.
 *   - return 0 on success, negative on error; vtb_last_error() gives the message (thread-local).
 *   - "bf16" pointers are `const void*` to raw 16-bit brain-floats; "f32" are float*.
 *   - re-entrant: no global mutable state except the resolved driver entry point (vtb_init).
 */
#ifndef VTB200_H
#define VTB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vtb_stream_t; /* == cudaStream_t */

const char* vtb_last_error(void);
int vtb_version(void);
/* Resolve cuTensorMapEncodeTiled through the runtime, query SM count.  Idempotent. */
int vtb_init(void);
/* Runtime switches (A/B measurements, cross-checks; also settable from the environment: VTB_OPTS=name=value,... is applied by the
 * Python loader).  Defaults are the measured-best settings.
 * "gemm_cluster" (default 1): GEMM tiles as CTA pairs (tcgen05 cta_group::2, 256 x BN tiles, each CTA stages its 128 rows of A and
 *   half of B) wherever legal; 0 = 1-CTA tiles; 3 = the round-1 rule (pairs only when pairs x splits fill the pair slots).
 * "gemm_colsum_pair" (default 1): a_colsum launches may use CTA pairs; 0 = 1-CTA tiles for them.
 * "gemm_helpers" (1/2, default 2): warps that issue the staged epilogue's TMA stores.
 * "gemm_bn_waste_pct" (default 120): a wider tile-N may waste this much more of the MMA columns than the next narrower one.
 * "attn_tc" (0/1, default 1): tcgen05/TMEM attention kernels where they apply (global attention, dh = 64, <= 256 keys);
 *   0 forces the mma.sync kernels.  "attn_tc_fwd_version" / "attn_tc_bwd_version" (1/2, default 2): persistent round-2 kernels or
 *   the round-1 ones.
 * "attn_wp" (0/1, default 1): one warp per (window, head) problem when nq, nkv <= 64 on the mma.sync path; 0 = one CTA per problem.
 * "attn_wt" (0/1, default 1): tcgen05 window kernels (two windows per 128-row tile) for WINDOW problems with dh = 32 and
 *   <= 64 tokens per window; 0 forces the mma.sync kernels.
 * "attn_ht" (0/1, default 1): tcgen05 halo kernels (dh = 32, <= 64 queries, <= 176 halo slots); 0 forces the mma.sync kernels.
 *   "attn_ht_dbg": timing knock-outs of the halo backward (results are wrong on purpose).
 * "ln_stream" (0/1, default 1): streaming (bulk-copy staged) LayerNorm kernels; 0 = register-resident kernels.
 * "input_variant" (1/2/3, default 2): vtb_input_batch kernel; 2 = the leaner variant (bit-identical, 5-17 % faster on the B200),
 *   3 = the same with 8 pixels per thread when W % 8 == 0; both need a 16-byte aligned table, else variant 1 runs. */
int vtb_set_option(const char* name, int32_t value);

/* ------------------------------------------------------------------------------------------------
 * GEMM on tcgen05 (TMA -> smem -> UMMA -> TMEM -> epilogue):  C[M,N] = alpha * sum_k A(m,k) B(n,k)
 * Replaces every nn.Linear / k=s Conv2d-as-GEMM forward, dgrad and wgrad on the path:
 *   layer.py:191-196 (FFN), vit.py:30,43 / swin:128,155 / pvt:40,52,67 / halo:63,107 / twins:66,77,91,130,150
 *   (attention projections), vit.py:73,76 / pvt.py:111 / pvt.py:27 / twins.py:54 (patch & reduce convs),
 *   swin:209,226 / halo:162 / twins:209 (patchify linears), vit.py:200 / swin:377 / pvt:278 / halo:219,278 (heads).
 * Operand layouts (row-major storage, ld in elements, bf16):
 *   a_mn_major=0: A stored [M][K] (K contiguous)      a_mn_major=1: A stored [K][M] (M contiguous)
 *   b_mn_major=0: B stored [N][K] (K contiguous)      b_mn_major=1: B stored [K][N] (N contiguous)
 *   forward  y = x W^T : A=x (0), B=W (0)     dgrad dx = dy W : A=dy (0), B=W (1)
 *   wgrad dW = dy^T x  : A=dy (1), B=x (1)
 * Epilogue, applied in this order to v = alpha*acc:
 *   v += bias[n]                                   (bias != NULL)
 *   VTB_EPI_SILU_DUAL : out  = bf16(v) ; v = silu(float(bf16(v))) ; stored to out2   (layer.py:193)
 *   VTB_EPI_SILU_GRAD : v *= silu'(aux[m,n])                                       (backward of it)
 *   v *= row_scale[m / rows_per_scale]             (DropPath keep-mask/keep-prob, layer.py:176-178)
 *   v += resid[m*ldr + n]                          (residual add, vit.py:60-61 etc.; EPI_NONE, f32 out only)
 *   out[m*ldo + n] = v (bf16 or f32); with accumulate=1: += (f32 only; split-K / grad accumulation,
 *                     TMA reduce-add when rows are 16-byte aligned, atomicAdd otherwise)
 * When out/out2/resid/aux rows are 16-byte aligned the epilogue is staged through shared memory and written
 * with TMA bulk tensor stores (residual / aux tiles TMA-prefetched); otherwise a direct per-thread path runs.
 * ---------------------------------------------------------------------------------------------- */
enum { VTB_EPI_NONE = 0, VTB_EPI_SILU_DUAL = 1, VTB_EPI_SILU_GRAD = 2 };

typedef struct {
  int32_t M, N, K;
  const void* A; int32_t lda; int32_t a_mn_major;
  const void* B; int32_t ldb; int32_t b_mn_major;
  void* out; int32_t ldo; int32_t out_f32;
  void* out2;                         /* bf16, same ld as out (SILU_DUAL only) */
  const float* bias;                  /* [N] or NULL */
  const float* resid; int32_t ldr;    /* f32 or NULL */
  const float* row_scale; int32_t rows_per_scale;
  const void* aux; int32_t ldaux;     /* bf16 pre-activation (SILU_GRAD) */
  int32_t epilogue;
  int32_t splits;                     /* split-K factor; 0 = auto; >1 requires accumulate=1 */
  int32_t accumulate;
  float alpha;
  /* optional, a_mn_major=1 only (weight gradients dW = dy^T x): a_colsum[m] += sum_k A(m, k), i.e. the column sums
   * of the dy operand = the bias gradient of the same Linear (layer.py:191 / vit.py:30 ...), accumulated by the
   * otherwise idle epilogue warps from the operand tiles already staged in shared memory, so the bias gradient
   * costs no extra pass over dy.  f32 [M], atomicAdd; needs EPI_NONE without resid / out2. */
  float* a_colsum;
} vtb_gemm_params;

int vtb_gemm_bf16(const vtb_gemm_params* p, vtb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm (nn.LayerNorm: vit.py:13,52,54,113 / swin:12,180,185,206,221,277 / pvt:9,31,86,89,112,202 /
 * halo:9,133,138,159,215,217 / twins:12,166-178,206,258).  fp32 statistics, affine.
 * Input rows are fp32 (the residual stream).  Optional patchify gather (swin:15-22, A4 in SURVEY):
 *   patch_s > 1: x is NHWC [B, Hin, Win, C]; row r = (b, by, bx) gathers features (sy*s+sx)*C + c,
 *   cols must equal s*s*C.   patch_s <= 1: x is [rows, cols].
 * y is bf16 (y_f32=0) or f32.  mean/rstd [rows] are saved for backward.
 * rowmod_add (optional, f32 [group_rows, cols]) is added AFTER normalisation by (r % group_rows)  (pvt.py:140).
 * ---------------------------------------------------------------------------------------------- */
int vtb_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps,
                      int64_t rows, int32_t cols, int32_t patch_s, int32_t Hin, int32_t Win,
                      void* y, int32_t y_f32, float* mean, float* rstd,
                      const float* rowmod_add, int32_t group_rows, vtb_stream_t stream);

/* Backward of the above.  dy is bf16 (dy_f32=0) or f32 [rows, cols].
 *   dx_row = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat))
 *   dx_out[r] = (dx_in ? dx_in[r] : 0) + dx_row          (f32; patchify mode scatters back to NHWC)
 *   if dx_bf16 != NULL: dx_bf16[r] = bf16(dx_out[r] * (row_scale ? row_scale[r / rows_per_scale] : 1))
 *   partial dgamma/dbeta are accumulated into dgamma/dbeta (f32 [cols], atomicAdd; caller zeroes or
 *   passes the .grad buffer to accumulate into).
 *   if dx_colsum != NULL (needs dx_bf16, dense rows, cols <= 768): dx_colsum[c] += sum_r float(dx_bf16[r, c]) — the
 *   bias gradient of the Linear that produced this residual stream (its DropPath scale folded in through
 *   row_scale), so the caller's next backward step needs no separate cast / column-sum pass over dx_out.
 */
int vtb_layernorm_bwd(const void* dy, int32_t dy_f32, const float* x, const float* gamma,
                      const float* mean, const float* rstd, int64_t rows, int32_t cols,
                      int32_t patch_s, int32_t Hin, int32_t Win, const float* dx_in, float* dx_out,
                      void* dx_bf16, const float* row_scale, int32_t rows_per_scale,
                      float* dgamma, float* dbeta, float* dx_colsum, vtb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-head attention, all four variants of the reference through one geometry descriptor:
 *   VTB_ATTN_GLOBAL  vit.py:27-45, pvt.py:32-69, twins.py:58-93   group g = image b
 *   VTB_ATTN_WINDOW  swin:103-160 (shift/bias/mask), twins.py:109-152   g = (b, window)
 *   VTB_ATTN_HALO    halo:57-114   g = (b, block); keys gathered from a (W+2h)^2 zero-padded halo
 * Q,K,V,O live in token-major buffers: element (token t, head h, d) = base[t*ld + h*dh + d].
 * S = scale * Q K^T + rel_bias[pos[i,j], h] ; S = -inf where mask[g % n_mask, i, j] ; P = softmax(S) ; O = P V.
 * Scores and probabilities never leave the SM.  lse [groups, H, Nq] (log-sum-exp) is saved for backward.
 * ---------------------------------------------------------------------------------------------- */
enum { VTB_ATTN_GLOBAL = 0, VTB_ATTN_WINDOW = 1, VTB_ATTN_HALO = 2 };

typedef struct {
  int32_t mode;
  int32_t batch, heads, dh;
  int32_t nq, nkv;             /* queries / key slots per group */
  int32_t Hs, Ws;              /* spatial map (WINDOW/HALO) */
  int32_t window, shift, halo; /* shift = floor(window/2) on shifted layers else 0 */
  float scale;                 /* 1/sqrt(dh), applied after QK^T (vit.py:37) */
  const void* q; int32_t ldq;
  const void* k; int32_t ldk;
  const void* v; int32_t ldv;
  void* o; int32_t ldo;
  float* lse;
  const float* rel_bias;       /* [n_pos, heads] f32 (rel_pos.weight) or NULL */
  const int32_t* pos;          /* [nq, nkv] or NULL */
  int32_t n_pos;               /* rows of rel_bias (swin 169, halo 253; <= 512) */
  const uint8_t* mask;         /* [n_mask, nq, mask_ld], 1 = masked (swin local_mask) or NULL */
  int32_t n_mask;
  int32_t mask_ld;             /* row pitch of mask in bytes; 0 = nkv.  64 lets the window kernels fetch rows with 16-byte loads */
  /* backward only */
  const void* dout; int32_t lddo;
  void* dq; int32_t lddq;      /* bf16 */
  void* dk; int32_t lddk;      /* bf16, or f32 atomics when dkv_f32 (HALO: tokens shared by blocks) */
  void* dv; int32_t lddv;
  int32_t dkv_f32;
  float* delta;                /* workspace [groups, H, Nq] */
  float* drel_bias;            /* [n_pos, heads] f32, atomicAdd */
  /* optional, lets WINDOW problems with a mask take the tcgen05 window kernels: the mask as one 64-bit word per row,
   * [n_mask][2][64] uint64: word [m][0][i] has bit j set where mask[m, i, j] != 0 (query rows, forward),
   * word [m][1][j] has bit i set where mask[m, i, j] != 0 (key rows, backward).  NULL with mask != NULL -> mma.sync path. */
  const void* mask_bits;
  /* optional, backward only: scratch for the tcgen05 HALO kernels — per-block partial dK / dV rows that a second kernel sums
   * per token in a fixed order (no atomics).  Size from vtb_attention_bwd_workspace_bytes(); NULL / too small -> the
   * mma.sync halo kernels run instead.  16-byte aligned. */
  void* ws;
  int64_t ws_bytes;
} vtb_attn_params;

/* Bytes of vtb_attn_params.ws that vtb_attention_bwd can use for this geometry (0: none needed). */
int64_t vtb_attention_bwd_workspace_bytes(const vtb_attn_params* p);

int vtb_attention_fwd(const vtb_attn_params* p, vtb_stream_t stream);
int vtb_attention_bwd(const vtb_attn_params* p, vtb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Validation mode (vtb200.ops.validation_mode(), forward only): the same host path and the same tcgen05 GEMM / LayerNorm
 * kernels, fed so that a whole-model forward meets the north star's "logits within rtol 1e-3 of the reference's fp32
 * forward", which bf16 operands cannot (the reference's own autocast run does not either: tests/golden/autocast_levels.json).
 *   - activations and weights stay fp32 between kernels (vtb_layernorm_fwd with y_f32, fp32 GEMM outputs);
 *   - vtb_split3_bf16 turns an fp32 operand [rows, K] into bf16 [rows, 3K] = [hi | lo | hi] (A side) or [hi | hi | lo]
 *     (B side), hi = bf16(x), lo = bf16(x - hi): one vtb_gemm_bf16 call with K' = 3K then accumulates
 *     a_hi b_hi + a_lo b_hi + a_hi b_lo in fp32 (error ~2^-16 relative per product instead of 2^-8);
 *   - vtb_attention_fwd_f32: exact-softmax attention on the CUDA cores for every geometry of vtb_attn_params;
 *     q / k / v / o are FLOAT buffers (leading dimensions in elements), lse optional;
 *   - vtb_silu_fwd_exact: x / (1 + exp(-x)) without the tanh.approx shortcut; vtb_patch_gather_f32: the patch gather of
 *     vtb_patch_gather with a float destination.
 * ---------------------------------------------------------------------------------------------- */
int vtb_split3_bf16(const float* src, int64_t lds, int64_t rows, int32_t cols, int32_t b_side, void* dst,
                    vtb_stream_t stream);
int vtb_attention_fwd_f32(const vtb_attn_params* p, vtb_stream_t stream);
int vtb_silu_fwd_exact(const float* x, float* y, int64_t n, vtb_stream_t stream);
int vtb_patch_gather_f32(const float* src, int32_t src_nchw, int32_t c_major, int32_t B, int32_t C, int32_t H,
                         int32_t W, int32_t p, float* dst, vtb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Small memory-bound helpers.
 * ---------------------------------------------------------------------------------------------- */
/* fp32 -> bf16 cast (autocast's weight cast, torch.cuda.amp.autocast at train.py:273). */
int vtb_cast_f32_bf16(const float* src, void* dst, int64_t n, vtb_stream_t stream);
/* bf16 <- f32 with row stride: dst[r, c] = bf16(src[r, c]) for strided views. */
int vtb_cast_f32_bf16_2d(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows,
                         int32_t cols, vtb_stream_t stream);
/* dst[r, c] = bf16(src[r, c] * row_scale[r / rows_per_scale])  (row_scale NULL => 1): the DropPath-scaled
 * bf16 copy of the residual-stream gradient that feeds the branch's dgrad/wgrad GEMMs (layer.py:178 adjoint). */
int vtb_scale_cast_bf16(const float* src, const float* row_scale, int32_t rows_per_scale, int64_t rows,
                        int32_t cols, void* dst, vtb_stream_t stream);
/* Same, and colsum[c] += sum_r dst[r, c] in the same pass (the bias gradient of the Linear that closes a branch:
 * the scaled gradient IS that Linear's output gradient).  colsum f32 [cols], accumulated with atomicAdd. */
int vtb_scale_cast_colsum_bf16(const float* src, const float* row_scale, int32_t rows_per_scale, int64_t rows,
                               int32_t cols, void* dst, float* colsum, vtb_stream_t stream);
/* Element dropout with a caller-drawn keep mask (uint8, one byte per element; the host draws it with torch's generator in the
 * reference's call order so that a shared seed gives the reference's masks):
 *   out[i] = resid[i] + row_scale[i / elems_per_scale] * (keep[i] ? x[i] * scale : 0)      (resid / row_scale NULL => 0 / 1)
 * x / out are f32 (is_f32) or bf16; out may alias x.  Replaces nn.Dropout in PositionwiseFeedForward (layer.py:194), on the
 * branch outputs and the token embedding of ViT (vit.py:57-61,102,146) and on PVT's patch embedding (pvt.py:127,141); the
 * residual form is the DropPath + residual step of a ViT branch (vit.py:60-61).  The adjoint is the same call on the gradient. */
int vtb_dropout(const void* x, const uint8_t* keep, float scale, int64_t n, int32_t is_f32, const float* resid,
                const float* row_scale, int64_t elems_per_scale, void* out, vtb_stream_t stream);
/* SiLU on f32 (halo_transformer.py:218) and its adjoint. */
int vtb_silu_fwd(const float* x, float* y, int64_t n, vtb_stream_t stream);
int vtb_silu_bwd(const float* x, const float* dy, float* dx, int64_t n, vtb_stream_t stream);
/* out[n] += sum_m X[m, n]  (bias gradients).  X bf16 [M, ld]. */
int vtb_colsum_bf16(const void* X, int64_t M, int32_t N, int32_t ld, float* out,
                    vtb_stream_t stream);
/* Patch gather to a bf16 GEMM operand (conv k=s stride=s as GEMM):
 *   src_nchw=1: src[b, c, y, x] (f32)            src_nchw=0: src[b, y, x, c] (f32 or bf16 by src_bf16)
 *   c_major=1 : feature = c*p*p + py*p + px (vit.py:73, pvt.py:111)
 *   c_major=0 : feature = (py*p + px)*C + c (patchify, swin:15-22)
 * dst bf16 [B*(H/p)*(W/p), p*p*C]. */
int vtb_patch_gather(const void* src, int32_t src_bf16, int32_t src_nchw, int32_t c_major,
                     int32_t B, int32_t C, int32_t H, int32_t W, int32_t p, void* dst,
                     vtb_stream_t stream);
/* Adjoint of vtb_patch_gather: dx (+)= dA[row, feat]  (dA bf16 or f32; dx f32, NHWC [b,y,x,c] or, with
 * dst_nchw=1, NCHW [b,c,y,x]).  accumulate=0 overwrites. */
int vtb_patch_scatter(const void* dA, int32_t dA_f32, int32_t c_major, int32_t B, int32_t C,
                      int32_t H, int32_t W, int32_t p, float* dx, int32_t accumulate, int32_t dst_nchw,
                      vtb_stream_t stream);
/* dst[b, w, h, :] = src[b, h, w, :]  (bf16 or f32 rows of C elements).  The Twins global attention feeds its
 * reduce-conv with `input.transpose(1, 2).reshape(B, C, H, W)` (twins.py:70): a transposed copy that is then
 * REINTERPRETED as NCHW — this kernel makes the copy, vtb_patch_gather(src_nchw=1) does the reinterpretation. */
int vtb_transpose_hw(const void* src, void* dst, int32_t is_f32, int32_t B, int32_t H, int32_t W, int32_t C,
                     vtb_stream_t stream);
/* Depthwise 3x3 conv (padding 1, no bias) + identity on NHWC f32: y = dwconv(x, w) + x  (twins.py:25-36,
 * PositionalEncodingGenerator; w is [C,1,3,3]) and its adjoints: dx = dwconv^T(dy, w) + dy, dw += sum dy*x. */
int vtb_dwconv3x3_fwd(const float* x, const float* w, int32_t B, int32_t H, int32_t W, int32_t C, float* y,
                      vtb_stream_t stream);
int vtb_dwconv3x3_bwd(const float* x, const float* w, const float* dy, int32_t B, int32_t H, int32_t W,
                      int32_t C, float* dx, float* dw, vtb_stream_t stream);
/* ViT token assembly (vit.py:141-143): x[b, 0, :] = cls + pos[0];  x[b, 1+p, :] = tok[b*n + p, :] + pos[1+p, :]
 * tok f32 [B*n, D] (patch GEMM output incl. conv bias), pos f32 [n+1, D], x f32 [B, n+1, D]. */
int vtb_vit_assemble_tokens(const float* tok, const float* cls, const float* pos, int32_t B, int32_t n,
                            int32_t D, float* x, vtb_stream_t stream);
/* x[g*row_stride_groups + c] = a[c] + b[c] for g < groups  (cls rows: pvt.py:136-140) */
int vtb_fill_rows(float* x, int64_t row_stride_groups, int32_t groups, int32_t cols,
                  const float* a, const float* b, vtb_stream_t stream);
/* out[c] += sum_g x[g*group_stride, c]  (cls_token / pos gradients), f32 */
int vtb_rowgroup_sum(const float* x, int64_t group_stride, int32_t groups, int32_t rows,
                     int32_t cols, float* out, vtb_stream_t stream);
/* Mean over `n` consecutive rows: out[g, :] = mean_r x[g*n + r, :]  (AdaptiveAvgPool2d(1), swin:281) and
 * its adjoint dx[g*n + r, :] = dy[g, :]/n. */
int vtb_mean_rows_fwd(const float* x, int32_t groups, int32_t n, int32_t cols, float* out,
                      vtb_stream_t stream);
int vtb_mean_rows_bwd(const float* dy, int32_t groups, int32_t n, int32_t cols, float* dx,
                      vtb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused DINO loss (loss.py:119-142, DINOLoss.forward): one launch computes
 *   loss += mean over pairs (iq, v != iq) and images of  -sum_k softmax((teacher[iq]-center)/t_teacher)[k]
 *                                                          * log_softmax(student[v]/t_student)[k]
 * and (dstudent != NULL) the gradient of that loss w.r.t. student.  student f32 [n_crops*batch, dim] (crop-major chunks,
 * the first two crops are the global ones), teacher f32 [2*batch, dim], center f32 [dim], loss f32 [1] (atomicAdd: zero it),
 * dstudent f32 [n_crops*batch, dim].  One CTA per image, two streaming passes over its rows.
 * ---------------------------------------------------------------------------------------------- */
int vtb_dino_loss(const float* student, const float* teacher, const float* center, int32_t n_crops, int32_t batch,
                  int32_t dim, float t_student, float t_teacher, float* loss, float* dstudent, vtb_stream_t stream);

/* DINO projection head (vit.py:206-262).  One warp per row.
 *   l2norm:      y = bf16(x / max(||x||_2, eps))  (F.normalize, vit.py:259; eps 1e-12), inv[r] = 1 / max(||x_r||, eps)
 *                dx = inv (dy - y (y . dy))
 *   weight_norm: w = bf16(v * g / ||v||_row)       (nn.utils.weight_norm of the last Linear, vit.py:244-248; g is [rows])
 *                dg = (dW . v) inv;  dv = g inv (dW - v (dW . v) inv^2)          (dg may be NULL: norm_last_layer)
 * Both emit the bf16 GEMM operand directly; x, v, dy, dW, dx, dv f32 [rows, cols] contiguous.
 *   gelu:        y = x Phi(x) (exact erf form, nn.GELU() vit.py:228,236), f32 and/or bf16 output; dx = dy (Phi + x phi). */
int vtb_l2norm_fwd(const float* x, int64_t rows, int32_t cols, float eps, void* y_bf16, float* inv, vtb_stream_t stream);
int vtb_l2norm_bwd(const float* dy, const float* x, const float* inv, int64_t rows, int32_t cols, float* dx,
                   vtb_stream_t stream);
int vtb_weight_norm_fwd(const float* v, const float* g, int64_t rows, int32_t cols, void* w_bf16, float* inv,
                        vtb_stream_t stream);
int vtb_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* inv, int64_t rows, int32_t cols,
                        float* dv, float* dg, vtb_stream_t stream);
int vtb_gelu_fwd(const float* x, float* y, void* y_bf16, int64_t n, vtb_stream_t stream);
int vtb_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, vtb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-tensor step-side kernels: the per-parameter loops of the training step restated as ONE
 * launch over a tensor list.  A list is given as HOST arrays of DEVICE pointers + element counts; the library packs
 * them into kernel parameters (no device-side table, nothing allocated) and splits lists longer than
 * VTB_MT_MAX_TENSORS into several launches.  All tensors f32 contiguous unless stated.  Scalars that the reference
 * holds on the device (clip coefficient) stay on the device: no host synchronisation anywhere.
 * ---------------------------------------------------------------------------------------------- */
#define VTB_MT_MAX_TENSORS 256
#define VTB_MT_CHUNK 8192 /* elements one CTA streams */
/* Number of CTAs (= partial sums) vtb_mt_sumsq uses for this list: sizes its `partials` workspace. */
int64_t vtb_mt_num_chunks(const int64_t* numel, int32_t n);
/* dst[i] = bf16(src[i]) for every tensor of the list (autocast's per-weight casts, train.py:273, in one launch). */
int vtb_mt_cast_f32_bf16(const void* const* src, void* const* dst, const int64_t* numel, int32_t n,
                         vtb_stream_t stream);
/* dst = dst*decay + src*(1-decay): train_util.py:70-84 (`accumulate`) and the DINO teacher update train_dino.py:257-261. */
int vtb_mt_ema(void* const* dst, const void* const* src, const int64_t* numel, int32_t n, double decay,
               vtb_stream_t stream);
/* Global gradient norm + clip coefficient (torch.nn.utils.clip_grad_norm_, train.py:294 / train_dino.py:243):
 *   out[0] = sqrt(sum of squares of every element), out[1] = min(1, max_norm / (out[0] + 1e-6)).
 * partials: f32 workspace of vtb_mt_num_chunks() floats.  Deterministic (fixed summation order). */
int vtb_mt_grad_norm(const void* const* grad, const int64_t* numel, int32_t n, float max_norm, float* partials,
                     float* out, vtb_stream_t stream);
/* x *= *scale (device scalar; the clip coefficient of vtb_mt_grad_norm).  Tensors are left untouched when *scale == 1. */
int vtb_mt_scale(void* const* x, const int64_t* numel, int32_t n, const float* scale, vtb_stream_t stream);
/* Adaptive gradient clipping, optimizer.py:12-26: per unit (row 0-slice of an ndim>1 tensor, the whole tensor otherwise)
 *   max_norm = max(||p_unit||, eps) * clipping;  g_unit *= max_norm / max(||g_unit||, 1e-6)  if ||g_unit|| >= max_norm.
 * units[i] = number of units of tensor i (shape[0] or 1); numel[i] % units[i] == 0. */
int vtb_mt_agc(const void* const* param, void* const* grad, const int64_t* numel, const int64_t* units, int32_t n,
               float clipping, float eps, vtb_stream_t stream);
/* One AdamW step (torch.optim.AdamW, the `adamw` optimizer of config/swin-transformer-s.conf:39-42 and
 * config/dino_deit-s-16.conf:52-55) over a parameter group:
 *   g' = g * (*grad_scale)            (grad_scale NULL => 1; the clip coefficient, folding clip_grad_norm_'s pass in)
 *   p *= 1 - lr*weight_decay;  m = m + (1-beta1)(g' - m);  v = beta2 v + (1-beta2) g'^2
 *   p -= (lr / (1-beta1^step)) * m / (sqrt(v)/sqrt(1-beta2^step) + eps)
 * p_bf16 (may be NULL, entries may be NULL): bf16 copy of the updated parameter written in the same pass (the operand
 * copy the next forward would otherwise cast). */
int vtb_mt_adamw(void* const* param, const void* const* grad, void* const* exp_avg, void* const* exp_avg_sq,
                 void* const* p_bf16, const int64_t* numel, int32_t n, double lr, double beta1, double beta2,
                 double eps, double weight_decay, int64_t step, const float* grad_scale, vtb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused MixLoss + accuracy (loss.py:53-86, train_util.py:53-67; the loss/metric of train.py:273-281):
 *   logp = log_softmax(logits);  t = inter*T(target1) + (1-inter)*T(target2),  T(c) = eps/n + (1-eps)[k == c]
 *   loss[0] += sum_rows sum_k t_k (log t_k - logp_k) * loss_scale          (KL, 0 log 0 = 0; loss_scale = 1/B for "mean")
 *   dlogits = (softmax(logits) - t) * loss_scale                           (dlogits may be NULL)
 *   row_loss[r] = sum_k t_k (log t_k - logp_k)                              (reduction "none"; row_loss may be NULL)
 *   correct[0] += #rows whose target1 logit is the maximum, correct[1] += #rows with it among the `topk` largest
 *   (correct NULL ok; with loss, row_loss and dlogits all NULL the call is `accuracy` alone)
 * logits f32 [rows, n_class] with row stride ld, targets int64 (target2 NULL => target1), inter f32 [rows] (NULL => 1),
 * loss f32 [1] (atomicAdd: zero it; may be NULL), correct int32 [2].
 * A row whose target1 / target2 lies outside [0, n_class) is skipped (no loss, zero gradient, no hit, nothing read out of
 * bounds): MixLoss defines no ignore_index, and with reduction "mean" the divisor stays the full row count.  A row whose
 * target logit is NaN counts as a miss.
 * ---------------------------------------------------------------------------------------------- */
int vtb_mix_loss(const float* logits, int64_t ld, const int64_t* target1, const int64_t* target2, const float* inter,
                 int32_t rows, int32_t n_class, double eps, float loss_scale, float* loss, float* row_loss,
                 float* dlogits, int32_t* correct, int32_t topk, vtb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Device input path: uint8 HWC source images -> normalised f32 NCHW batch, with the per-sample
 * mixup / cutmix / RandomErasing of the reference's loader applied in the same pass.  Stands in for
 *   mix_dataset.py:37-90 (MixDataset.__getitem__: Image.blend :66 / paste :84 on uint8, mul + add_ :63 / slice copy :79
 *   on tensors), factory.py:163-174 (ToTensor + Normalize) and transforms.py:377-407 (RandomErasing._erase, modes
 *   "pixel" and "const"; factory.py:178-182).
 * src   u8  [n_src, H, W, 3] device;  out f32 [batch, 3, H, W] device;  mean3 / std3: three HOST floats each.
 * table i32 [batch, 24] device, one row per output image (the random decisions, drawn on the host in the reference's
 * order by device_input.MixSampler):
 *   0 src1   1 src2 (partner)   2 mode (0 none, 1 mixup, 2 cutmix)   3 domain (0 = mix in uint8 like PIL, then
 *   normalise, then erase box A on the result: mix_before_aug = true;  1 = normalise + erase each source (box A ->
 *   src1, box B -> src2), then mix the tensors: mix_before_aug = false)
 *   4 w1 (f32 bits): domain 0 -> PIL blend alpha = (float)(1 - ratio);  domain 1 -> (float)ratio
 *   5 w2 (f32 bits): domain 1 -> (float)(1 - ratio), the `alpha` of add_;  unused otherwise
 *   6..9 cutmix box x1, y1, x2, y2 (half-open)    10..13 erase box A top, left, h, w (h = 0: none)    14..17 erase box B
 *   18, 19 noise seeds of boxes A / B    20 erase mode (0 zeros, 1 N(0,1) per pixel: Philox4x32-10, counter (x, y, 0, 0),
 *   key (seed, 0x7674B200), Box-Muller -- restated bit-for-bit in oracle/input_ops.py)    21..23 reserved (0)
 * Results: bit-identical to torchvision/PIL for modes none / cutmix and the uint8 blend; tensor mixup within 1 ulp
 * (ATen's add_ may or may not contract to an FMA); erase noise is N(0,1) but not torch's CPU stream.
 * ---------------------------------------------------------------------------------------------- */
int vtb_input_batch(const uint8_t* src, int32_t n_src, const int32_t* table, int32_t batch, int32_t H, int32_t W,
                    const float* mean3, const float* std3, float* out, vtb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VTB200_H */
