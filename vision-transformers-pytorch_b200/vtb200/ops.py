"""Tensor-level wrappers over the C-ABI (include/vtb200.h).

torch is used only for device memory, the current stream and dtype bookkeeping; every FLOP and every
byte moved on the hot path happens inside libvtb200.so.  LAUNCHES counts the library calls (the
`gpu_launches` claim in bench.py).
"""
import ctypes as C
import math
import threading

import torch

from . import lib as _l

LAUNCHES = 0
# When set to a list, every op appends (name, flops, start_event, end_event): bench.py's live roofline pass
# and per-kernel time breakdown (CUDA events on the launching stream).
PROFILE = None


class _Prof:
    __slots__ = ("name", "flops", "nbytes", "e0")

    def __init__(self, name, flops=0.0, nbytes=0.0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e0.record()

    def __exit__(self, *exc):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((self.name, self.flops, self.e0, e1, self.nbytes))
        return False


class _NoProf:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NOPROF = _NoProf()


def _prof(name, flops=0.0, nbytes=0.0):
    """flops / nbytes: ALGORITHMIC work of the call (operands and outputs touched once), for the live roofline pass."""
    return _NOPROF if PROFILE is None else _Prof(name, flops, nbytes)

BF16 = torch.bfloat16
F32 = torch.float32


class _ZeroArena(threading.local):
    """Zero-initialised outputs (split-K weight gradients, bias / LayerNorm-parameter / table gradients: ~150 per ViT-B
    step, ~330 per Swin-S step) are carved out of 64 MiB chunks that are cleared with ONE memset each instead of one
    fill kernel per tensor.  A chunk stays alive for as long as any tensor carved from it does."""
    CHUNK = 64 << 20
    buf = None
    off = 0
    capturing = False  # a chunk allocated while a CUDA graph is being captured belongs to that graph's pool

    def take(self, shape, dtype, device):
        numel = 1
        for d in shape:
            numel *= int(d)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        if nbytes == 0 or nbytes > self.CHUNK // 4:
            return torch.zeros(shape, dtype=dtype, device=device)
        need = (nbytes + 255) // 256 * 256
        cap = torch.cuda.is_current_stream_capturing()
        if (self.buf is None or cap != self.capturing or self.buf.device != torch.device(device)
                or self.off + need > self.CHUNK):
            self.buf = torch.zeros(self.CHUNK, dtype=torch.uint8, device=device)
            self.off, self.capturing = 0, cap
        t = self.buf[self.off:self.off + nbytes].view(dtype).view(shape)
        self.off += need
        return t


_arena = _ZeroArena()


def zeros(shape, dtype=F32, device="cuda"):
    """torch.zeros for the many small accumulate-into outputs of a step (see _ZeroArena)."""
    if isinstance(shape, int):
        shape = (shape,)
    return _arena.take(tuple(shape), dtype, device)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _chk2d(t, dtype, name):
    if t.dtype != dtype or t.dim() != 2 or t.stride(1) != 1 or not t.is_cuda:
        raise ValueError(f"vtb200.{name}: expected 2-D row-major {dtype} CUDA tensor, got "
                         f"{t.dtype} {tuple(t.shape)} strides {t.stride()} on {t.device}")


# ------------------------------------------------------------------------------------------------ validation mode
# Forward-only switch for parity checks at the north star's tolerance (logits within rtol 1e-3 of the reference's fp32
# forward): activations and weights stay fp32 between kernels, every GEMM runs on the same tcgen05 kernel with 3-way split
# bf16 operands (K' = 3K: a_hi b_hi + a_lo b_hi + a_hi b_lo accumulated in fp32), attention runs the exact-softmax fp32
# kernel, SiLU without the tanh.approx shortcut.  See include/vtb200.h ("Validation mode").
PRECISE = False


class validation_mode:
    """with vtb200.ops.validation_mode(): logits = model(x)   — requires torch.no_grad()."""

    def __enter__(self):
        global PRECISE
        if torch.is_grad_enabled():
            raise RuntimeError("vtb200 validation mode is forward-only: enter it under torch.no_grad()")
        self._prev, PRECISE = PRECISE, True
        return self

    def __exit__(self, *exc):
        global PRECISE
        PRECISE = self._prev
        return False


def act_dtype():
    """dtype of the activations that travel between kernels (GEMM operands, attention inputs)."""
    return F32 if PRECISE else BF16


def split3(x, b_side):
    """fp32 [rows, K] (row-strided view allowed) -> bf16 [rows, 3K]: [hi | lo | hi] (A side) or [hi | hi | lo] (B side)."""
    lib = _l.get()
    _chk2d(x, F32, "split3")
    rows, K = x.shape
    dst = torch.empty((rows, 3 * K), dtype=BF16, device=x.device)
    with _prof("split3_bf16"):
        _l.check(lib.vtb_split3_bf16(_p(x), x.stride(0), rows, K, int(b_side), _p(dst), _stream()), lib)
    _count()
    return dst


def _gemm_precise(a, b, *, out, bias, resid, row_scale, rows_per_scale, epilogue, out2):
    """C = A B^T (+ epilogue) with fp32 operands through the split-operand tcgen05 GEMM; C is fp32."""
    if a.dtype != F32 or b.dtype != F32:
        raise ValueError("vtb200.gemm (validation mode): fp32 operands expected")
    if out is not None and out.dtype != F32:
        raise ValueError("vtb200.gemm (validation mode): fp32 output expected")
    a3, b3 = split3(a, False), split3(b, True)
    if epilogue == _l.EPI_SILU_DUAL:  # out = pre-activation, out2 = SiLU(out): exact SiLU as its own pass
        pre = gemm(a3, b3, out=out, out_dtype=F32, bias=bias)
        if out2 is not None:
            silu_fwd(pre, out=out2)
        return pre
    if epilogue != _l.EPI_NONE:
        raise ValueError("vtb200.gemm (validation mode): forward epilogues only")
    return gemm(a3, b3, out=out, out_dtype=F32, bias=bias, resid=resid, row_scale=row_scale, rows_per_scale=rows_per_scale)


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=BF16, bias=None, resid=None,
         row_scale=None, rows_per_scale=0, epilogue=_l.EPI_NONE, aux=None, out2=None, accumulate=False,
         splits=0, alpha=1.0, a_colsum=None):
    """C[M,N] = alpha * sum_k A(m,k) B(n,k) with the fused epilogue of vtb_gemm_bf16.

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] or [K,N] (b_mn=True); both bf16 row-major
    (leading stride may exceed the row length: views into wider buffers are fine).
    """
    lib = _l.get()
    if PRECISE and (a.dtype == F32 or b.dtype == F32):
        if a_mn or b_mn or aux is not None or accumulate or a_colsum is not None or alpha != 1.0:
            raise ValueError("vtb200.gemm (validation mode): forward products only")
        return _gemm_precise(a, b, out=out, bias=bias, resid=resid, row_scale=row_scale, rows_per_scale=rows_per_scale,
                             epilogue=epilogue, out2=out2)
    _chk2d(a, BF16, "gemm(a)")
    _chk2d(b, BF16, "gemm(b)")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"vtb200.gemm: K mismatch {K} vs {Kb}")
    rows_out = M
    if out is None:
        out = zeros((M, N), out_dtype, a.device) if accumulate else torch.empty((M, N), dtype=out_dtype, device=a.device)
    _chk2d(out, out.dtype, "gemm(out)")
    if out.shape[0] < rows_out or out.shape[1] != N:
        raise ValueError(f"vtb200.gemm: out shape {tuple(out.shape)} vs ({rows_out},{N})")
    p = _l.GemmParams()
    p.M, p.N, p.K = M, N, K
    p.A, p.lda, p.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn)
    p.B, p.ldb, p.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn)
    p.out, p.ldo, p.out_f32 = out.data_ptr(), out.stride(0), int(out.dtype == F32)
    if out.dtype not in (F32, BF16):
        raise ValueError("vtb200.gemm: out must be f32 or bf16")
    if out2 is not None:
        _chk2d(out2, BF16, "gemm(out2)")
        if out2.stride(0) != out.stride(0) or out2.shape != out.shape:
            raise ValueError("vtb200.gemm: out2 must match out")
        p.out2 = out2.data_ptr()
    if bias is not None:
        if bias.dtype != F32 or bias.numel() != N or not bias.is_contiguous():
            raise ValueError("vtb200.gemm: bias must be contiguous f32 [N]")
        p.bias = bias.data_ptr()
    if resid is not None:
        _chk2d(resid, F32, "gemm(resid)")
        p.resid, p.ldr = resid.data_ptr(), resid.stride(0)
    if row_scale is not None:
        if row_scale.dtype != F32 or not row_scale.is_contiguous() or rows_per_scale <= 0:
            raise ValueError("vtb200.gemm: row_scale must be contiguous f32 with rows_per_scale > 0")
        if row_scale.numel() * rows_per_scale < M:
            raise ValueError("vtb200.gemm: row_scale too short")
        p.row_scale, p.rows_per_scale = row_scale.data_ptr(), rows_per_scale
    if aux is not None:
        _chk2d(aux, BF16, "gemm(aux)")
        p.aux, p.ldaux = aux.data_ptr(), aux.stride(0)
    p.epilogue, p.splits, p.accumulate = epilogue, splits, int(accumulate)
    p.alpha = alpha
    if a_colsum is not None:  # column sums of the MN-major A operand (the bias gradient that goes with a wgrad)
        if not a_mn or a_colsum.dtype != F32 or a_colsum.numel() != M or not a_colsum.is_contiguous():
            raise ValueError("vtb200.gemm: a_colsum needs a_mn=True and a contiguous f32 [M] accumulator")
        p.a_colsum = a_colsum.data_ptr()
    kind = ("wgrad" if a_mn else ("dgrad" if b_mn else "fwd"))
    osz = 4 if out.dtype == F32 else 2
    nbytes = 2.0 * K * (M + N) + float(M) * N * (osz * (2 if accumulate else 1) + (2 if out2 is not None else 0)
                                                  + (4 if resid is not None else 0) + (2 if aux is not None else 0))
    with _prof(f"gemm_{kind}[{M}x{N}x{K}]", 2.0 * M * N * K, nbytes):
        _l.check(lib.vtb_gemm_bf16(C.byref(p), _stream()), lib)
    _count()
    return out


def layernorm_fwd(x, gamma, beta, eps, *, out_dtype=BF16, patchify=None, rowmod_add=None, group_rows=0):
    """x f32 [rows, cols] (or NHWC [B,H,W,C] with patchify=(s,H,W)) -> y, mean, rstd."""
    lib = _l.get()
    if x.dtype != F32 or not x.is_contiguous():
        raise ValueError("vtb200.layernorm_fwd: x must be contiguous f32")
    if patchify is None:
        cols = x.shape[-1]
        rows = x.numel() // cols
        s, H, W = 0, 0, 0
    else:
        s, H, W = patchify
        Cc = x.shape[-1]
        cols = s * s * Cc
        rows = x.numel() // cols
    if PRECISE:
        out_dtype = F32
    y = torch.empty((rows, cols), dtype=out_dtype, device=x.device)
    mean = torch.empty(rows, dtype=F32, device=x.device)
    rstd = torch.empty(rows, dtype=F32, device=x.device)
    with _prof("layernorm_fwd"):
        _l.check(lib.vtb_layernorm_fwd(_p(x), _p(gamma), _p(beta), float(eps), rows, cols, s, H, W, _p(y),
                                       int(out_dtype == F32), _p(mean), _p(rstd), _p(rowmod_add),
                                       group_rows, _stream()), lib)
    _count()
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, *, dx_in=None, dx_out=None, patchify=None, want_bf16=False,
                  row_scale=None, rows_per_scale=0, dgamma=None, dbeta=None, colsum_out=None):
    """Returns dx (f32, same shape as x), optional bf16 scaled copy, dgamma, dbeta (accumulated).
    colsum_out (f32 [cols], pre-zeroed) additionally receives the column sums of the bf16 copy."""
    lib = _l.get()
    cols = dy.shape[-1]
    rows = dy.numel() // cols
    if not dy.is_contiguous() or dy.dtype not in (F32, BF16):
        raise ValueError("vtb200.layernorm_bwd: dy must be contiguous f32/bf16")
    s, H, W = patchify if patchify is not None else (0, 0, 0)
    if dx_out is None:
        dx_out = torch.empty_like(x)
    dxb = torch.empty(x.shape, dtype=BF16, device=x.device) if want_bf16 else None
    if dgamma is None:
        dgamma = zeros(cols, F32, x.device)
    if dbeta is None:
        dbeta = zeros(cols, F32, x.device)
    with _prof("layernorm_bwd"):
        _l.check(lib.vtb_layernorm_bwd(_p(dy), int(dy.dtype == F32), _p(x), _p(gamma), _p(mean), _p(rstd),
                                       rows, cols, s, H, W, _p(dx_in), _p(dx_out), _p(dxb), _p(row_scale),
                                       rows_per_scale, _p(dgamma), _p(dbeta), _p(colsum_out), _stream()), lib)
    _count()
    return dx_out, dxb, dgamma, dbeta


_MASK_BITS = {}


def _mask_bits(mask, n):
    """[n_mask, 2, 64] int64 row words of a uint8 mask [n_mask, n, pitch] (include/vtb200.h: mask_bits): word
    [m, 0, i] bit j / word [m, 1, j] bit i set where mask[m, i, j] != 0.  Built once per mask buffer (the masks are
    constant module buffers) with torch integer ops."""
    key = (mask.data_ptr(), mask._version, tuple(mask.shape), n)
    bits = _MASK_BITS.get(key)
    if bits is None:
        m = (mask[:, :, :n] != 0).to(torch.int64)
        sh = torch.arange(n, device=mask.device, dtype=torch.int64)
        bits = torch.zeros((mask.shape[0], 2, 64), dtype=torch.int64, device=mask.device)
        bits[:, 0, :n] = (m << sh.view(1, 1, n)).sum(2)
        bits[:, 1, :n] = (m << sh.view(1, n, 1)).sum(1)
        if len(_MASK_BITS) > 256:
            _MASK_BITS.clear()
        _MASK_BITS[key] = bits
    return bits


class AttnSpec:
    """Geometry of one attention call (mirrors vtb_attn_params)."""

    def __init__(self, mode, batch, heads, dh, nq, nkv, Hs=0, Ws=0, window=0, shift=0, halo=0,
                 rel_bias=None, pos=None, mask=None):
        self.mode, self.batch, self.heads, self.dh = mode, batch, heads, dh
        self.nq, self.nkv, self.Hs, self.Ws = nq, nkv, Hs, Ws
        self.window, self.shift, self.halo = window, shift, halo
        self.rel_bias, self.pos, self.mask = rel_bias, pos, mask
        self.scale = 1.0 / math.sqrt(dh)

    @property
    def groups(self):
        if self.mode == _l.ATTN_GLOBAL:
            return self.batch
        return self.batch * (self.Hs // self.window) * (self.Ws // self.window)

    def fill(self, p):
        p.mode, p.batch, p.heads, p.dh = self.mode, self.batch, self.heads, self.dh
        p.nq, p.nkv, p.Hs, p.Ws = self.nq, self.nkv, self.Hs, self.Ws
        p.window, p.shift, p.halo = self.window, self.shift, self.halo
        p.scale = self.scale
        if self.rel_bias is not None:
            if self.rel_bias.dtype != F32 or not self.rel_bias.is_contiguous():
                raise ValueError("vtb200.attention: rel_bias must be contiguous f32 [n_pos, heads]")
            if self.pos.dtype != torch.int32 or not self.pos.is_contiguous():
                raise ValueError("vtb200.attention: pos must be contiguous int32 [nq, nkv]")
            p.rel_bias, p.pos, p.n_pos = self.rel_bias.data_ptr(), self.pos.data_ptr(), self.rel_bias.shape[0]
        if self.mask is not None:
            if self.mask.dtype != torch.uint8 or not self.mask.is_contiguous() or self.mask.dim() != 3:
                raise ValueError("vtb200.attention: mask must be contiguous uint8 [n_mask, nq, nkv or padded pitch]")
            if self.mask.shape[1] != self.nq or self.mask.shape[2] < self.nkv:
                raise ValueError("vtb200.attention: mask shape does not match nq / nkv")
            p.mask, p.n_mask, p.mask_ld = self.mask.data_ptr(), self.mask.shape[0], self.mask.shape[2]
            if self.mode == _l.ATTN_WINDOW and self.nq <= 64 and self.nkv == self.nq:
                self._bits = _mask_bits(self.mask, self.nq)  # keep alive for the call
                p.mask_bits = self._bits.data_ptr()


def _qkv_fill(p, q, k, v):
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _chk2d(t, BF16, f"attention({nm})")
    p.q, p.ldq = q.data_ptr(), q.stride(0)
    p.k, p.ldk = k.data_ptr(), k.stride(0)
    p.v, p.ldv = v.data_ptr(), v.stride(0)


def _attention_fwd_precise(spec, q, k, v):
    """Validation mode: fp32 q / k / v views -> fp32 o (exact softmax on the CUDA cores, same geometry descriptor)."""
    lib = _l.get()
    p = _l.AttnParams()
    spec.fill(p)
    p.mask_bits = None
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _chk2d(t, F32, f"attention({nm}, validation mode)")
    p.q, p.ldq = q.data_ptr(), q.stride(0)
    p.k, p.ldk = k.data_ptr(), k.stride(0)
    p.v, p.ldv = v.data_ptr(), v.stride(0)
    o = torch.empty((q.shape[0], spec.heads * spec.dh), dtype=F32, device=q.device)
    lse = torch.empty((spec.groups, spec.heads, spec.nq), dtype=F32, device=q.device)
    p.o, p.ldo, p.lse = o.data_ptr(), o.stride(0), lse.data_ptr()
    with _prof("attention_fwd_f32"):
        _l.check(lib.vtb_attention_fwd_f32(C.byref(p), _stream()), lib)
    _count()
    return o, lse


def attention_fwd(spec, q, k, v):
    """q [Tq, >=H*dh], k/v [Tkv, >=H*dh] bf16 views (row-major). Returns o [Tq, H*dh] bf16, lse."""
    lib = _l.get()
    if PRECISE:
        return _attention_fwd_precise(spec, q, k, v)
    p = _l.AttnParams()
    spec.fill(p)
    _qkv_fill(p, q, k, v)
    o = torch.empty((q.shape[0], spec.heads * spec.dh), dtype=BF16, device=q.device)
    lse = torch.empty((spec.groups, spec.heads, spec.nq), dtype=F32, device=q.device)
    p.o, p.ldo, p.lse = o.data_ptr(), o.stride(0), lse.data_ptr()
    hd = spec.heads * spec.dh
    with _prof("attention_fwd", 4.0 * spec.groups * spec.heads * spec.nq * spec.nkv * spec.dh,
               2.0 * hd * (2 * q.shape[0] + 2 * k.shape[0])):  # q, k, v read once, o written once (bf16)
        _l.check(lib.vtb_attention_fwd(C.byref(p), _stream()), lib)
    _count()
    return o, lse


def attention_bwd(spec, q, k, v, o, lse, dout, dq, dk, dv, drel_bias=None, dkv_f32=False):
    """Writes dq/dk/dv (views with the same layout as q/k/v; dk/dv f32 + pre-zeroed when dkv_f32)."""
    lib = _l.get()
    p = _l.AttnParams()
    spec.fill(p)
    _qkv_fill(p, q, k, v)
    _chk2d(o, BF16, "attention_bwd(o)")
    _chk2d(dout, BF16, "attention_bwd(dout)")
    delta = torch.empty_like(lse)
    p.o, p.ldo, p.lse = o.data_ptr(), o.stride(0), lse.data_ptr()
    p.dout, p.lddo = dout.data_ptr(), dout.stride(0)
    p.dq, p.lddq = dq.data_ptr(), dq.stride(0)
    p.dk, p.lddk = dk.data_ptr(), dk.stride(0)
    p.dv, p.lddv = dv.data_ptr(), dv.stride(0)
    p.dkv_f32 = int(dkv_f32)
    p.delta = delta.data_ptr()
    if drel_bias is not None:
        p.drel_bias = drel_bias.data_ptr()
    ws_bytes = lib.vtb_attention_bwd_workspace_bytes(C.byref(p))
    if ws_bytes > 0:  # tcgen05 halo kernels: per-block partial dK / dV rows, summed per token by a second kernel
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device)
        p.ws, p.ws_bytes = ws.data_ptr(), ws_bytes
    hd = spec.heads * spec.dh
    with _prof("attention_bwd", 10.0 * spec.groups * spec.heads * spec.nq * spec.nkv * spec.dh,
               2.0 * hd * (4 * q.shape[0] + 4 * k.shape[0])):  # q, k, v, o, do read; dq, dk, dv written (bf16)
        _l.check(lib.vtb_attention_bwd(C.byref(p), _stream()), lib)
    _count(2)


# Set by vtb200.multi.enable_weight_arena: maps an f32 weight to its bf16 copy in a model's arena (refreshed by one
# multi-tensor cast per top-level forward) or None.  WEIGHT_EPOCH is bumped by every library call that rewrites
# parameters in place (multi-tensor EMA / AdamW), which torch's version counters cannot see.
WEIGHT_LOOKUP = None
WEIGHT_EPOCH = 0


def assert_weight_fresh(w):
    """Backward-side guard of the weight arena (vtb200.multi.WeightArena): a bf16 weight view saved by a forward must still
    hold the weights of that forward when the backward reads it."""
    tag = getattr(w, "_vtb_arena_gen", None)
    if tag is not None and tag[0].generation != tag[1]:
        raise RuntimeError("vtb200: this graph's forward ran before a weight update and its backward runs after a later "
                           "forward re-cast the weight arena; the saved bf16 weights are gone.  Run backward before the next "
                           "forward of the updated model, or disable_weight_arena(model).")


def cast_bf16(src):
    lib = _l.get()
    if src.dtype != F32 or not src.is_contiguous():
        raise ValueError("vtb200.cast_bf16: contiguous f32 expected")
    if PRECISE:
        return src  # validation mode: weights reach the GEMM in fp32 and are split there
    if WEIGHT_LOOKUP is not None:
        hit = WEIGHT_LOOKUP(src)
        if hit is not None:
            return hit
    dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    if src.numel():
        with _prof("cast_f32_bf16"):
            _l.check(lib.vtb_cast_f32_bf16(_p(src), _p(dst), src.numel(), _stream()), lib)
        _count()
    return dst


def cast_bf16_2d(src, dst):
    """dst[r, c] = bf16(src[r, c]) for 2-D row-strided views (src f32, dst bf16, same shape)."""
    lib = _l.get()
    _chk2d(src, F32, "cast_bf16_2d(src)")
    _chk2d(dst, BF16, "cast_bf16_2d(dst)")
    if src.shape != dst.shape:
        raise ValueError("vtb200.cast_bf16_2d: shape mismatch")
    with _prof("cast_f32_bf16_2d"):
        _l.check(lib.vtb_cast_f32_bf16_2d(_p(src), src.stride(0), _p(dst), dst.stride(0), src.shape[0],
                                          src.shape[1], _stream()), lib)
    _count()
    return dst


def scale_cast_bf16(src, row_scale=None, rows_per_scale=0):
    """bf16(src * row_scale[row // rows_per_scale]); src f32 [..., cols] contiguous."""
    lib = _l.get()
    if src.dtype != F32 or not src.is_contiguous():
        raise ValueError("vtb200.scale_cast_bf16: contiguous f32 expected")
    cols = src.shape[-1]
    rows = src.numel() // cols
    if PRECISE and row_scale is None:
        return src.view(rows, cols)  # validation mode: the operand stays fp32
    dst = torch.empty((rows, cols), dtype=BF16, device=src.device)
    with _prof("scale_cast_bf16"):
        _l.check(lib.vtb_scale_cast_bf16(_p(src), _p(row_scale), rows_per_scale, rows, cols, _p(dst),
                                         _stream()), lib)
    _count()
    return dst


def scale_cast_colsum_bf16(src, row_scale=None, rows_per_scale=0):
    """(bf16(src * row_scale[row // rows_per_scale]), its column sums) in one pass over src."""
    lib = _l.get()
    if src.dtype != F32 or not src.is_contiguous():
        raise ValueError("vtb200.scale_cast_colsum_bf16: contiguous f32 expected")
    cols = src.shape[-1]
    rows = src.numel() // cols
    dst = torch.empty((rows, cols), dtype=BF16, device=src.device)
    cs = zeros(cols, F32, src.device)
    with _prof("scale_cast_colsum_bf16"):
        _l.check(lib.vtb_scale_cast_colsum_bf16(_p(src), _p(row_scale), rows_per_scale, rows, cols, _p(dst), _p(cs),
                                                _stream()), lib)
    _count()
    return dst, cs


def colsum(x, out=None):
    lib = _l.get()
    _chk2d(x, BF16, "colsum")
    M, N = x.shape
    if out is None:
        out = zeros(N, F32, x.device)
    with _prof("colsum_bf16"):
        _l.check(lib.vtb_colsum_bf16(_p(x), M, N, x.stride(0), _p(out), _stream()), lib)
    _count()
    return out


def dropout(x, keep, scale, *, out=None, resid=None, row_scale=None, rows_per_scale=0):
    """out = resid + row_scale[row] * (keep ? x * scale : 0) for contiguous f32 / bf16 x and a bool / uint8 keep mask of the
    same shape (nn.Dropout with a caller-drawn mask; its adjoint is the same call on the gradient).  out may be x."""
    lib = _l.get()
    if x.dtype not in (F32, BF16) or not x.is_contiguous() or not keep.is_contiguous() or keep.numel() != x.numel() \
            or keep.element_size() != 1:
        raise ValueError("vtb200.dropout: contiguous f32 / bf16 input and a one-byte keep mask of the same size expected")
    if out is None:
        out = torch.empty_like(x)
    if out.dtype != x.dtype or not out.is_contiguous() or out.numel() != x.numel():
        raise ValueError("vtb200.dropout: out must match x")
    if resid is not None and (x.dtype != F32 or resid.dtype != F32 or not resid.is_contiguous() or resid.numel() != x.numel()):
        raise ValueError("vtb200.dropout: resid needs contiguous f32 operands of the same size")
    eps = 0
    if row_scale is not None:
        if rows_per_scale <= 0 or row_scale.dtype != F32:
            raise ValueError("vtb200.dropout: row_scale needs rows_per_scale > 0 and f32 scales")
        eps = rows_per_scale * x.shape[-1]
    with _prof("dropout"):
        _l.check(lib.vtb_dropout(_p(x), _p(keep), float(scale), x.numel(), int(x.dtype == F32), _p(resid), _p(row_scale),
                                 eps, _p(out), _stream()), lib)
    _count()
    return out


def patch_gather(src, *, nchw, c_major, B, Cc, H, W, p):
    lib = _l.get()
    if not src.is_contiguous() or src.dtype not in (F32, BF16):
        raise ValueError("vtb200.patch_gather: contiguous f32/bf16 expected")
    rows = B * (H // p) * (W // p)
    if PRECISE:
        if src.dtype != F32:
            raise ValueError("vtb200.patch_gather (validation mode): f32 source expected")
        dst = torch.empty((rows, p * p * Cc), dtype=F32, device=src.device)
        with _prof("patch_gather_f32"):
            _l.check(lib.vtb_patch_gather_f32(_p(src), int(nchw), int(c_major), B, Cc, H, W, p, _p(dst), _stream()), lib)
        _count()
        return dst
    dst = torch.empty((rows, p * p * Cc), dtype=BF16, device=src.device)
    with _prof("patch_gather"):
        _l.check(lib.vtb_patch_gather(_p(src), int(src.dtype == BF16), int(nchw), int(c_major), B, Cc, H, W,
                                      p, _p(dst), _stream()), lib)
    _count()
    return dst


def patch_scatter(dA, *, c_major, B, Cc, H, W, p, dx=None, accumulate=False, dst_nchw=False):
    lib = _l.get()
    if not dA.is_contiguous() or dA.dtype not in (F32, BF16):
        raise ValueError("vtb200.patch_scatter: contiguous f32/bf16 expected")
    if dx is None:
        dx = torch.empty((B, Cc, H, W) if dst_nchw else (B, H, W, Cc), dtype=F32, device=dA.device)
    with _prof("patch_scatter"):
        _l.check(lib.vtb_patch_scatter(_p(dA), int(dA.dtype == F32), int(c_major), B, Cc, H, W, p, _p(dx),
                                       int(accumulate), int(dst_nchw), _stream()), lib)
    _count()
    return dx


def vit_assemble_tokens(tok, cls, pos, B, n, D):
    """x[b,0] = cls + pos[0]; x[b,1+p] = tok[b*n+p] + pos[1+p]  ->  f32 [B, n+1, D]."""
    lib = _l.get()
    x = torch.empty((B, n + 1, D), dtype=F32, device=tok.device)
    with _prof("vit_assemble_tokens"):
        _l.check(lib.vtb_vit_assemble_tokens(_p(tok), _p(cls), _p(pos), B, n, D, _p(x), _stream()), lib)
    _count()
    return x


def fill_rows(x, group_stride, groups, cols, a, b=None):
    lib = _l.get()
    with _prof("fill_rows"):
        _l.check(lib.vtb_fill_rows(_p(x), group_stride, groups, cols, _p(a), _p(b), _stream()), lib)
    _count()


def rowgroup_sum(x, group_stride, groups, rows, cols, out):
    lib = _l.get()
    with _prof("rowgroup_sum"):
        _l.check(lib.vtb_rowgroup_sum(_p(x), group_stride, groups, rows, cols, _p(out), _stream()), lib)
    _count()


def mean_rows_fwd(x, groups, n, cols):
    lib = _l.get()
    out = torch.empty((groups, cols), dtype=F32, device=x.device)
    with _prof("mean_rows_fwd"):
        _l.check(lib.vtb_mean_rows_fwd(_p(x), groups, n, cols, _p(out), _stream()), lib)
    _count()
    return out


def mean_rows_bwd(dy, groups, n, cols):
    lib = _l.get()
    dx = torch.empty((groups * n, cols), dtype=F32, device=dy.device)
    with _prof("mean_rows_bwd"):
        _l.check(lib.vtb_mean_rows_bwd(_p(dy), groups, n, cols, _p(dx), _stream()), lib)
    _count()
    return dx


def silu_fwd(x, out=None):
    lib = _l.get()
    y = torch.empty_like(x) if out is None else out
    if y.dtype != F32 or x.dtype != F32 or not y.is_contiguous() or not x.is_contiguous() or y.numel() != x.numel():
        raise ValueError("vtb200.silu_fwd: contiguous f32 tensors of one size expected")
    with _prof("silu_fwd"):
        fn = lib.vtb_silu_fwd_exact if PRECISE else lib.vtb_silu_fwd
        _l.check(fn(_p(x), _p(y), x.numel(), _stream()), lib)
    _count()
    return y


def silu_bwd(x, dy):
    lib = _l.get()
    dx = torch.empty_like(x)
    with _prof("silu_bwd"):
        _l.check(lib.vtb_silu_bwd(_p(x), _p(dy), _p(dx), x.numel(), _stream()), lib)
    _count()
    return dx


def transpose_hw(x, B, H, W, Cc):
    """[B,H,W,C] -> [B,W,H,C] copy (bf16 or f32)."""
    lib = _l.get()
    if not x.is_contiguous() or x.dtype not in (F32, BF16):
        raise ValueError("vtb200.transpose_hw: contiguous f32/bf16 expected")
    dst = torch.empty((B, W, H, Cc), dtype=x.dtype, device=x.device)
    with _prof("transpose_hw"):
        _l.check(lib.vtb_transpose_hw(_p(x), _p(dst), int(x.dtype == F32), B, H, W, Cc, _stream()), lib)
    _count()
    return dst


def dwconv3x3_fwd(x, w):
    lib = _l.get()
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    with _prof("dwconv3x3_fwd"):
        _l.check(lib.vtb_dwconv3x3_fwd(_p(x), _p(w), B, H, W, Cc, _p(y), _stream()), lib)
    _count()
    return y


def dwconv3x3_bwd(x, w, dy):
    lib = _l.get()
    B, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    dw = torch.zeros((Cc, 1, 3, 3), dtype=F32, device=x.device)
    with _prof("dwconv3x3_bwd"):
        _l.check(lib.vtb_dwconv3x3_bwd(_p(x), _p(w), _p(dy), B, H, W, Cc, _p(dx), _p(dw), _stream()), lib)
    _count()
    return dx, dw


def dino_loss(student, teacher, center, n_crops, t_student, t_teacher, want_grad=True):
    """Fused DINOLoss.forward (loss.py:119-142) and its gradient.  student f32 [n_crops*B, K], teacher f32 [2*B, K],
    center f32 [1, K] or [K].  Returns (loss f32 [1], dstudent f32 like student or None)."""
    lib = _l.get()
    for t, nm in ((student, "student"), (teacher, "teacher"), (center, "center")):
        if t.dtype != F32 or not t.is_contiguous():
            raise ValueError(f"vtb200.dino_loss: {nm} must be contiguous f32")
    K = student.shape[-1]
    B = teacher.shape[0] // 2
    if teacher.shape != (2 * B, K) or student.shape != (n_crops * B, K) or center.numel() != K:
        raise ValueError("vtb200.dino_loss: shapes must be student [n_crops*B, K], teacher [2*B, K], center [K]")
    loss = zeros(1, F32, student.device)
    dstudent = torch.empty_like(student) if want_grad else None
    with _prof("dino_loss", 0.0, 4.0 * (2 * (student.numel() + teacher.numel()) + (student.numel() if want_grad else 0))):
        _l.check(lib.vtb_dino_loss(_p(student), _p(teacher), _p(center), n_crops, B, K, float(t_student),
                                   float(t_teacher), _p(loss), _p(dstudent), _stream()), lib)
    _count()
    return loss, dstudent


# ------------------------------------------------------------------------------------------------ DINO head rows
def l2norm_fwd(x, eps=1e-12):
    """F.normalize(x, dim=-1) as the bf16 GEMM operand.  x f32 [rows, cols] -> (y bf16 [rows, cols], inv f32 [rows])."""
    lib = _l.get()
    _chk2d(x, F32, "l2norm_fwd(x)")
    if not x.is_contiguous():
        raise ValueError("vtb200.l2norm_fwd: contiguous rows expected")
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    inv = torch.empty(x.shape[0], dtype=F32, device=x.device)
    with _prof("l2norm_fwd", 0.0, 6.0 * x.numel()):
        _l.check(lib.vtb_l2norm_fwd(_p(x), x.shape[0], x.shape[1], float(eps), _p(y), _p(inv), _stream()), lib)
    _count()
    return y, inv


def l2norm_bwd(dy, x, inv):
    lib = _l.get()
    for t, nm in ((dy, "dy"), (x, "x")):
        _chk2d(t, F32, f"l2norm_bwd({nm})")
        if not t.is_contiguous():
            raise ValueError("vtb200.l2norm_bwd: contiguous rows expected")
    dx = torch.empty_like(x)
    with _prof("l2norm_bwd", 0.0, 12.0 * x.numel()):
        _l.check(lib.vtb_l2norm_bwd(_p(dy), _p(x), _p(inv), x.shape[0], x.shape[1], _p(dx), _stream()), lib)
    _count()
    return dx


def weight_norm_fwd(v, g):
    """nn.utils.weight_norm (dim=0): w = v * g / ||v||_row, emitted as bf16.  v f32 [rows, cols], g f32 with `rows`
    elements -> (w bf16 [rows, cols], inv f32 [rows])."""
    lib = _l.get()
    _chk2d(v, F32, "weight_norm_fwd(v)")
    if not v.is_contiguous() or g.dtype != F32 or g.numel() != v.shape[0] or not g.is_contiguous():
        raise ValueError("vtb200.weight_norm_fwd: v [rows, cols] and g [rows(, 1)] contiguous f32 expected")
    w = torch.empty(v.shape, dtype=BF16, device=v.device)
    inv = torch.empty(v.shape[0], dtype=F32, device=v.device)
    with _prof("weight_norm_fwd", 0.0, 6.0 * v.numel()):
        _l.check(lib.vtb_weight_norm_fwd(_p(v), _p(g), v.shape[0], v.shape[1], _p(w), _p(inv), _stream()), lib)
    _count()
    return w, inv


def weight_norm_bwd(dw, v, g, inv, want_dg=True):
    lib = _l.get()
    for t, nm in ((dw, "dw"), (v, "v")):
        _chk2d(t, F32, f"weight_norm_bwd({nm})")
        if not t.is_contiguous():
            raise ValueError("vtb200.weight_norm_bwd: contiguous rows expected")
    dv = torch.empty_like(v)
    dg = torch.empty(g.shape, dtype=F32, device=v.device) if want_dg else None
    with _prof("weight_norm_bwd", 0.0, 12.0 * v.numel()):
        _l.check(lib.vtb_weight_norm_bwd(_p(dw), _p(v), _p(g), _p(inv), v.shape[0], v.shape[1], _p(dv), _p(dg),
                                         _stream()), lib)
    _count()
    return dv, dg


def gelu_fwd(x, want_f32=False):
    """nn.GELU() (exact): x f32 -> bf16 copy of the activation (the next Linear's operand) and, if asked, f32."""
    lib = _l.get()
    if x.dtype != F32 or not x.is_contiguous():
        raise ValueError("vtb200.gelu_fwd: contiguous f32 expected")
    yb = torch.empty(x.shape, dtype=BF16, device=x.device)
    y = torch.empty_like(x) if want_f32 else None
    with _prof("gelu_fwd", 0.0, (10.0 if want_f32 else 6.0) * x.numel()):
        _l.check(lib.vtb_gelu_fwd(_p(x), _p(y), _p(yb), x.numel(), _stream()), lib)
    _count()
    return yb, y


def gelu_bwd(x, dy):
    lib = _l.get()
    if x.dtype != F32 or dy.dtype != F32 or not x.is_contiguous() or not dy.is_contiguous() or x.shape != dy.shape:
        raise ValueError("vtb200.gelu_bwd: contiguous f32 tensors of one shape expected")
    dx = torch.empty_like(x)
    with _prof("gelu_bwd", 0.0, 12.0 * x.numel()):
        _l.check(lib.vtb_gelu_bwd(_p(x), _p(dy), _p(dx), x.numel(), _stream()), lib)
    _count()
    return dx
