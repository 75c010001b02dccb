"""Tensor-list wrappers over the multi-tensor C-ABI (include/vtb200.h, `vtb_mt_*`, `vtb_mix_loss`).

The step-side loops of the reference (train.py:285-299, train_dino.py:236-261, train_util.py:70-84, optimizer.py:12-26)
walk the parameter list in Python and launch 2-8 tiny ATen kernels per tensor; here a list is packed once into host
arrays of device pointers and one library call covers it.  No CPU fallback: every function raises without the library.
"""
import ctypes as C

import torch

from . import lib as _l
from . import ops as _ops
from .ops import BF16, F32, _count, _p, _prof, _stream

MT_CHUNK = 8192  # VTB_MT_CHUNK


class TensorList:
    """Host-side pack of a list of CUDA tensors: `ptrs` (void*[n]) and `numel` (int64[n]).  Keeps the tensors alive."""

    __slots__ = ("tensors", "ptrs", "numel", "n", "total")

    def __init__(self, tensors, dtype=F32, name="list"):
        tensors = list(tensors)
        for i, t in enumerate(tensors):
            if t is None:
                continue
            if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
                raise ValueError(f"vtb200.multi: {name}[{i}] must be a contiguous {dtype} CUDA tensor, got "
                                 f"{t.dtype} {tuple(t.shape)} on {t.device}")
        self.tensors = tensors
        self.n = len(tensors)
        self.ptrs = (C.c_void_p * max(self.n, 1))(*[None if t is None else t.data_ptr() for t in tensors])
        self.numel = (C.c_int64 * max(self.n, 1))(*[0 if t is None else t.numel() for t in tensors])
        self.total = sum(0 if t is None else t.numel() for t in tensors)

    def same_shapes(self, other, name):
        if self.n != other.n or any(a != b for a, b in zip(self.numel[:self.n], other.numel[:other.n])):
            raise ValueError(f"vtb200.multi.{name}: the two lists must hold tensors of equal sizes")


def _as_list(x, dtype=F32, name="list"):
    return x if isinstance(x, TensorList) else TensorList(x, dtype, name)


def num_chunks(tl):
    return int(_l.load().vtb_mt_num_chunks(tl.numel, tl.n))


def cast_bf16(src, dst):
    """dst[i] = bf16(src[i]) for every tensor: autocast's per-weight casts (train.py:273) in one launch."""
    lib = _l.get()
    src, dst = _as_list(src, F32, "src"), _as_list(dst, BF16, "dst")
    src.same_shapes(dst, "cast_bf16")
    with _prof("mt_cast_f32_bf16", 0.0, 6.0 * src.total):
        _l.check(lib.vtb_mt_cast_f32_bf16(src.ptrs, dst.ptrs, src.numel, src.n, _stream()), lib)
    _count(-(-src.n // 256))
    return dst


def ema(dst, src, decay):
    """dst = dst*decay + src*(1-decay)  (train_util.py:70-84, train_dino.py:257-261)."""
    lib = _l.get()
    dst, src = _as_list(dst, F32, "dst"), _as_list(src, F32, "src")
    dst.same_shapes(src, "ema")
    with _prof("mt_ema", 0.0, 12.0 * src.total):
        _l.check(lib.vtb_mt_ema(dst.ptrs, src.ptrs, src.numel, src.n, float(decay), _stream()), lib)
    _ops.WEIGHT_EPOCH += 1
    _count(-(-src.n // 256))


def grad_norm(grads, max_norm):
    """-> f32 [2] on the device: (global L2 norm, min(1, max_norm / (norm + 1e-6))).  No host synchronisation."""
    lib = _l.get()
    grads = _as_list(grads, F32, "grads")
    dev = grads.tensors[0].device if grads.n else "cuda"
    out = torch.empty(2, dtype=F32, device=dev)
    partials = torch.empty(max(num_chunks(grads), 1), dtype=F32, device=dev)
    with _prof("mt_grad_norm", 0.0, 4.0 * grads.total):
        _l.check(lib.vtb_mt_grad_norm(grads.ptrs, grads.numel, grads.n, float(max_norm), _p(partials), _p(out),
                                      _stream()), lib)
    _count(-(-grads.n // 256) + 1)
    return out


def scale(tensors, scale_dev):
    """x *= scale_dev[0] for every tensor (device scalar; skipped inside the kernel when it is exactly 1)."""
    lib = _l.get()
    tensors = _as_list(tensors, F32, "tensors")
    if scale_dev.dtype != F32 or not scale_dev.is_cuda:
        raise ValueError("vtb200.multi.scale: scale must be an f32 CUDA tensor")
    with _prof("mt_scale", 0.0, 8.0 * tensors.total):
        _l.check(lib.vtb_mt_scale(tensors.ptrs, tensors.numel, tensors.n, _p(scale_dev), _stream()), lib)
    _count(-(-tensors.n // 256))


def agc(params, grads, clipping=0.01, eps=1e-3):
    """Unit-wise adaptive gradient clipping in place on `grads` (optimizer.py:12-26)."""
    lib = _l.get()
    units = [p.shape[0] if p.dim() > 1 else 1 for p in params]
    params, grads = _as_list(params, F32, "params"), _as_list(grads, F32, "grads")
    params.same_shapes(grads, "agc")
    units_c = (C.c_int64 * max(params.n, 1))(*units)
    with _prof("mt_agc", 0.0, 12.0 * params.total):
        _l.check(lib.vtb_mt_agc(params.ptrs, grads.ptrs, params.numel, units_c, params.n, float(clipping), float(eps),
                                _stream()), lib)
    _count(-(-params.n // 256))


def adamw(params, grads, exp_avgs, exp_avg_sqs, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None,
          bf16_out=None):
    """One torch.optim.AdamW step over a parameter group; `grad_scale` (f32 device scalar) multiplies the gradients
    first (the clip coefficient), `bf16_out` (list with None holes) receives bf16 copies of the updated parameters."""
    lib = _l.get()
    params, grads = _as_list(params, F32, "params"), _as_list(grads, F32, "grads")
    exp_avgs, exp_avg_sqs = _as_list(exp_avgs, F32, "exp_avgs"), _as_list(exp_avg_sqs, F32, "exp_avg_sqs")
    for other in (grads, exp_avgs, exp_avg_sqs):
        params.same_shapes(other, "adamw")
    pb = None
    if bf16_out is not None:
        pb = _as_list(bf16_out, BF16, "bf16_out")
        if pb.n != params.n or any(b and b != a for a, b in zip(params.numel[:params.n], pb.numel[:pb.n])):
            raise ValueError("vtb200.multi.adamw: bf16_out must match params (None for no copy)")
    with _prof("mt_adamw", 0.0, (28.0 + (2.0 if pb else 0.0)) * params.total):
        _l.check(lib.vtb_mt_adamw(params.ptrs, grads.ptrs, exp_avgs.ptrs, exp_avg_sqs.ptrs, pb.ptrs if pb else None,
                                  params.numel, params.n, float(lr), float(beta1), float(beta2), float(eps),
                                  float(weight_decay), int(step), _p(grad_scale), _stream()), lib)
    _ops.WEIGHT_EPOCH += 1
    _count(-(-params.n // 256))


def mix_loss(logits, target1, target2=None, inter=None, *, eps=0.0, loss_scale=None, want_loss=True, want_grad=True,
             want_rows=False, want_correct=False, topk=5):
    """Fused MixLoss (+ gradient, + top-1/top-5 hit counts of target1): loss.py:53-86, train_util.py:53-67.
    logits f32 [B, n_class] (row stride allowed), targets int64 [B], inter f32 [B] or None.
    -> (loss f32 [1], row_loss f32 [B] or None, dlogits f32 [B, n_class] or None, correct int32 [2] or None)."""
    lib = _l.get()
    if logits.dtype != F32 or logits.dim() != 2 or logits.stride(1) != 1 or not logits.is_cuda:
        raise ValueError("vtb200.multi.mix_loss: logits must be a 2-D row-major f32 CUDA tensor")
    B, n = logits.shape
    for t, nm in ((target1, "target1"), (target2, "target2")):
        if t is not None and (t.dtype != torch.int64 or t.shape != (B,) or not t.is_contiguous() or not t.is_cuda):
            raise ValueError(f"vtb200.multi.mix_loss: {nm} must be a contiguous int64 CUDA tensor of shape [{B}]")
    if inter is not None and (inter.dtype != F32 or inter.numel() != B or not inter.is_contiguous()):
        raise ValueError(f"vtb200.multi.mix_loss: interpolation must be contiguous f32 with {B} elements")
    dev = logits.device
    loss = _ops.zeros(1, F32, dev) if want_loss else None
    rows = torch.empty(B, dtype=F32, device=dev) if want_rows else None
    dlogits = torch.empty((B, n), dtype=F32, device=dev) if want_grad else None
    correct = _ops.zeros(2, torch.int32, dev) if want_correct else None
    if loss_scale is None:
        loss_scale = 1.0 / max(B, 1)
    with _prof("mix_loss", 0.0, 4.0 * B * n * (2 if want_grad else 1)):
        _l.check(lib.vtb_mix_loss(_p(logits), logits.stride(0), _p(target1), _p(target2), _p(inter), B, n, float(eps),
                                  float(loss_scale), _p(loss), _p(rows), _p(dlogits), _p(correct), int(topk), _stream()),
                 lib)
    _count()
    return loss, rows, dlogits, correct


# ------------------------------------------------------------------------------------------------ weight arena
class WeightArena:
    """bf16 operand copies of a model's weights, refreshed by ONE multi-tensor cast at the start of each top-level
    forward instead of one cast launch per Linear (autocast's behaviour, train.py:273).

    Opt-in (`enable_weight_arena(model)`): the contract is that parameters only change BETWEEN top-level forward
    calls (optimizer step, EMA, load_state_dict), which is how train.py / train_dino.py use the models.  Lookups are
    guarded: an entry is used only if the parameter object is still alive, still owns the same storage address and its
    autograd version is the one that was cast; anything else falls back to a per-call cast.

    The arena is ONE shared buffer that every refresh overwrites in place, and the branch Functions keep the views they were
    handed for their backward (dgrad reads the bf16 weight).  A graph whose forward ran BEFORE a weight update and whose backward
    runs AFTER the next forward would therefore differentiate against the new weights.  That misuse is detected, not silently
    computed: every view carries the arena's content generation (bumped by a refresh that follows a parameter change), and
    `ops.assert_weight_fresh` — called by every dgrad — raises if the generation moved on.  Two forwards with unchanged weights
    (gradient accumulation, the DINO crops) keep the generation and stay legal."""

    def __init__(self, module):
        # weight-normed layers (`weight_v` / `weight_g`, vit.py:244) are excluded: their operand is built by vtb_weight_norm_fwd
        self.params = [p for n, p in module.named_parameters() if p.dim() >= 2 and p.dtype == F32 and p.is_cuda
                       and p.is_contiguous() and not n.endswith(("weight_v", "weight_g"))]
        if not self.params:
            raise ValueError("vtb200.WeightArena: no f32 CUDA weight matrices in this module (move it to the GPU first)")
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 127) // 128 * 128  # 256-byte aligned slices
        self.flat = torch.empty(total, dtype=BF16, device=self.params[0].device)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, self.params)]
        self.src = self.dst = None
        self.ptrs = ()
        self.versions = [-1] * len(self.params)
        self.epoch = -1
        self.index = {}
        self.generation = 0  # content generation: bumped when a refresh follows a change of any parameter

    def refresh(self):
        ptrs = tuple(p.data_ptr() for p in self.params)
        versions = [p._version for p in self.params]
        if ptrs != self.ptrs or versions != self.versions or self.epoch != _ops.WEIGHT_EPOCH:
            self.generation += 1
        if ptrs != self.ptrs:  # first call, or a parameter's storage was swapped (load_state_dict(assign=True), .to())
            self.src = TensorList([p.detach() for p in self.params], F32, "weights")
            self.dst = TensorList(self.views, BF16, "arena")
            self.ptrs = ptrs
            self.index = {ptr: i for i, ptr in enumerate(ptrs)}
        cast_bf16(self.src, self.dst)
        self.versions = versions
        self.epoch = _ops.WEIGHT_EPOCH

    def lookup(self, t):
        i = self.index.get(t.data_ptr())
        if i is None or self.epoch != _ops.WEIGHT_EPOCH:
            return None
        p = self.params[i]
        if p.data_ptr() != t.data_ptr() or p._version != self.versions[i] or t.numel() != p.numel():
            return None
        v = self.views[i].view(t.shape)
        v._vtb_arena_gen = (self, self.generation)  # checked by ops.assert_weight_fresh in the backward pass
        return v


_ARENAS = []


def enable_weight_arena(module):
    """Attach a WeightArena to `module` (a top-level model already on the GPU) and refresh it before every forward.
    Returns the arena; `disable_weight_arena(module)` removes it."""
    arena = WeightArena(module)
    handle = module.register_forward_pre_hook(lambda m, args: arena.refresh())
    module._vtb_weight_arena = (arena, handle)
    _ARENAS.append(arena)
    _ops.WEIGHT_LOOKUP = _lookup
    return arena


def disable_weight_arena(module):
    arena, handle = module.__dict__.pop("_vtb_weight_arena")
    handle.remove()
    _ARENAS.remove(arena)
    if not _ARENAS:
        _ops.WEIGHT_LOOKUP = None


def _lookup(t):
    for arena in _ARENAS:
        v = arena.lookup(t)
        if v is not None:
            return v
    return None
