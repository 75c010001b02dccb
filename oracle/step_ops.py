"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — CPU restatement, in numpy float32, of the step-side loops the
multi-tensor kernels replace.  Each function follows the reference line it cites; where the arithmetic lives in torch
(AdamW, clip_grad_norm_) it restates torch 2.11's published single-tensor algorithm.  Pinned by
tests/golden/step_ops.pt, which oracle/make_step_golden.py generates by running the reference's own functions
(optimizer.py, train_util.py, loss.py) and torch.optim.AdamW / nn.utils.clip_grad_norm_ on the same inputs.
"""
import math

import numpy as np

F = np.float32


def ema(dst, src, decay):
    """train_util.py:77 / train_dino.py:261: `p1.mul_(decay).add_(p2, alpha=1 - decay)`, per tensor."""
    d, a = F(decay), F(1.0 - decay)
    return [(x.astype(F) * d + y.astype(F) * a).astype(F) for x, y in zip(dst, src)]


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (train.py:294): total = ||(||g_i||)_i||_2; coef = clamp(max_norm / (total + 1e-6),
    max=1); every g_i *= coef.  -> (clipped grads, total norm)."""
    total = F(math.sqrt(sum(float(np.sum(g.astype(np.float64) ** 2)) for g in grads)))
    coef = F(max_norm) / (total + F(1e-6))
    coef = F(1.0) if coef > 1 else coef
    return [(g * coef).astype(F) for g in grads], total


def unitwise_norm(x):
    """optimizer.py:4-9: the norm of the whole tensor (ndim <= 1) or of each slice along dim 0, kept broadcastable."""
    if x.ndim <= 1:
        return np.sqrt(np.sum(x.astype(np.float64) ** 2)).astype(F)
    axes = tuple(range(1, x.ndim))
    return np.sqrt(np.sum(x.astype(np.float64) ** 2, axis=axes, keepdims=True)).astype(F)


def adaptive_grad_clip(params, grads, clipping=0.01, eps=1e-3):
    """optimizer.py:12-26."""
    out = []
    for p, g in zip(params, grads):
        max_norm = np.maximum(unitwise_norm(p), F(eps)) * F(clipping)
        g_norm = unitwise_norm(g)
        clipped = g * (max_norm / np.maximum(g_norm, F(1e-6)))
        out.append(np.where(g_norm < max_norm, g, clipped).astype(F))
    return out


def adamw_step(p, g, m, v, *, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    """torch/optim/adamw.py `_single_tensor_adamw` (the `adamw` of config/swin-transformer-s.conf:39-42):
    decoupled decay, lerp of the first moment, second moment, bias corrections, addcdiv.  -> (p, m, v)."""
    g = (g * F(grad_scale)).astype(F)
    p = (p * F(1.0 - lr * weight_decay)).astype(F)
    m = (m + F(1.0 - beta1) * (g - m)).astype(F)
    v = (v * F(beta2) + F(1.0 - beta2) * g * g).astype(F)
    step_size = F(lr / (1.0 - beta1 ** step))
    bc2_sqrt = F(math.sqrt(1.0 - beta2 ** step))
    denom = (np.sqrt(v) / bc2_sqrt + F(eps)).astype(F)
    p = (p - step_size * (m / denom)).astype(F)
    return p, m, v


def mix_loss(logits, target1, target2, inter, eps=0.0, reduction="mean"):
    """loss.py:60-86.  -> (loss, d loss / d logits) in float64 arithmetic on float32 inputs."""
    x = logits.astype(np.float64)
    B, n = x.shape
    logp = x - x.max(-1, keepdims=True)
    logp = logp - np.log(np.exp(logp).sum(-1, keepdims=True))

    def true(t):
        d = np.full((B, n), eps / n)
        d[np.arange(B), t] = 1 - eps + eps / n
        return d

    w = np.asarray(inter, dtype=np.float64).reshape(-1, 1)
    t = w * true(target1) + (1 - w) * true(target2)
    with np.errstate(divide="ignore", invalid="ignore"):
        kl = np.where(t > 0, t * np.log(t), 0.0) - t * logp  # F.kl_div: xlogy(t, t) - t * input
    rows = kl.sum(-1)
    grad = np.exp(logp) * t.sum(-1, keepdims=True) - t
    if reduction == "none":
        return rows, grad
    if reduction == "mean":
        return rows.sum() / B, grad / B
    return rows.sum(), grad


def accuracy(logits, target, topk=(1,)):
    """train_util.py:53-67: percentage of rows whose target is among the k largest logits."""
    B = logits.shape[0]
    rank = (logits > logits[np.arange(B), target][:, None]).sum(-1)
    return [float((rank < k).sum()) * 100.0 / B for k in topk]
