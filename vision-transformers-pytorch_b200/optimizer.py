"""Drop-in for the reference's `optimizer.py` (adaptive_grad_clip, optimizer.py:12-26) plus the two other per-parameter
loops of the reference's update step — `nn.utils.clip_grad_norm_` (train.py:294, train_dino.py:243) and the `adamw`
optimizer its configs select (config/swin-transformer-s.conf:39-42, config/dino_deit-s-16.conf:52-55; tensorfn maps the
name to torch.optim.AdamW) — each as one multi-tensor launch over the parameter list (vtb_mt_agc / vtb_mt_grad_norm /
vtb_mt_scale / vtb_mt_adamw).  Nothing here synchronises with the host.
"""
import torch

from vtb200 import multi


@torch.no_grad()
def adaptive_grad_clip(parameters, clipping=0.01, eps=1e-3):
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]

    params = [p for p in parameters if p.grad is not None]
    if params:
        multi.agc([p.detach() for p in params], [p.grad for p in params], clipping, eps)


@torch.no_grad()
def clip_grad_norm_(parameters, max_norm, norm_type=2.0, defer_to=None):
    """torch.nn.utils.clip_grad_norm_ for the 2-norm.  Returns the total norm as a 0-dim device tensor (no sync).
    `defer_to`: an `AdamW` of this module — the clip coefficient is then folded into its next step() instead of a
    separate rescaling pass over the gradients."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("vtb200 clip_grad_norm_: only norm_type=2 is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros((), device="cuda")
    out = multi.grad_norm(grads, max_norm)
    if defer_to is not None:
        defer_to.grad_scale = out[1:2]
    else:
        multi.scale(grads, out[1:2])
    return out[0]


class AdamW(torch.optim.Optimizer):
    """torch.optim.AdamW (decoupled weight decay) with the same constructor, param_groups and state_dict layout
    (`step`, `exp_avg`, `exp_avg_sq` per parameter); one launch per parameter group."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False):
        if amsgrad or maximize:
            raise NotImplementedError("vtb200 AdamW: amsgrad / maximize are not implemented")
        if lr < 0 or eps < 0 or weight_decay < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("vtb200 AdamW: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False))
        self.grad_scale = None  # f32 device scalar set by clip_grad_norm_(..., defer_to=self); consumed by one step

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            buckets = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("AdamW does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                buckets.setdefault(int(st["step"]), []).append((p, st))
            for step, items in buckets.items():
                multi.adamw([p.detach() for p, _ in items], [p.grad for p, _ in items],
                            [st["exp_avg"] for _, st in items], [st["exp_avg_sq"] for _, st in items],
                            lr=float(group["lr"]), beta1=group["betas"][0], beta2=group["betas"][1], eps=group["eps"],
                            weight_decay=group["weight_decay"], step=step, grad_scale=self.grad_scale)
        self.grad_scale = None
        return loss
