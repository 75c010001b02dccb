"""Generates tests/golden/step_ops.pt by running the REFERENCE's own step-side functions — optimizer.py
(adaptive_grad_clip), train_util.py (accumulate, accuracy), loss.py (MixLoss) — and the torch functions its training
loop calls (torch.optim.AdamW, nn.utils.clip_grad_norm_; train.py:285-299) on small seeded inputs.  Build container only
(needs /root/reference):  python oracle/make_step_golden.py
TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.
"""
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader  # noqa: E402

SHAPES = [(7, 5), (13,), (3, 4, 2, 2), (1, 1, 6), (1,), (40, 33)]


def tensors(seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) * scale for s in SHAPES]


class Bag(nn.Module):
    def __init__(self, ts):
        super().__init__()
        self.ps = nn.ParameterList([nn.Parameter(t.clone()) for t in ts])


def main():
    ref_opt = ref_loader.load_reference_module("optimizer")
    ref_tu = ref_loader.load_reference_module("train_util")
    ref_loss = ref_loader.load_reference_module("loss")
    out = {"shapes": SHAPES}

    # EMA: train_util.accumulate on two parameter bags
    a, b = Bag(tensors(1)), Bag(tensors(2))
    out["ema_dst"], out["ema_src"], out["ema_decay"] = tensors(1), tensors(2), 0.996
    ref_tu.accumulate(a, b, decay=0.996)
    out["ema_out"] = [p.detach().clone() for p in a.ps]

    # clip_grad_norm_ (clipping and not clipping)
    for tag, max_norm in (("clip", 3.0), ("noclip", 1e4)):
        bag = Bag(tensors(3))
        for p, g in zip(bag.ps, tensors(4)):
            p.grad = g.clone()
        total = nn.utils.clip_grad_norm_(list(bag.parameters()), max_norm)
        out[f"{tag}_max_norm"], out[f"{tag}_total"] = max_norm, total.clone()
        out[f"{tag}_out"] = [p.grad.clone() for p in bag.ps]
    out["clip_grads"] = tensors(4)

    # adaptive_grad_clip: gradients 30x larger than the weights so that most units clip, some do not
    bag = Bag(tensors(5))
    grads = tensors(6, scale=0.02)
    grads[0][2] *= 1e-3  # one unit well below its threshold
    for p, g in zip(bag.ps, grads):
        p.grad = g.clone()
    ref_opt.adaptive_grad_clip(list(bag.parameters()), clipping=0.01, eps=1e-3)
    out["agc_params"], out["agc_grads"] = tensors(5), grads
    out["agc_out"] = [p.grad.clone() for p in bag.ps]

    # AdamW: three steps, two groups (decay / no decay), the gradients change per step
    bag = Bag(tensors(7))
    ps = list(bag.parameters())
    hp = dict(lr=2.5e-4, betas=(0.9, 0.999), eps=1e-8)
    opt = torch.optim.AdamW([{"params": ps[:3], "weight_decay": 0.05}, {"params": ps[3:], "weight_decay": 0.0}], **hp)
    out["adamw_params"], out["adamw_hp"], out["adamw_wd"] = tensors(7), hp, [0.05] * 3 + [0.0] * (len(ps) - 3)
    out["adamw_grads"], out["adamw_out"] = [], []
    for step in range(3):
        gs = tensors(10 + step)
        for p, g in zip(ps, gs):
            p.grad = g.clone()
        opt.step()
        out["adamw_grads"].append(gs)
        out["adamw_out"].append([p.detach().clone() for p in ps])
    out["adamw_exp_avg"] = [opt.state[p]["exp_avg"].clone() for p in ps]
    out["adamw_exp_avg_sq"] = [opt.state[p]["exp_avg_sq"].clone() for p in ps]

    # MixLoss (+ gradient through autograd) and accuracy
    g = torch.Generator().manual_seed(20)
    B, n = 12, 37
    logits = (torch.randn(B, n, generator=g) * 3).requires_grad_()
    t1 = torch.randint(0, n, (B,), generator=g)
    t2 = torch.randint(0, n, (B,), generator=g)
    t2[:3] = t1[:3]  # equal labels (no mixing partner)
    inter = torch.rand(B, generator=g)
    inter[3] = 1.0
    out["mix_logits"], out["mix_t1"], out["mix_t2"], out["mix_inter"] = logits.detach().clone(), t1, t2, inter
    for eps in (0.1, 0.0):
        for red in ("mean", "none", "sum"):
            logits.grad = None
            loss = ref_loss.MixLoss(eps=eps, reduction=red)(logits, t1, t2, inter)
            loss.sum().backward()
            out[f"mix_{eps}_{red}_loss"], out[f"mix_{eps}_{red}_grad"] = loss.detach().clone(), logits.grad.clone()
    out["acc_1_5"] = [t.clone() for t in ref_tu.accuracy(logits.detach(), t1, topk=(1, 5))]
    out["acc_1_3_5"] = [t.clone() for t in ref_tu.accuracy(logits.detach(), t1, topk=(1, 3, 5))]

    # host-side helpers
    out["cosine"] = ref_tu.cosine_schedule(1.0, 0.1, 12, warmup=4, warmup_start=0.0)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "step_ops.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
