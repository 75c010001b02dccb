"""Drop-in for the reference's `train_util.py`: same names and call signatures.  `accuracy` (train_util.py:53-67) and
`accumulate` (train_util.py:70-84) run as single library launches (vtb_mix_loss in its accuracy-only mode, vtb_mt_ema);
the rest is host-side bookkeeping restated from the reference's behaviour."""
import math

import torch

from vtb200 import multi


def cosine_schedule(base, final, step, warmup=0, warmup_start=0):
    """train_util.py:6-22: linear warm-up from `warmup_start` to `base`, then half a cosine from `base` to `final`."""
    head = torch.linspace(warmup_start, base, warmup).tolist() if warmup > 0 else []
    n = step - warmup
    # the reference iterates over a torch.arange, so the cosine argument pi*i/n is rounded to float32 before math.cos
    arg = (math.pi * torch.arange(n) / max(n, 1)).tolist()
    tail = [final + 0.5 * (base - final) * (1 + math.cos(a)) for a in arg]
    return head + torch.tensor(tail).tolist()


def cancel_last_layer_grad(epoch, model, freeze):
    """train_util.py:25-31: drop the gradients of every parameter whose name contains "last" while epoch < freeze."""
    if epoch >= freeze:
        return
    for name, p in model.named_parameters():
        if "last" in name:
            p.grad = None


class Meter(object):
    """Running average (train_util.py:34-50)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


@torch.no_grad()
def accuracy(output, target, topk=(1,)):
    """precision@k in percent, one 0-dim tensor per k (train_util.py:53-67).  One launch covers k = 1 and one more k."""
    batch_size = target.size(0)
    logits = output if output.dtype == torch.float32 else output.float()
    hits = {}
    others = sorted({k for k in topk if k != 1}) or [1]
    for k in others:
        _, _, _, correct = multi.mix_loss(logits, target, want_loss=False, want_grad=False, want_correct=True, topk=k)
        hits[1], hits[k] = correct[0], correct[1]
    return [hits[k].float().mul_(100.0 / batch_size) for k in topk]


@torch.no_grad()
def accumulate(model1, model2, decay=0.99999, ema_bn=False):
    """model1 = model1*decay + model2*(1-decay) over the parameters matched by name (train_util.py:70-84)."""
    par1 = dict(model1.named_parameters())
    par2 = dict(model2.named_parameters())
    dst = [par1[k].detach() for k in par1.keys()]
    src = [par2[k].detach() for k in par1.keys()]
    if ema_bn:
        buf1 = dict(model1.named_buffers())
        buf2 = dict(model2.named_buffers())
        for k in buf1.keys():
            if "running_mean" in k or "running_var" in k:
                dst.append(buf1[k])
                src.append(buf2[k])
    multi.ema(dst, src, decay)


def add_weight_decay(named_parameters, weight_decay, check_skip_fn):
    """train_util.py:87-112: split trainable parameters into a no-decay and a decay group."""
    groups = {True: ([], []), False: ([], [])}
    for name, p in named_parameters:
        if not p.requires_grad:
            continue
        params, names = groups[bool(check_skip_fn(name, p))]
        params.append(p)
        names.append(name)
    (no_decay, no_decay_names), (decay, decay_names) = groups[True], groups[False]
    return (
        ({"params": no_decay, "weight_decay": 0.0, "no_decay": True}, {"params": decay, "weight_decay": weight_decay}),
        (no_decay_names, decay_names),
    )
