"""TEST INFRASTRUCTURE ONLY — tests/golden/dino_ops.pt from the UNMODIFIED reference (build container only).

Pins the two DINO pieces of the hot-path table whose oracle had no reference anchor (a7, f1):
  * `DINOHead` (models/vit.py:206-262): forward output and the autograd gradients of every parameter and of the input for
    three configurations (depth 3 / norm_last_layer False as in config/dino_deit-s-16.conf, depth 1 with the frozen
    weight_g, depth 2), parameters re-randomised so that weight_g != 1 and the biases are non-zero;
  * `DINOLoss` (loss.py:89-152): loss value, gradient w.r.t. the student logits, and the `center` buffer after one and
    two calls (update_center runs on a single-process gloo group: world size 1, the all-reduce is the identity), for a
    non-zero starting centre, both temperature regimes (warm-up epoch 0 and the final temperature) and 4 / 10 crops.
Run:  python oracle/make_dino_golden.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "dino_ops.pt")

HEADS = {
    "conf": dict(in_dim=64, out_dim=512, use_bn=False, norm_last_layer=False, depth=3, dim_ff=128, dim_bottleneck=32),
    "depth1_frozen_g": dict(in_dim=48, out_dim=256, use_bn=False, norm_last_layer=True, depth=1, dim_ff=128, dim_bottleneck=32),
    "depth2": dict(in_dim=32, out_dim=384, use_bn=False, norm_last_layer=False, depth=2, dim_ff=64, dim_bottleneck=16),
}
LOSSES = {
    # name: (out_dim, n_crop, warmup_T, T, warmup_epochs, n_epoch, epoch, batch per crop)
    "conf_like": (512, 10, 0.04, 0.04, 0, 10, 3, 4),
    "warmup": (256, 4, 0.04, 0.07, 5, 10, 0, 3),
    "late": (256, 4, 0.04, 0.07, 5, 10, 8, 3),
}


def main():
    ref = ref_loader.load()
    ref_loss = ref_loader.load_reference_module("loss")
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    fx = {"heads": {}, "losses": {}, "torch": torch.__version__}
    for seed, (name, kw) in enumerate(HEADS.items(), start=300):
        torch.manual_seed(seed)
        head = ref.vit.DINOHead(**kw)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for k, p in head.named_parameters():
                if k.endswith("weight_g"):
                    p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
                elif k.endswith("bias"):
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.05 * torch.randn(p.shape, generator=g))
        x = torch.randn(12, kw["in_dim"], generator=g, requires_grad=True)
        out = head(x)
        probe = torch.randn(out.shape, generator=g)
        (out * probe).sum().backward()
        fx["heads"][name] = dict(ctor=kw, state_dict={k: v.detach().clone() for k, v in head.state_dict().items()}, x=x.detach(),
                                 probe=probe, output=out.detach(), dx=x.grad.clone(),
                                 grads={k: (p.grad.clone() if p.grad is not None else None) for k, p in head.named_parameters()})
        print(f"head {name}: out {tuple(out.shape)}")
    for seed, (name, (K, n_crop, wt, tt, we, ne, epoch, bpc)) in enumerate(LOSSES.items(), start=400):
        g = torch.Generator().manual_seed(seed)
        mod = ref_loss.DINOLoss(K, n_crop, wt, tt, we, ne)
        center0 = 0.3 * torch.randn(1, K, generator=g)
        mod.center.copy_(center0)
        student = (2.0 * torch.randn(n_crop * bpc, K, generator=g)).requires_grad_(True)
        teacher = 2.0 * torch.randn(2 * bpc, K, generator=g)
        loss1 = mod(student, teacher, epoch)
        loss1.backward()
        center1 = mod.center.clone()
        student2 = 2.0 * torch.randn(n_crop * bpc, K, generator=g)
        teacher2 = 2.0 * torch.randn(2 * bpc, K, generator=g)
        loss2 = mod(student2, teacher2, epoch)   # sees the centre updated by the first call
        fx["losses"][name] = dict(ctor=(K, n_crop, wt, tt, we, ne), epoch=epoch, center0=center0, student=student.detach(),
                                  teacher=teacher, loss=loss1.detach(), dstudent=student.grad.clone(), center1=center1,
                                  student2=student2, teacher2=teacher2, loss2=loss2.detach(), center2=mod.center.clone(),
                                  temperature=mod.teacher_temperature_schedule[epoch])
        print(f"loss {name}: {loss1.item():.6f} -> {loss2.item():.6f}")
    torch.save(fx, OUT)
    print(f"-> {OUT} ({os.path.getsize(OUT) / 1e6:.2f} MB)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
