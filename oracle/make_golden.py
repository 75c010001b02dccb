"""TEST INFRASTRUCTURE ONLY — generates tests/golden/* from the UNMODIFIED reference (build container
only; needs /root/reference).  Run:  python oracle/make_golden.py

For every transformer family a tiny configuration (kernel-legal sizes: head dim 32, channel counts % 8) is
built from the reference's own classes, parameters are re-randomised with oracle.restate.randomize_ (seeded;
exercises rel_pos / LayerNorm affines), and the reference's forward output plus the autograd
gradients of loss = sum(out * probe) w.r.t. every parameter are stored.  HaloTransformer's own backward
raises (in-place residual, halo_transformer.py:147-148), so its gradients come from the out-of-place
restatement in oracle/restate.py and the fixture says so.

Also writes tests/golden/structure.json: state_dict key -> shape/dtype snapshots and parameter counts of the
full-size BASELINE configurations, and SHA-256 of the Swin/Halo integer buffers (structural pins, §8c).
"""
import hashlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, restate as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (family, ctor kwargs / args, input shape(s), oracle kwargs)
    "vit_tiny": ("vit", dict(head=None, image_size=32, window_size=8, depth=2, dim=64, n_head=2, dim_ff=128,
                             dropout=0., drop_attn=0., drop_ff=0., drop_path=0.), [(2, 3, 32, 32)],
                 dict(patch=8, depth=2, heads=2)),
    "vit_multicrop": ("vit", dict(head=None, image_size=32, window_size=8, depth=2, dim=64, n_head=2, dim_ff=128,
                                  dropout=0., drop_attn=0., drop_ff=0., drop_path=0.),
                      [(2, 3, 32, 32), (2, 3, 32, 32), (2, 3, 16, 16), (2, 3, 16, 16)],
                      dict(patch=8, depth=2, heads=2)),
    "swin_w2": ("swin", dict(image_size=(64, 64), n_class=10, depths=(2, 2, 2, 2), dims=(32, 32, 64, 64), dim_head=32,
                             n_heads=(1, 1, 2, 2), dim_ffs=(64, 64, 128, 128), window_size=2), [(2, 3, 64, 64)],
                dict(depths=(2, 2, 2, 2), n_heads=(1, 1, 2, 2), dim_head=32, window=2)),
    "swin_w7": ("swin", dict(image_size=(224, 224), n_class=10, depths=(2, 2, 2, 2), dims=(32, 32, 32, 32), dim_head=32,
                             n_heads=(1, 1, 1, 1), dim_ffs=(32, 32, 32, 32), window_size=7), [(1, 3, 224, 224)],
                dict(depths=(2, 2, 2, 2), n_heads=(1, 1, 1, 1), dim_head=32, window=7)),
    "pvt_tiny": ("pvt", dict(image_size=64, n_class=10, in_dim=3, depths=(1, 1, 2, 1), patch_embed_dims=(32, 64, 64, 64),
                             n_heads=(1, 2, 2, 2), dim_ffs=(64, 128, 128, 128), reductions=(8, 4, 2, 1)),
                 [(2, 3, 64, 64)], dict(depths=(1, 1, 2, 1), n_heads=(1, 2, 2, 2), reductions=(8, 4, 2, 1))),
    "halo_w2": ("halo", dict(image_size=(64, 64), n_class=10, depths=(1, 1, 2, 1), dims=(32, 32, 64, 64), dim_head=32,
                             n_heads=(1, 1, 2, 2), dim_ffs=(64, 64, 128, 128), window_size=2, halo_size=1),
                [(2, 3, 64, 64)], dict(depths=(1, 1, 2, 1), n_heads=(1, 1, 2, 2), dim_head=32, window=2, halo=1)),
    "halo_w7": ("halo", dict(image_size=(224, 224), n_class=10, depths=(1, 1, 1, 1), dims=(32, 32, 32, 32), dim_head=32,
                             n_heads=(1, 1, 1, 1), dim_ffs=(32, 32, 32, 32), window_size=7, halo_size=3),
                [(1, 3, 224, 224)], dict(depths=(1, 1, 1, 1), n_heads=(1, 1, 1, 1), dim_head=32, window=7, halo=3)),
    "twins_w2": ("twins", dict(n_class=10, depths=(1, 2, 1, 1), dims=(32, 32, 64, 64), dim_head=32, n_heads=(1, 1, 2, 2),
                               dim_ffs=(64, 64, 128, 128), window_size=2), [(2, 3, 64, 64)],
                 dict(depths=(1, 2, 1, 1), n_heads=(1, 1, 2, 2), dim_head=32, window=2)),
    "twins_w7": ("twins", dict(n_class=10, depths=(1, 1, 1, 1), dims=(32, 32, 32, 32), dim_head=32, n_heads=(1, 1, 1, 1),
                               dim_ffs=(32, 32, 32, 32), window_size=7), [(1, 3, 224, 224)],
                 dict(depths=(1, 1, 1, 1), n_heads=(1, 1, 1, 1), dim_head=32, window=7)),
}

FULL = {
    "vit_b16": ("vit", dict(head=None, image_size=224, window_size=16, depth=12, dim=768, n_head=12, dim_ff=3072,
                            dropout=0., drop_attn=0., drop_ff=0., drop_path=0.)),
    "vit_tiny16": ("vit", dict(head=None, image_size=224, window_size=16, depth=12, dim=192, n_head=3, dim_ff=768,
                               dropout=0., drop_attn=0., drop_ff=0., drop_path=0.)),
    "swin_s": ("swin", dict(image_size=(224, 224), n_class=1000, depths=(2, 2, 18, 2), dims=(96, 192, 384, 768),
                            dim_head=32, n_heads=(3, 6, 12, 24), dim_ffs=(384, 768, 1536, 3072), window_size=7,
                            drop_path=0.3)),
    "pvt_small": ("pvt", dict(image_size=224, n_class=1000, in_dim=3, depths=(3, 4, 6, 3),
                              patch_embed_dims=(64, 128, 320, 512), n_heads=(1, 2, 5, 8),
                              dim_ffs=(512, 1024, 1280, 2048), reductions=(8, 4, 2, 1))),
    "halo_t": ("halo", dict(image_size=(224, 224), n_class=1000, depths=(2, 2, 6, 2), dims=(96, 192, 384, 768),
                            dim_head=32, n_heads=(3, 6, 12, 24), dim_ffs=(384, 768, 1536, 3072), window_size=7,
                            halo_size=3)),
    "twins_s": ("twins", dict(n_class=1000, depths=(2, 2, 10, 4), dims=(64, 128, 256, 512), dim_head=32,
                              n_heads=(2, 4, 8, 16), dim_ffs=(256, 512, 1024, 2048), window_size=7, drop_path=0.2)),
}


def build(ref, family, kw):
    if family == "vit":
        return ref.vit.VisionTransformer(**kw)
    if family == "swin":
        return ref.swin_transformer.SwinTransformer(**kw)
    if family == "pvt":
        return ref.pvt.PyramidVisionTransformer(**kw)
    if family == "halo":
        return ref.halo_transformer.HaloTransformer(**kw)
    if family == "twins":
        return ref.twins.TwinsSVT(**kw)
    raise KeyError(family)


ORACLE_FWD = {"vit": R.vit_forward, "swin": R.swin_forward, "pvt": R.pvt_forward, "halo": R.halo_forward,
              "twins": R.twins_forward}


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    for seed, (name, (family, kw, in_shapes, okw)) in enumerate(CASES.items(), start=100):
        if os.path.exists(os.path.join(OUT, name + ".pt")) and "--all" not in sys.argv:
            print(f"{name}: exists, kept (pass --all to regenerate)")
            continue
        torch.manual_seed(seed)
        model = R.randomize_(build(ref, family, kw).eval(), seed)
        g = torch.Generator().manual_seed(seed + 1000)
        xs = [torch.randn(s, generator=g) for s in in_shapes]
        inp = xs if len(xs) > 1 else xs[0]
        grads_from = "reference"
        params = dict(model.named_parameters())
        if family == "halo":
            with torch.no_grad():
                out = model(inp.clone())
            sd = {k: (v.detach().clone().requires_grad_(v.is_floating_point())) for k, v in model.state_dict().items()}
            out2 = ORACLE_FWD[family](sd, inp, **okw)
            assert torch.allclose(out, out2, atol=1e-5, rtol=1e-4), (out - out2).abs().max()
            probe = torch.randn(out.shape, generator=g)
            (out2 * probe).sum().backward()
            grads = {k: sd[k].grad.clone() for k in params}
            grads_from = "oracle restatement (reference backward raises: halo_transformer.py:147-148)"
        else:
            out = model(inp)
            probe = torch.randn(out.shape, generator=g)
            (out * probe).sum().backward()
            grads = {k: p.grad.clone() for k, p in params.items()}
        fx = dict(family=family, ctor=kw, oracle_kwargs=okw, inputs=xs, probe=probe, output=out.detach(),
                  state_dict={k: v.detach().clone() for k, v in model.state_dict().items()}, grads=grads,
                  grads_from=grads_from, seed=seed, torch=torch.__version__)
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"{name}: out {tuple(out.shape)} params {sum(p.numel() for p in params.values())} "
              f"-> {os.path.getsize(path) / 1e6:.2f} MB")

    structure = {}
    for name, (family, kw) in FULL.items():
        model = build(ref, family, kw)
        sd = model.state_dict()
        entry = dict(n_params=sum(p.numel() for p in model.parameters()), n_entries=len(sd),
                     param_order=[k for k, _ in model.named_parameters()],
                     keys={k: [list(v.shape), str(v.dtype)] for k, v in sd.items()})
        h = hashlib.sha256()
        nbytes = 0
        for k, v in model.named_buffers():
            h.update(k.encode())
            h.update(v.to(torch.int64).numpy().tobytes())
            nbytes += v.numel() * v.element_size()
        entry["buffers_sha256"], entry["buffer_bytes"] = h.hexdigest(), nbytes
        structure[name] = entry
        print(name, entry["n_params"], entry["n_entries"], nbytes)
    with open(os.path.join(OUT, "structure.json"), "w") as f:
        json.dump(structure, f)


if __name__ == "__main__":
    main()
