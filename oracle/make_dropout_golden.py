"""TEST INFRASTRUCTURE ONLY — generates tests/golden/dropout_ops.pt from the UNMODIFIED reference in TRAIN mode with
element dropout switched on (build container only; needs /root/reference).  Run:  python oracle/make_dropout_golden.py

nn.Dropout sites of the path: the Dropout inside PositionwiseFeedForward (layer.py:194,
`drop_ff`), on both branch outputs and the token embedding of ViT (vit.py:56-61,102,146, `dropout`), on PVT's patch
embedding (pvt.py:127,141).  The reference's modules run unchanged; only torch.nn.functional.dropout is wrapped so that
each call's keep mask is RECORDED: the wrapper draws the mask with the original F.dropout on a tensor of ones of the
input's shape / dtype — the same generator call the module would have made (tests/test_oracle.py checks on the live
reference that this reproduces nn.Dropout's own mask bit for bit) — and multiplies.  Stored per case: state_dict, input,
the masks in call order, output, every parameter gradient of sum(out * probe).  The product replays the masks
(tests/test_models_gpu.py::test_element_dropout_matches_reference_golden); drop_attn stays 0 (rejected by the product).
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, restate as R  # noqa: E402
from oracle.make_golden import build  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "dropout_ops.pt")

CASES = {
    # name: (family, ctor kwargs, input shape, oracle kwargs, oracle dropout sites)
    "vit": ("vit", dict(head=None, image_size=32, window_size=8, depth=2, dim=64, n_head=2, dim_ff=128, dropout=0.1,
                        drop_attn=0., drop_ff=0.25, drop_path=0.), (4, 3, 32, 32), dict(patch=8, depth=2, heads=2),
            ("pos", "branch", "ffn")),
    "vit_ff_only": ("vit", dict(head=None, image_size=32, window_size=8, depth=2, dim=64, n_head=2, dim_ff=128, dropout=0.,
                                drop_attn=0., drop_ff=0.5, drop_path=0.), (4, 3, 32, 32), dict(patch=8, depth=2, heads=2),
                    ("ffn",)),
    "swin": ("swin", dict(image_size=(64, 64), n_class=10, depths=(2, 2, 2, 2), dims=(32, 32, 64, 64), dim_head=32,
                          n_heads=(1, 1, 2, 2), dim_ffs=(64, 64, 128, 128), window_size=2, drop_ff=0.2), (2, 3, 64, 64),
             dict(depths=(2, 2, 2, 2), n_heads=(1, 1, 2, 2), dim_head=32, window=2), ("ffn",)),
    "pvt": ("pvt", dict(image_size=64, n_class=10, in_dim=3, depths=(1, 1, 2, 1), patch_embed_dims=(32, 64, 64, 64),
                        n_heads=(1, 2, 2, 2), dim_ffs=(64, 128, 128, 128), reductions=(8, 4, 2, 1), drop_ff=0.2),
            (2, 3, 64, 64), dict(depths=(1, 1, 2, 1), n_heads=(1, 2, 2, 2), reductions=(8, 4, 2, 1)), ("patch", "ffn")),
}


class RecordedDropout:
    """Context: F.dropout records (keep mask, 1 / (1 - p)) per call that actually drops."""

    def __init__(self):
        self.masks = []

    def __enter__(self):
        self.orig = F.dropout
        rec = self

        def dropout(input, p=0.5, training=True, inplace=False):
            if not training or p == 0:
                return input
            m = rec.orig(torch.ones_like(input), p, True)
            rec.masks.append((m.ne(0), 1.0 / (1.0 - p)))
            return input * m

        F.dropout = dropout
        return self

    def __exit__(self, *exc):
        F.dropout = self.orig


def run_case(ref, family, kw, in_shape, seed):
    torch.manual_seed(seed)
    model = R.randomize_(build(ref, family, kw), seed).train()
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(in_shape, generator=g)
    torch.manual_seed(seed + 7)
    with RecordedDropout() as rec:
        out = model(x)
    probe = torch.randn(out.shape, generator=g)
    (out * probe).sum().backward()
    return model, x, rec.masks, out.detach(), probe, {k: p.grad.clone() for k, p in model.named_parameters()}


def main():
    ref = ref_loader.load()
    fx = {}
    for seed, (name, (family, kw, in_shape, okw, sites)) in enumerate(CASES.items(), start=300):
        model, x, masks, out, probe, grads = run_case(ref, family, kw, in_shape, seed)
        # the same run with nn.Dropout's own masks (F.dropout untouched) under the same generator state must agree
        torch.manual_seed(seed)
        model2 = R.randomize_(build(ref, family, kw), seed).train()
        torch.manual_seed(seed + 7)
        out2 = model2(x)
        assert torch.equal(out, out2.detach()), (name, (out - out2).abs().max())
        fx[name] = dict(family=family, ctor=kw, oracle_kwargs=okw, sites=sites, input=x, probe=probe, output=out,
                        masks=[(m.clone(), s) for m, s in masks], grads=grads, seed=seed,
                        state_dict={k: v.detach().clone() for k, v in model.state_dict().items()})
        print(f"{name}: {len(masks)} dropout calls, kept fractions "
              + " ".join(f"{m.float().mean():.2f}" for m, _ in masks[:6]) + " ...")
    fx["torch"] = torch.__version__
    torch.save(fx, OUT)
    print(f"-> {OUT} {os.path.getsize(OUT) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
