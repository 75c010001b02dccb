"""Pins for the oracle (oracle/restate.py): golden vectors generated from the real reference, and — in the
build container, where /root/reference exists — the reference modules themselves."""
import pytest
import torch

from conftest import load_golden
from oracle import ref_loader, restate as R

CASES = ["vit_tiny", "vit_multicrop", "swin_w2", "swin_w7", "pvt_tiny", "halo_w2", "halo_w7", "twins_w2", "twins_w7"]
FWD = {"vit": R.vit_forward, "swin": R.swin_forward, "pvt": R.pvt_forward, "halo": R.halo_forward,
       "twins": R.twins_forward}


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_golden_forward_and_grads(name):
    fx = load_golden(name)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in fx["state_dict"].items()}
    inp = fx["inputs"] if len(fx["inputs"]) > 1 else fx["inputs"][0]
    out = FWD[fx["family"]](sd, inp, **fx["oracle_kwargs"])
    assert out.shape == fx["output"].shape
    assert rel(out, fx["output"]) < 5e-6  # fp32 CPU, same maths, different op order
    (out * fx["probe"]).sum().backward()
    for k, g in fx["grads"].items():
        assert sd[k].grad is not None, k
        assert rel(sd[k].grad, g) < 2e-4, (k, rel(sd[k].grad, g))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", CASES)
def test_golden_reproduces_from_reference(name):
    """The committed fixture is what the unmodified reference computes today."""
    from oracle.make_golden import build

    fx = load_golden(name)
    ref = ref_loader.load()
    model = build(ref, fx["family"], fx["ctor"]).eval()
    model.load_state_dict(fx["state_dict"], strict=True)
    inp = [x.clone() for x in fx["inputs"]]
    with torch.no_grad():
        out = model(inp if len(inp) > 1 else inp[0])
    assert rel(out, fx["output"]) < 1e-6


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_integer_tables_match_reference_bit_exact():
    ref = ref_loader.load()
    for hs, shift in [(56, True), (28, True), (14, True), (7, True), (56, False), (7, False)]:
        a = ref.swin_transformer.MultiHeadedLocalAttention(16, 2, 8, (hs, hs), 7, shift)
        pos, mask = R.swin_tables(hs, hs, 7, shift)
        assert torch.equal(pos, a.pos)
        if shift:
            assert torch.equal(mask, a.local_mask)
    a = ref.halo_transformer.MultiHeadedHaloAttention(16, 2, 8, 7, 3)
    assert torch.equal(R.halo_pos_table(7, 3), a.pos)
    assert a.pos.unique().numel() == 253 - 0 or a.pos.max().item() + 1 == a.rel_pos.weight.shape[0]


def test_drop_path_scales_are_consumed_in_branch_order():
    fx = load_golden("vit_tiny")
    sd = fx["state_dict"]
    x = fx["inputs"][0]
    B = x.shape[0]
    ones = [torch.ones(B)] * 4
    out = R.vit_forward(sd, x, dp_scales=ones, **fx["oracle_kwargs"])
    assert rel(out, fx["output"]) < 5e-6
    zeros = [torch.zeros(B)] * 4
    out0 = R.vit_forward(sd, x, dp_scales=zeros, **fx["oracle_kwargs"])
    # every branch dropped: the residual stream is just the embedding -> differs from the full model
    assert rel(out0, fx["output"]) > 1e-3


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_reference_bf16_autocast_error_level():
    """What "matching the reference" can mean in bf16: the UNMODIFIED reference under its own training dtype flow
    (torch.autocast, train.py:273; here on CPU) differs from its own fp32 output by 3e-3 .. 1e-2 rel-L2 on the golden
    configurations, and at most a quarter of its logits are within the north star's rtol 1e-3.  The bf16 bars of the GPU
    suite (outputs 2e-2 rel-L2) sit just above this level; tests/golden/autocast_levels.json records it."""
    import json
    import os

    from conftest import GOLDEN
    from oracle.make_golden import build

    levels = json.load(open(os.path.join(GOLDEN, "autocast_levels.json")))["cases"]
    ref = ref_loader.load()
    for name in ("vit_tiny", "swin_w7", "pvt_tiny"):
        fx = load_golden(name)
        model = build(ref, fx["family"], fx["ctor"]).eval()
        model.load_state_dict(fx["state_dict"], strict=True)
        inp = [x.clone() for x in fx["inputs"]]
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            out = model(inp if len(inp) > 1 else inp[0]).float()
        err = rel(out, fx["output"])
        assert 1e-3 < err < 2e-2, (name, err)  # above the fp32 target, below the GPU suite's bf16 bar
        assert err == pytest.approx(levels[name]["rel_l2"], rel=0.5), (name, err, levels[name])
    assert all(v["rel_l2"] > 1e-3 and v["share_within_rtol_1e-3"] <= 0.5 for v in levels.values())


# ------------------------------------------------------------------------------------------------ DINO pins (a7, f1)
def _dino_fx():
    return load_golden("dino_ops")


def test_oracle_dino_head_matches_reference_golden():
    """R.dino_head (the checker of the DINOHead kernels) against outputs and autograd gradients of the reference's OWN
    DINOHead (models/vit.py:206-262; tests/golden/dino_ops.pt written by oracle/make_dino_golden.py)."""
    from oracle import restate as R

    for name, h in _dino_fx()["heads"].items():
        sd = {k: v.clone().requires_grad_(True) for k, v in h["state_dict"].items()}
        x = h["x"].clone().requires_grad_(True)
        out = R.dino_head(sd, x, pre="")
        assert rel(out, h["output"]) < 2e-6, (name, rel(out, h["output"]))
        (out * h["probe"]).sum().backward()
        assert rel(x.grad, h["dx"]) < 2e-5, name
        for k, g in h["grads"].items():
            if g is None:   # frozen weight_g (norm_last_layer=True)
                continue
            assert rel(sd[k].grad, g) < 2e-5, (name, k, rel(sd[k].grad, g))


def test_oracle_dino_loss_matches_reference_golden():
    """R.dino_loss against the reference's OWN DINOLoss (loss.py:89-152): value, student gradient, and the centre after
    one and two calls (update_center at world size 1), in the warm-up and the final temperature regime."""
    from oracle import restate as R

    for name, c in _dino_fx()["losses"].items():
        K, n_crop = c["ctor"][0], c["ctor"][1]
        s = c["student"].clone().requires_grad_(True)
        loss = R.dino_loss(s, c["teacher"], c["center0"], n_crop, 0.1, c["temperature"])
        assert abs(loss.item() - c["loss"].item()) < 2e-6 * abs(c["loss"].item()), name
        loss.backward()
        assert rel(s.grad, c["dstudent"]) < 2e-5, name
        center1 = c["center0"] * 0.9 + c["teacher"].sum(0, keepdim=True) / c["teacher"].shape[0] * 0.1   # loss.py:144-152
        assert rel(center1, c["center1"]) < 1e-6, name
        loss2 = R.dino_loss(c["student2"], c["teacher2"], center1, n_crop, 0.1, c["temperature"])
        assert abs(loss2.item() - c["loss2"].item()) < 2e-6 * abs(c["loss2"].item()), name


def test_dino_golden_reproduces_from_reference():
    """In the build container: the committed DINO fixture is what the live reference classes produce."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    ref = ref_loader.load()
    for name, h in _dino_fx()["heads"].items():
        head = ref.vit.DINOHead(**h["ctor"])
        head.load_state_dict(h["state_dict"])
        assert rel(head(h["x"]), h["output"]) < 1e-6, name


# ------------------------------------------------------------------------------------------- element dropout (train mode)
DROPOUT_CASES = ["vit", "vit_ff_only", "swin", "pvt"]


@pytest.mark.parametrize("name", DROPOUT_CASES)
def test_oracle_element_dropout_sites_match_reference_golden(name):
    """The restatement's nn.Dropout sites (layer.py:194, vit.py:60-61,146, pvt.py:141), replaying the keep masks the
    UNMODIFIED reference ran with in train mode (oracle/make_dropout_golden.py): same output, same gradients."""
    fx = load_golden("dropout_ops")[name]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in fx["state_dict"].items()}
    with R.element_dropout(fx["masks"], fx["sites"]):
        out = FWD[fx["family"]](sd, fx["input"], **fx["oracle_kwargs"])
    assert rel(out, fx["output"]) < 5e-6
    (out * fx["probe"]).sum().backward()
    for k, g in fx["grads"].items():
        assert rel(sd[k].grad, g) < 2e-4, (k, rel(sd[k].grad, g))
    # without the masks the oracle is the eval-mode forward: a different result (the fixture really dropped something)
    assert rel(FWD[fx["family"]](sd, fx["input"], **fx["oracle_kwargs"]), fx["output"]) > 1e-2


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", DROPOUT_CASES)
def test_dropout_golden_reproduces_from_reference_with_its_own_masks(name):
    """The fixture equals the unmodified reference in train mode with nn.Dropout's OWN masks under the recorded generator
    state — i.e. the recorded masks are the masks torch draws, in the order the reference draws them."""
    from oracle.make_golden import build

    fx = load_golden("dropout_ops")[name]
    ref = ref_loader.load()
    model = build(ref, fx["family"], fx["ctor"]).train()
    model.load_state_dict(fx["state_dict"], strict=True)
    torch.manual_seed(fx["seed"] + 7)
    out = model(fx["input"].clone())
    assert torch.equal(out.detach(), fx["output"])


def test_make_dropout_keep_draws_what_nn_dropout_draws():
    """vtb200.blocks.make_dropout_keep (the product's mask source) = the mask of nn.Dropout on a tensor of that shape /
    dtype under the same generator state; nothing is drawn in eval mode or at p = 0."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vision-transformers-pytorch_b200"))
    from vtb200.blocks import make_dropout_keep

    for dtype in (torch.float32, torch.bfloat16):
        x = torch.randn(3, 17, 40).to(dtype) + 3  # no exact zeros
        torch.manual_seed(11)
        want = torch.nn.Dropout(0.3).train()(x).ne(0)
        torch.manual_seed(11)
        keep, scale = make_dropout_keep(True, 0.3, x.shape, dtype, "cpu")
        assert keep.dtype == torch.bool and torch.equal(keep, want) and abs(scale - 1 / 0.7) < 1e-12
    state = torch.get_rng_state()
    assert make_dropout_keep(False, 0.3, (4, 4), torch.float32, "cpu") is None
    assert make_dropout_keep(True, 0.0, (4, 4), torch.float32, "cpu") is None
    assert torch.equal(torch.get_rng_state(), state)
    keep, scale = make_dropout_keep(True, 1.0, (4, 4), torch.float32, "cpu")
    assert not keep.any() and scale == 0.0
