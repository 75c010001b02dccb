"""Device input path: MixDataset mixup / cutmix, ToTensor + Normalize, RandomErasing as one kernel.

CPU: (1) the numpy oracle (oracle/input_ops.py) against tests/golden/input_ops.pt — batches produced by the REFERENCE's
own mix_dataset.MixDataset / transforms.RandomErasing with torchvision and PIL (oracle/make_input_golden.py) — with the
decisions re-drawn by the product's host sampler from the same seed: images bit-exact (1 ulp for tensor mixup), labels and
ratios equal; (2) Philox against the Random123 known answers, the blend against PIL over all 65 536 byte pairs;
(3) the kernel SOURCE compiled for the host (tests/kernel_emulation.py) against the oracle: index / byte arithmetic.
GPU: the CUDA kernel through the C-ABI against the same golden vectors and the oracle (vector and scalar variants, every
mode), noise statistics, and linearity / idempotence properties at the BASELINE size (256 x 224 x 224).
(The file sorts last on purpose: this kernel was written after the round's GPU budget was spent.)
"""
import ctypes as C
import functools
import hashlib
import os
import random

import numpy as np
import pytest
import torch

from conftest import load_golden


@pytest.fixture(scope="module")
def G():
    return load_golden("input_ops")


def redraw(case, noise_seed=0):
    """Decisions + table for a golden case from the product's sampler, seeded like the reference run."""
    import device_input as D

    s = D.MixSampler(case["mixup"], case["cutmix"], case["erasing"], case["mix_before_aug"],
                     rng=random.Random(case["seed"]), noise_seed=noise_seed)
    ds = [s.sample(i, case["n"], case["H"], case["W"]) for i in range(case["n"])]
    slots = {i: i for i in range(case["n"])}
    return ds, D.pack_table(ds, slots, case["mix_before_aug"], erase_mode="const")


def ulp_diff(a, b):
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia, ib = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia), np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib).max(initial=0)


def check_against_golden(case, got, name):
    want = case["img"].numpy()
    assert got.shape == want.shape
    if "tensor" in name and case["mixup"] > 0:
        assert ulp_diff(got, want) <= 1, name  # ATen's add_(x, alpha) may or may not be an FMA
    else:
        assert np.array_equal(got.view(np.int32), want.view(np.int32)), name


# ------------------------------------------------------------------------------------------------ CPU: sampler + oracle
def test_sampler_redraws_reference_decisions(G):
    seen = set()
    for name, case in G["cases"].items():
        ds, _ = redraw(case)
        labels = [i * 7 % 5 for i in range(case["n"])]
        assert [labels[d.index] for d in ds] == case["label1"], name
        assert [labels[d.partner] for d in ds] == case["label2"], name
        assert [float(d.ratio) for d in ds] == case["ratio"], name  # same float arithmetic, exact
        seen |= {d.mode for d in ds}
        if case["erasing"] > 0:
            assert any(d.erase_a[2] > 0 for d in ds), name  # the fixture really erases
    assert seen == {0, 1, 2}


def test_rand_bbox_matches_reference(G):
    import device_input as D

    rng = random.Random(5)
    for size, ratio, want in G["bbox"]:
        assert D.rand_bbox(size, ratio, rng) == tuple(want)


def test_oracle_matches_reference_batches(G):
    from oracle import input_ops as O

    for name, case in G["cases"].items():
        _, table = redraw(case)
        check_against_golden(case, O.input_batch(case["u8"].numpy(), table, G["mean"], G["std"]), name)


def test_sampler_and_oracle_match_live_reference_on_random_configs():
    """Build container only: 24 seeded random loader configurations (sizes, mixup / cutmix / erasing strengths, both
    `mix_before_aug` orders) through the reference's OWN MixDataset + RandomErasing + torchvision + PIL, against the host
    sampler + numpy oracle: labels and ratios equal, images bit-identical (<= 1 ulp for tensor-domain mixup)."""
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree only exists in the build container")
    from PIL import Image
    from torchvision import transforms as T

    import device_input as D
    from oracle import input_ops as O

    ref_mix = ref_loader.load_reference_module("mix_dataset")
    ref_tf = ref_loader.load_reference_module("transforms")

    class Images:
        def __init__(self, u8, transform):
            self.u8, self.transform = u8, transform

        def __len__(self):
            return len(self.u8)

        def __getitem__(self, i):
            return self.transform(Image.fromarray(self.u8[i])), int(i)

    cfg = random.Random(77)
    modes = set()
    for trial in range(24):
        H, W, n = cfg.randint(4, 30), cfg.randint(4, 30), cfg.randint(2, 6)
        mixup, cutmix = cfg.choice([0.0, 0.2, 0.8]), cfg.choice([0.0, 0.5, 1.0])
        erasing, before, seed = cfg.choice([0.0, 0.3, 0.9]), cfg.random() < 0.5, cfg.randrange(10 ** 6)
        u8 = np.random.default_rng(trial).integers(0, 256, (n, H, W, 3), dtype=np.uint8)
        tail = [T.ToTensor(), T.Normalize(mean=list(O.MEAN), std=list(O.STD))]
        if erasing > 0:
            tail.append(ref_tf.RandomErasing(erasing, mode="const", max_count=1, num_splits=0, device="cpu"))
        if before:
            ds = ref_mix.MixDataset(Images(u8, lambda im: im), T.Compose(tail), mixup, cutmix)
        else:
            ds = ref_mix.MixDataset(Images(u8, T.Compose(tail)), T.Compose([]), mixup, cutmix)
        random.seed(seed)
        items = [ds[i] for i in range(n)]
        s = D.MixSampler(mixup, cutmix, erasing, before, rng=random.Random(seed))
        dec = [s.sample(i, n, H, W) for i in range(n)]
        what = (trial, H, W, n, mixup, cutmix, erasing, before)
        assert [d.index for d in dec] == [it[1] for it in items] and [d.partner for d in dec] == [it[2] for it in items], what
        assert [float(d.ratio) for d in dec] == [float(it[3]) for it in items], what
        got = O.input_batch(u8, D.pack_table(dec, {i: i for i in range(n)}, before, "const"))
        want = np.stack([it[0].numpy() for it in items])
        if not before and mixup > 0:
            assert ulp_diff(got, want) <= 1, what
        else:
            assert np.array_equal(got.view(np.int32), want.view(np.int32)), what
        modes |= {d.mode for d in dec}
    assert modes == {0, 1, 2}


def test_oracle_philox_known_answers():
    from oracle import input_ops as O

    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]  # Random123 kat_vectors, philox4x32 10
    for ctr, key, want in kat:
        assert tuple(int(v) for v in O.philox4x32_10(*ctr, *key)) == want


def test_oracle_blend_matches_pil_on_every_byte_pair(G):
    from oracle import input_ops as O

    a = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1)
    for alpha, sha in zip(G["blend"]["alphas"], G["blend"]["sha256"]):
        assert hashlib.sha256(O.pil_blend(a, a.T.copy(), alpha).tobytes()).hexdigest() == sha, alpha


def test_oracle_noise_is_standard_normal():
    from oracle import input_ops as O

    ys, xs = np.meshgrid(np.arange(224), np.arange(224), indexing="ij")
    z = O.erase_noise(1234, ys, xs).astype(np.float64)
    assert np.isfinite(z).all() and abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs((z ** 4).mean() - 3) < 0.1 and abs(np.corrcoef(z.reshape(3, -1))[0, 1]) < 0.01
    assert not np.array_equal(z, O.erase_noise(1235, ys, xs))


def test_check_table_rejects_out_of_range_rows():
    import device_input as D

    d = D.Decision(0)
    good = D.pack_table([d], {0: 0})
    D.check_table(good, 1, 8, 8)
    for col, val in ((0, 1), (1, -1), (2, 3), (3, 2), (12, 9), (11, -1), (20, 2)):
        bad = good.copy()
        bad[0, col] = val
        with pytest.raises(ValueError):
            D.check_table(bad, 1, 8, 8)
    with pytest.raises(ValueError):
        D.check_table(good.astype(np.int64), 1, 8, 8)
    with pytest.raises(ValueError):
        D.pack_table([d], {0: 0}, erase_mode="rand")


def test_device_input_fails_loudly_without_cuda():
    import device_input as D

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        D.DeviceInput()


# ------------------------------------------------------------------------------------------------ CPU: kernel source on the host
def run_emulated(u8, table, mean, std, variant=1):
    import kernel_emulation as K

    lib = K.build("input.cu")
    getattr(lib, "_Z21vtb_input_variant_seti")(variant)  # the C++ setter behind vtb_set_option("input_variant", v)
    S, H, W, _ = u8.shape
    B = table.shape[0]
    u8 = np.ascontiguousarray(u8)
    table = np.ascontiguousarray(table, np.int32)
    out = np.full((B, 3, H, W), np.nan, np.float32)
    f3 = C.c_float * 3
    rc = lib.vtb_input_batch(u8.ctypes.data_as(C.c_void_p), S, table.ctypes.data_as(C.c_void_p), B, H, W, f3(*mean),
                             f3(*std), out.ctypes.data_as(C.c_void_p), None)
    assert rc == 0
    return out


def mixed_table(n, H, W, seed, erase_mode):
    """Every mode x domain, erase boxes on both sources, boxes touching the borders."""
    import device_input as D

    rng = random.Random(seed)
    ds = []
    for i in range(n):
        d = D.Decision(i)
        d.partner = (i + 1 + rng.randrange(n - 1)) % n
        d.mode = i % 3
        d.weight = rng.random()
        x1, x2 = sorted(rng.randrange(W + 1) for _ in range(2))
        y1, y2 = sorted(rng.randrange(H + 1) for _ in range(2))
        d.box = (x1, y1, x2, y2) if d.mode == 2 else (0, 0, 0, 0)
        for attr in ("erase_a", "erase_b"):
            h, w = rng.randrange(H), rng.randrange(W)
            setattr(d, attr, (rng.randint(0, H - h), rng.randint(0, W - w), h, w) if rng.random() < 0.7 else (0, 0, 0, 0))
        d.seed_a, d.seed_b = rng.getrandbits(32), rng.getrandbits(32)
        ds.append(d)
    slots = {i: i for i in range(n)}
    t = np.concatenate([D.pack_table(ds, slots, True, erase_mode), D.pack_table(ds, slots, False, erase_mode)])
    D.check_table(t, n, H, W)
    return t


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_kernel_source_on_host_matches_golden_and_oracle(G, variant):
    from oracle import input_ops as O

    run_emulated = functools.partial(globals()["run_emulated"], variant=variant)
    for name, case in G["cases"].items():
        _, table = redraw(case)
        check_against_golden(case, run_emulated(case["u8"].numpy(), table, G["mean"], G["std"]), name)
    # 44 % 4 == 0 -> the 4-pixel path; 37 -> the scalar path; both with every mode, erase modes and a non-default mean / std
    # 44 / 1100 / 301: tiles spanning several rows, rows spanning several tiles, odd widths (scalar path)
    for (H, W), emode, mean, std in (((36, 44), "pixel", O.MEAN, O.STD), ((19, 37), "pixel", (0.5, 0.4, 0.3), (0.2, 0.25, 0.3)),
                                     ((36, 44), "const", O.MEAN, O.STD), ((5, 1100), "pixel", O.MEAN, O.STD),
                                     ((20, 56), "pixel", O.MEAN, O.STD), ((9, 1104), "const", O.MEAN, O.STD),
                                     ((3, 301), "const", O.MEAN, O.STD)):
        u8 = np.random.default_rng(H).integers(0, 256, (7, H, W, 3), dtype=np.uint8)
        table = mixed_table(7, H, W, seed=W, erase_mode=emode)
        got, want = run_emulated(u8, table, mean, std), O.input_batch(u8, table, mean, std)
        assert not np.isnan(got).any()
        # identical outside the noise boxes up to the fma of tensor mixup; libm log / sin / cos inside them
        assert np.abs(got - want).max() <= 2e-5
        exact = (got.view(np.int32) == want.view(np.int32)).mean()
        assert exact > 0.5, exact


def test_kernel_source_on_host_edge_cases():
    """Empty batch, a 1 x 1 image, a batch that reuses one source for every row, a partner equal to the image itself."""
    import device_input as D
    from oracle import input_ops as O

    u8 = np.random.default_rng(0).integers(0, 256, (3, 6, 8, 3), dtype=np.uint8)
    assert run_emulated(u8, np.zeros((0, D.TABLE_COLS), np.int32), O.MEAN, O.STD).shape == (0, 3, 6, 8)
    one = np.array([[[[0, 128, 255]]]], np.uint8)
    d = D.Decision(0)
    got = run_emulated(one, D.pack_table([d], {0: 0}), O.MEAN, O.STD, variant=2)
    assert np.array_equal(got, O.input_batch(one, D.pack_table([d], {0: 0})))
    ds = []
    for i in range(5):  # every row reads source 1; mixing an image with itself must return it (both domains)
        d = D.Decision(1)
        d.mode, d.weight, d.box = 1 + i % 2, 0.3, (1, 1, 7, 5)
        ds.append(d)
    for before in (True, False):
        table = D.pack_table(ds, {1: 1}, before, "const")
        for variant in (1, 2, 3):
            got = run_emulated(u8, table, O.MEAN, O.STD, variant=variant)
            plain = O._to_chw(u8[1], O.normalize_lut())
            tol = 0 if before else 5e-7  # r x + (1 - r) x in float32
            assert np.abs(got - plain[None]).max() <= tol, (before, variant)


def test_kernel_source_on_host_random_shapes_and_tables():
    """Seeded random image sizes (vector and scalar paths, tiles that straddle rows and images), batch sizes and decision
    rows: both kernel variants, built for the host, against the oracle."""
    from oracle import input_ops as O

    rng = random.Random(2024)
    for trial in range(12):
        H, W = rng.randint(1, 40), rng.choice([4, 8, 12, 16, 36, 72, 104]) if trial % 2 else rng.randint(1, 70)
        n = rng.randint(2, 5)
        u8 = np.random.default_rng(trial).integers(0, 256, (n, H, W, 3), dtype=np.uint8)
        table = mixed_table(n, H, W, seed=trial, erase_mode=rng.choice(["pixel", "const"]))
        table = table[np.random.default_rng(trial).permutation(len(table))][:rng.randint(1, len(table))]
        want = O.input_batch(u8, table)
        for variant in (1, 2, 3):
            got = run_emulated(u8, table, O.MEAN, O.STD, variant=variant)
            assert not np.isnan(got).any(), (trial, variant, H, W)
            assert np.abs(got - want).max() <= 2e-5, (trial, variant, H, W)


def test_device_input_wrapper_over_host_build(G, monkeypatch):
    """The Python side of the boundary (DeviceInput, make_batch: validation, table upload, ctypes call, empty batch) run
    end to end on the CPU against the host build of the kernel source, with the CUDA-only torch calls stubbed out."""
    import contextlib
    import types

    import device_input as D
    import kernel_emulation as K
    from oracle import input_ops as O
    from vtb200 import lib

    if torch.cuda.is_available():
        pytest.skip("CUDA present: the GPU tests cover the wrapper")
    emul = K.build("input.cu")
    emul.vtb_input_batch.argtypes = lib.load().vtb_input_batch.argtypes
    getattr(emul, "_Z21vtb_input_variant_seti")(1)
    monkeypatch.setattr(lib, "get", lambda: emul)
    monkeypatch.setattr(lib, "check", lambda rc, handle=None: None if rc == 0 else pytest.fail(f"rc={rc}"))
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, stream: None, raising=False)
    pipe = D.DeviceInput(G["mean"], G["std"], device="cpu")
    for name, case in G["cases"].items():
        _, table = redraw(case)
        check_against_golden(case, pipe(case["u8"], table).numpy(), name)
    assert pipe(torch.zeros((2, 8, 8, 3), dtype=torch.uint8), np.zeros((0, D.TABLE_COLS), np.int32)).shape == (0, 3, 8, 8)
    with pytest.raises(ValueError):
        pipe(torch.zeros((2, 8, 8, 4), dtype=torch.uint8), np.zeros((0, D.TABLE_COLS), np.int32))  # not RGB
    bad = D.pack_table([D.Decision(0)], {0: 5})
    with pytest.raises(ValueError):
        pipe(torch.zeros((2, 8, 8, 3), dtype=torch.uint8), bad)  # source slot 5 of 2
    n, H, W = 12, 32, 32
    data = np.random.default_rng(0).integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    labels = list(range(100, 100 + n))
    mk = lambda: D.MixSampler(0.8, 1.0, 0.25, mix_before_aug=True, rng=random.Random(9), noise_seed=4)  # noqa: E731
    batch, l1, l2, ratio = D.make_batch([0, 1, 2, 3, 4, 5], lambda i: data[i], n, mk(), D.DeviceInput(device="cpu"), labels)
    sampler = mk()
    ds = [sampler.sample(i, n, H, W) for i in range(6)]
    want = O.input_batch(data, D.pack_table(ds, {i: i for i in range(n)}, True, "pixel"))
    assert np.abs(batch.numpy() - want).max() <= 2e-5
    assert l1.tolist() == labels[:6] and l2.tolist() == [labels[d.partner] for d in ds]
    assert ratio.tolist() == [float(d.ratio) for d in ds]


# ------------------------------------------------------------------------------------------------ GPU
def run_device(u8, table, mean, std, variant=1):
    import device_input as D
    from vtb200 import lib

    pipe = D.DeviceInput(mean, std)
    lib.set_option("input_variant", variant)
    try:
        out = pipe(torch.from_numpy(np.ascontiguousarray(u8)), table)
        torch.cuda.synchronize()
    finally:
        lib.set_option("input_variant", 2)  # the library default
    return out.cpu().numpy()


@pytest.mark.gpu
def test_gpu_empty_batch_is_legal():
    import device_input as D

    out = D.DeviceInput()(torch.zeros((2, 8, 8, 3), dtype=torch.uint8), np.zeros((0, D.TABLE_COLS), np.int32))
    assert out.shape == (0, 3, 8, 8) and out.is_cuda


@pytest.mark.gpu
def test_gpu_matches_reference_batches(G):
    for name, case in G["cases"].items():
        _, table = redraw(case)
        check_against_golden(case, run_device(case["u8"].numpy(), table, G["mean"], G["std"]), name)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,emode", [(36, 44, "pixel"), (19, 37, "pixel"), (36, 44, "const"), (224, 224, "pixel")])
def test_gpu_matches_oracle_every_mode(H, W, emode):
    from oracle import input_ops as O

    u8 = np.random.default_rng(H).integers(0, 256, (7, H, W, 3), dtype=np.uint8)
    table = mixed_table(7, H, W, seed=W, erase_mode=emode)
    got, want = run_device(u8, table, O.MEAN, O.STD), O.input_batch(u8, table, O.MEAN, O.STD)
    assert np.abs(got - want).max() <= 2e-5  # noise: GPU libm vs numpy (log / sin / cos, |z| < 6); 1 ulp on tensor mixup
    if emode == "const":
        plain = np.isin(table[:, 2], (0, 2)) | (table[:, 3] == 0)  # everything but tensor mixup is bit-exact
        assert np.array_equal(got[plain].view(np.int32), want[plain].view(np.int32))


@pytest.mark.gpu
def test_gpu_noise_statistics_and_full_size_properties():
    """BASELINE size (256 x 3 x 224 x 224).  Properties: mode "none" equals the normalisation table applied to the bytes
    (checked on the device, bit-exact); a fully erased image is N(0, 1); cutmix with the full box returns the partner;
    uint8 mixup with alpha 0 / 1 returns img1 / the partner (PIL semantics); the launch is deterministic."""
    import device_input as D
    from oracle import input_ops as O

    B, H, W = 256, 224, 224
    g = torch.Generator().manual_seed(3)
    src = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g)
    ds = []
    for i in range(B):
        d = D.Decision(i)
        d.partner = (i + 1) % B
        kind = i % 4
        if kind == 1:
            d.mode, d.box = 2, (0, 0, W, H)
        elif kind == 2:
            d.mode, d.weight = 1, float(i % 8 == 2)  # blend alpha = 1 - weight: 0 or 1
        elif kind == 3:
            d.erase_a, d.seed_a = (0, 0, H, W), 1000 + i
        ds.append(d)
    table = D.pack_table(ds, {i: i for i in range(B)}, True, "pixel")
    pipe = D.DeviceInput()
    out = pipe(src, table)
    again = pipe(src, table)
    assert torch.equal(out, again)
    lut = torch.from_numpy(O.normalize_lut()).cuda()
    dev = src.cuda().long()
    plain = torch.stack([lut[c][dev[..., c]] for c in range(3)], 1)  # [B, 3, H, W]
    idx = torch.arange(B)
    partner = (idx + 1) % B
    assert torch.equal(out[idx % 4 == 0], plain[idx % 4 == 0])
    assert torch.equal(out[idx % 4 == 1], plain[partner[idx % 4 == 1]])
    assert torch.equal(out[idx % 8 == 2], plain[idx % 8 == 2])          # weight 1 -> alpha 0 -> img1
    assert torch.equal(out[idx % 8 == 6], plain[partner[idx % 8 == 6]])  # weight 0 -> alpha 1 -> partner
    z = out[idx % 4 == 3].double()
    assert torch.isfinite(z).all() and abs(z.mean().item()) < 3e-3 and abs(z.std().item() - 1) < 3e-3
    assert abs((z ** 4).mean().item() - 3) < 0.03
    assert not torch.equal(out[3], out[7])  # different seeds, different noise
    want = O.erase_noise(1003, *np.meshgrid(np.arange(H), np.arange(W), indexing="ij"))
    assert np.abs(out[3].cpu().numpy() - want).max() <= 2e-5


@pytest.mark.gpu
def test_gpu_make_batch_yields_the_loader_tuple():
    """make_batch = what `for input, label1, label2, ratio in loader` (train.py:265) receives, built on the device."""
    import device_input as D
    from oracle import input_ops as O

    n, H, W = 12, 32, 32
    data = np.random.default_rng(0).integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    labels = list(range(100, 100 + n))
    sampler = D.MixSampler(0.8, 1.0, 0.25, mix_before_aug=True, rng=random.Random(9), noise_seed=4)
    batch, l1, l2, ratio = D.make_batch([0, 1, 2, 3, 4, 5], lambda i: data[i], n, sampler, D.DeviceInput(), labels)
    assert batch.shape == (6, 3, H, W) and batch.dtype == torch.float32 and batch.is_cuda
    assert l1.tolist() == labels[:6] and all(100 <= v < 100 + n for v in l2.tolist())
    assert ratio.dtype == torch.float64 and ((ratio >= 0) & (ratio <= 1)).all()
    # the same decisions through the oracle
    sampler = D.MixSampler(0.8, 1.0, 0.25, mix_before_aug=True, rng=random.Random(9), noise_seed=4)
    ds = [sampler.sample(i, n, H, W) for i in range(6)]
    table = D.pack_table(ds, {i: i for i in range(n)}, True, "pixel")
    want = O.input_batch(data, table)
    assert np.abs(batch.cpu().numpy() - want).max() <= 2e-5


# ------------------------------------------------------------------------------------------------ DeferredMeter (train_util)
@pytest.mark.gpu
def test_gpu_deferred_meter_matches_blocking_meter():
    """The non-blocking meter ends with the same val / sum / count / avg as `Meter` fed by `.item()`,
    also when more values are in flight than it has slots."""
    import train_util as T

    vals = torch.arange(1, 41, dtype=torch.float32, device="cuda") * 0.25
    plain, deferred = T.Meter(), T.DeferredMeter(slots=8)
    for i, v in enumerate(vals):
        plain.update(v.item() * 2.0, i % 3 + 1)
        deferred.update_async(v, i % 3 + 1, scale=2.0)
    deferred.sync()
    assert deferred.count == plain.count and deferred.val == plain.val
    assert deferred.sum == pytest.approx(plain.sum, rel=1e-12) and deferred.avg == pytest.approx(plain.avg, rel=1e-12)
    assert not deferred._pending


# ------------------------------------------------------------------------------------------------ every kernel variant on the GPU
# (variant 2 is the library default since round 2: bit-identical to the reference's golden batches like the other two and
#  5-17 % faster on the B200, profiles/r02_input_variants.log)
@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("H,W,emode", [(36, 44, "pixel"), (19, 37, "const"), (5, 1100, "pixel"), (224, 224, "pixel")])
def test_gpu_lean_variants_match_oracle_and_golden(G, H, W, emode, variant):
    from oracle import input_ops as O

    u8 = np.random.default_rng(H).integers(0, 256, (7, H, W, 3), dtype=np.uint8)
    table = mixed_table(7, H, W, seed=W, erase_mode=emode)
    got, want = run_device(u8, table, O.MEAN, O.STD, variant=variant), O.input_batch(u8, table, O.MEAN, O.STD)
    assert np.abs(got - want).max() <= 2e-5
    for name, case in G["cases"].items():
        _, t = redraw(case)
        check_against_golden(case, run_device(case["u8"].numpy(), t, G["mean"], G["std"], variant=variant), name)
