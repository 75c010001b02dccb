"""Stand-alone GPU check of vtb_input_batch WITHOUT torch (starts in seconds: ctypes + numpy + libcudart): every mode x
domain x erase mode, vector and scalar variants, BASELINE image size, against the numpy oracle and the reference's golden
batches.  Test infrastructure (imports oracle/): run as `python tools/input_selftest.py` on the GPU box."""
import ctypes as C
import os
import sys
import time

import numpy as np

t_start = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import input_ops as O  # noqa: E402

EMULATE = os.environ.get("VTB_SELFTEST_EMULATE") == "1"  # dry run of THIS script on a box without a GPU (host build of the source)
rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
if EMULATE:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import kernel_emulation

    lib = kernel_emulation.build("input.cu")
else:
    lib = C.CDLL(os.path.join(ROOT, "vision-transformers-pytorch_b200", "vtb200", "libvtb200.so"))
if not EMULATE:
    lib.vtb_last_error.restype = C.c_char_p
VARIANT = 3 if "--variant3" in sys.argv else (2 if "--variant2" in sys.argv else 1)  # kernel variant (vtb_set_option("input_variant", v))
if EMULATE:
    getattr(lib, "_Z21vtb_input_variant_seti")(VARIANT)
else:
    lib.vtb_set_option.argtypes = [C.c_char_p, C.c_int32]
    assert lib.vtb_set_option(b"input_variant", VARIANT) == 0
print(f"input_selftest: kernel variant {VARIANT}{' (host emulation)' if EMULATE else ''}", flush=True)
lib.vtb_input_batch.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float),
                                C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
rt.cudaFree.argtypes = [C.c_void_p]
rt.cudaGetErrorString.restype = C.c_char_p


def ck(rc, what):
    if rc != 0:
        raise SystemExit(f"FAIL {what}: cuda error {rc} {rt.cudaGetErrorString(rc)}")


def dev_alloc(nbytes):
    p = C.c_void_p()
    ck(rt.cudaMalloc(C.byref(p), max(nbytes, 1)), "cudaMalloc")
    return p


def run(u8, table, mean=O.MEAN, std=O.STD):
    u8, table = np.ascontiguousarray(u8), np.ascontiguousarray(table, np.int32)
    S, H, W, _ = u8.shape
    B = table.shape[0]
    out = np.full((B, 3, H, W), np.nan, np.float32)
    f3 = C.c_float * 3
    if EMULATE:
        assert lib.vtb_input_batch(u8.ctypes.data, S, table.ctypes.data, B, H, W, f3(*mean), f3(*std), out.ctypes.data, None) == 0
        return out
    d_src, d_tab, d_out = dev_alloc(u8.nbytes), dev_alloc(table.nbytes), dev_alloc(out.nbytes)
    ck(rt.cudaMemcpy(d_src, u8.ctypes.data, u8.nbytes, 1), "H2D src")
    ck(rt.cudaMemcpy(d_tab, table.ctypes.data, table.nbytes, 1), "H2D table")
    ck(rt.cudaMemcpy(d_out, out.ctypes.data, out.nbytes, 1), "H2D out (NaN fill)")
    rc = lib.vtb_input_batch(d_src, S, d_tab, B, H, W, f3(*mean), f3(*std), d_out, None)
    if rc != 0:
        raise SystemExit(f"FAIL vtb_input_batch rc={rc}: {lib.vtb_last_error().decode()}")
    ck(rt.cudaDeviceSynchronize(), "sync")
    ck(rt.cudaMemcpy(out.ctypes.data, d_out, out.nbytes, 2), "D2H")
    for p in (d_src, d_tab, d_out):
        rt.cudaFree(p)
    return out


def table_all_modes(n, H, W, seed, emode):
    rng = np.random.default_rng(seed)
    rows = []
    for domain in (0, 1):
        for i in range(n):
            t = np.zeros(O.TABLE_COLS, np.int32)
            w = np.float32(rng.random())
            t[0], t[1], t[2], t[3] = i, (i + 1 + rng.integers(n - 1)) % n, i % 3, domain
            t[4] = np.array([1 - w if domain == 0 else w], np.float32).view(np.int32)[0]
            t[5] = np.array([1 - w], np.float32).view(np.int32)[0]
            if t[2] == 2:
                xs, ys = np.sort(rng.integers(0, W + 1, 2)), np.sort(rng.integers(0, H + 1, 2))
                t[6:10] = xs[0], ys[0], xs[1], ys[1]
            for c0 in (10, 14):
                if rng.random() < 0.7:
                    h, w_ = int(rng.integers(0, H)), int(rng.integers(0, W))
                    t[c0:c0 + 4] = rng.integers(0, H - h + 1), rng.integers(0, W - w_ + 1), h, w_
            t[18:20] = rng.integers(-2 ** 31, 2 ** 31, 2)
            t[20] = emode
            rows.append(t)
    return np.stack(rows)


NPZ = os.path.join(ROOT, "tools", "_selftest_input.npz")


def prepare():
    """Build container: golden cases + their tables (product sampler) + the BASELINE-size bench tables -> NPZ (no torch needed
    to read it on the GPU box)."""
    import random

    import torch

    sys.path.insert(0, os.path.join(ROOT, "vision-transformers-pytorch_b200"))
    import device_input as D

    G = torch.load(os.path.join(ROOT, "tests", "golden", "input_ops.pt"), map_location="cpu", weights_only=False)
    out = {"mean": np.array(G["mean"], np.float64), "std": np.array(G["std"], np.float64)}
    for name, case in G["cases"].items():
        s = D.MixSampler(case["mixup"], case["cutmix"], case["erasing"], case["mix_before_aug"], rng=random.Random(case["seed"]))
        ds = [s.sample(i, case["n"], case["H"], case["W"]) for i in range(case["n"])]
        out[f"g_{name}_table"] = D.pack_table(ds, {i: i for i in range(case["n"])}, case["mix_before_aug"], "const")
        out[f"g_{name}_u8"], out[f"g_{name}_img"] = case["u8"].numpy(), case["img"].numpy()
    B = 256
    for tag, kw in (("none", dict(mixup=0.0, cutmix=0.0, erasing=0.0)), ("swin_conf", dict(mixup=0.8, cutmix=1.0, erasing=0.25)),
                    ("tensor_order", dict(mixup=0.8, cutmix=1.0, erasing=0.25, mix_before_aug=False))):
        s = D.MixSampler(rng=random.Random(0), **kw)
        ds = [s.sample(i, B, 224, 224) for i in range(B)]
        out[f"b_{tag}"] = D.pack_table(ds, {i: i for i in range(B)}, s.mix_before_aug, "pixel")
    np.savez_compressed(NPZ, **out)
    print("wrote", NPZ, os.path.getsize(NPZ), "bytes")


def golden_npz():
    global fails
    Z = np.load(NPZ)
    for key in [k for k in Z.files if k.endswith("_table") and k.startswith("g_")]:
        name = key[2:-6]
        got, want = run(Z[f"g_{name}_u8"], Z[key], Z["mean"], Z["std"]), Z[f"g_{name}_img"]
        ok = np.abs(got - want).max() <= 5e-7 if "tensor_mix" in name else np.array_equal(got.view(np.int32), want.view(np.int32))
        fails += not ok
        print(f"{'PASS' if ok else 'FAIL'} golden {name}: max |gpu - reference| = {np.abs(got - want).max():.3e}", flush=True)


def bench():
    """BASELINE batch (256 x 224 x 224): device-resident uint8 sources, L2 flushed between launches, CUDA events."""
    Z = np.load(NPZ)
    B, H, W = 256, 224, 224
    peak = 6543.1
    try:
        import json

        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    u8 = np.random.default_rng(1).integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    d_src, d_out, d_flush = dev_alloc(u8.nbytes), dev_alloc(B * 3 * H * W * 4), dev_alloc(256 << 20)
    ck(rt.cudaMemcpy(d_src, u8.ctypes.data, u8.nbytes, 1), "H2D src")
    rt.cudaEventCreate.argtypes = [C.POINTER(C.c_void_p)]
    rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
    rt.cudaEventSynchronize.argtypes = [C.c_void_p]
    rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
    e0, e1 = C.c_void_p(), C.c_void_p()
    ck(rt.cudaEventCreate(C.byref(e0)), "event")
    ck(rt.cudaEventCreate(C.byref(e1)), "event")
    f3 = C.c_float * 3
    for tag in ("swin_conf", "none", "tensor_order"):
        table = np.ascontiguousarray(Z[f"b_{tag}"], np.int32)
        d_tab = dev_alloc(table.nbytes)
        ck(rt.cudaMemcpy(d_tab, table.ctypes.data, table.nbytes, 1), "H2D table")
        mixed = int((table[:, 2] != 0).sum())
        nbytes = B * H * W * 12 + (B + mixed) * H * W * 3  # output written once, every source byte read once per use
        ts = []
        for it in range(9):
            ck(rt.cudaMemset(d_flush, it, 256 << 20), "flush")
            ck(rt.cudaEventRecord(e0, None), "record")
            rc = lib.vtb_input_batch(d_src, B, d_tab, B, H, W, f3(*O.MEAN), f3(*O.STD), d_out, None)
            ck(rt.cudaEventRecord(e1, None), "record")
            ck(rt.cudaEventSynchronize(e1), "sync")
            if rc != 0:
                raise SystemExit(f"FAIL vtb_input_batch rc={rc}: {lib.vtb_last_error().decode()}")
            ms = C.c_float()
            ck(rt.cudaEventElapsedTime(C.byref(ms), e0, e1), "elapsed")
            if it >= 3:
                ts.append(ms.value)
        ms = sum(ts) / len(ts)
        print(f"BENCH {tag:13s} B=256 224x224 ({mixed} mixed): {ms*1e3:7.1f} us  {B/ms*1e3:9.0f} img/s  {nbytes/1e6:6.1f} MB  "
              f"{nbytes/ms/1e6:6.0f} GB/s = {nbytes/ms/1e6/peak*100:4.1f} % of {peak:.0f} GB/s (min {min(ts)*1e3:.1f} us)", flush=True)
        rt.cudaFree(d_tab)


if "--prepare" in sys.argv:
    prepare()
    sys.exit(0)
if "--bench-only" in sys.argv:
    ck(rt.cudaSetDevice(0), "cudaSetDevice")
    bench()
    sys.exit(0)
if not EMULATE:
    ck(rt.cudaSetDevice(0), "cudaSetDevice")
fails = 0
for (H, W), emode in (((36, 44), 1), ((19, 37), 1), ((36, 44), 0), ((224, 224), 1)):
    u8 = np.random.default_rng(H).integers(0, 256, (7, H, W, 3), dtype=np.uint8)
    table = table_all_modes(7, H, W, W, emode)
    got, want = run(u8, table), O.input_batch(u8, table)
    err = float(np.nanmax(np.abs(got - want))) if not np.isnan(got).any() else float("nan")
    exact = float((got.view(np.int32) == want.view(np.int32)).mean())
    ok = err <= 2e-5
    if emode == 0:  # no noise: everything but tensor-domain mixup must be bit-identical
        plain = np.isin(table[:, 2], (0, 2)) | (table[:, 3] == 0)
        ok = ok and np.array_equal(got[plain].view(np.int32), want[plain].view(np.int32))
    fails += not ok
    print(f"{'PASS' if ok else 'FAIL'} H={H} W={W} erase_mode={emode}: max |gpu - oracle| = {err:.3e}, bit-identical {exact*100:.2f} %",
          flush=True)

if os.path.exists(NPZ):
    golden_npz()
    if not EMULATE:
        bench()
    print(f"input_selftest: {'ALL PASS' if fails == 0 else str(fails) + ' FAILED'} in {time.time() - t_start:.1f} s", flush=True)
    sys.exit(1 if fails else 0)

try:  # the reference's own batches (golden), decisions re-drawn by the product's sampler; needs torch only to unpickle
    if os.environ.get("VTB_SELFTEST_NO_GOLDEN") == "1":
        raise RuntimeError("VTB_SELFTEST_NO_GOLDEN=1")
    import random

    import torch  # noqa: F401

    sys.path.insert(0, os.path.join(ROOT, "vision-transformers-pytorch_b200"))
    import device_input as D

    G = torch.load(os.path.join(ROOT, "tests", "golden", "input_ops.pt"), map_location="cpu", weights_only=False)
    for name, case in G["cases"].items():
        s = D.MixSampler(case["mixup"], case["cutmix"], case["erasing"], case["mix_before_aug"], rng=random.Random(case["seed"]))
        ds = [s.sample(i, case["n"], case["H"], case["W"]) for i in range(case["n"])]
        table = D.pack_table(ds, {i: i for i in range(case["n"])}, case["mix_before_aug"], "const")
        got, want = run(case["u8"].numpy(), table, G["mean"], G["std"]), case["img"].numpy()
        if "tensor" in name and case["mixup"] > 0:
            ok = np.abs(got - want).max() <= 5e-7
        else:
            ok = np.array_equal(got.view(np.int32), want.view(np.int32))
        fails += not ok
        print(f"{'PASS' if ok else 'FAIL'} golden {name}: max |gpu - reference| = {np.abs(got - want).max():.3e}", flush=True)
except Exception as exc:  # noqa: BLE001
    print("golden leg skipped:", repr(exc))
print(f"input_selftest: {'ALL PASS' if fails == 0 else str(fails) + ' FAILED'} in {time.time() - t_start:.1f} s", flush=True)
sys.exit(1 if fails else 0)
