"""Torch-free check + micro-benchmark of the multi-tensor step-side kernels (vtb_mt_ema / grad_norm / scale / adamw /
cast_f32_bf16) through the C-ABI (ctypes + numpy + libcudart; see tools/cudart_ctypes.py) on the ViT-B/16 parameter list
(152 tensors, 86.6 M elements): a ragged small list is first checked against the numpy oracle (oracle/step_ops.py —
this file is test infrastructure), then each kernel is timed with CUDA events against the HBM roofline."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vision-transformers-pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cudart_ctypes as cu  # noqa: E402
from oracle import step_ops as S  # noqa: E402
from vtb200 import lib as L  # noqa: E402

cu.init()
lib = L.load()
L.check(lib.vtb_init(), lib)
F32, BF16 = np.float32, np.uint16
PEAK_GB = 6543.1
try:
    PEAK_GB = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


class DevList:
    """A tensor list on the device: host arrays of device pointers + element counts (the C-ABI's list format)."""

    def __init__(self, arrays=None, shapes=None, dtype=F32, seed=None):
        if arrays is not None:
            self.bufs = [cu.Buf(a.shape, dtype).upload(a) for a in arrays]
        else:
            self.bufs = [cu.Buf(s, dtype) for s in shapes]
            for b in self.bufs:
                b.fill_from(seed) if seed is not None else b.zero()
        self.n = len(self.bufs)
        self.ptrs = (C.c_void_p * self.n)(*[b.addr for b in self.bufs])
        self.numel = (C.c_int64 * self.n)(*[int(np.prod(b.shape)) for b in self.bufs])
        self.total = sum(int(np.prod(b.shape)) for b in self.bufs)

    def download(self):
        return [b.download() for b in self.bufs]


def grad_norm(grads, max_norm):
    chunks = int(lib.vtb_mt_num_chunks(grads.numel, grads.n))
    partials, out = cu.Buf(max(chunks, 1), F32), cu.Buf(2, F32)
    L.check(lib.vtb_mt_grad_norm(grads.ptrs, grads.numel, grads.n, float(max_norm), partials.ptr, out.ptr, None), lib)
    return partials, out


def self_check():
    rng = np.random.default_rng(0)
    shapes = [(7, 5), (13,), (3, 4, 2, 2), (1,), (40, 33), (9000,), (257, 129)]
    mk = lambda scale=1.0: [(rng.standard_normal(s) * scale).astype(F32) for s in shapes]  # noqa: E731
    ok = True
    # EMA
    dst, src = mk(), mk()
    d, s = DevList(dst), DevList(src)
    L.check(lib.vtb_mt_ema(d.ptrs, s.ptrs, s.numel, s.n, 0.996, None), lib)
    err = max(np.abs(a - b).max() for a, b in zip(d.download(), S.ema(dst, src, 0.996)))
    ok &= err <= 2.4e-7
    print(f"{'PASS' if err <= 2.4e-7 else 'FAIL'} ema: max abs err {err:.2e}", flush=True)
    # gradient norm + clip coefficient, then the rescale
    g = mk(3.0)
    dg = DevList(g)
    _, out = grad_norm(dg, 5.0)
    L.check(lib.vtb_mt_scale(dg.ptrs, dg.numel, dg.n, C.c_void_p(out.addr + 4), None), lib)
    want, total = S.clip_grad_norm(g, 5.0)
    got = out.download()
    e1 = abs(got[0] - total) / total
    e2 = max(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) for a, b in zip(dg.download(), want))
    ok &= e1 < 1e-6 and e2 < 1e-6
    print(f"{'PASS' if e1 < 1e-6 and e2 < 1e-6 else 'FAIL'} grad_norm + scale: norm rel err {e1:.2e}, clipped grads {e2:.2e}",
          flush=True)
    # AdamW, two steps
    p, gr = mk(), mk()
    m, v = [np.zeros_like(x) for x in p], [np.zeros_like(x) for x in p]
    dp, dgr, dm, dv = DevList(p), DevList(gr), DevList(m), DevList(v)
    hp = dict(lr=2.5e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.05)
    for step in (1, 2):
        L.check(lib.vtb_mt_adamw(dp.ptrs, dgr.ptrs, dm.ptrs, dv.ptrs, None, dp.numel, dp.n, hp["lr"], hp["beta1"], hp["beta2"],
                                 hp["eps"], hp["weight_decay"], step, None, None), lib)
        for i in range(len(p)):
            p[i], m[i], v[i] = S.adamw_step(p[i], gr[i], m[i], v[i], step=step, **hp)
    e3 = max(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30) for a, b in zip(dp.download(), p))
    ok &= e3 < 1e-6
    print(f"{'PASS' if e3 < 1e-6 else 'FAIL'} adamw x2: params rel err {e3:.2e}", flush=True)
    # bf16 cast
    x = mk()
    dx, db = DevList(x), DevList(shapes=shapes, dtype=BF16)
    L.check(lib.vtb_mt_cast_f32_bf16(dx.ptrs, db.ptrs, dx.numel, dx.n, None), lib)
    same = all(np.array_equal(a, cu.to_bf16_bits(b)) for a, b in zip(db.download(), x))
    ok &= same
    print(f"{'PASS' if same else 'FAIL'} cast_f32_bf16: bit-identical to round-to-nearest-even", flush=True)
    cu.ck(cu.rt.cudaDeviceSynchronize(), "sync")
    return ok


def vit_b16_shapes():
    D, FF, depth = 768, 3072, 12
    shapes = [(1, 1, D), (1, 197, D), (D, 3, 16, 16), (D,)]
    for _ in range(depth):
        shapes += [(D,), (D,), (3 * D, D), (3 * D,), (D, D), (D,), (D,), (D,), (FF, D), (FF,), (D, FF), (D,)]
    return shapes + [(D,), (D,), (1000, D), (1000,)]


if not self_check():
    raise SystemExit("FAIL: multi-tensor self-check")
rng = np.random.default_rng(1)
n_seed = 8 << 20
seed = cu.Buf(n_seed, F32).upload((0.02 * rng.standard_normal(n_seed)).astype(F32))
shapes = vit_b16_shapes()
P, G, M, V = (DevList(shapes=shapes, seed=seed) for _ in range(4))
V2 = DevList(shapes=shapes)  # exp_avg_sq starts at zero (non-negative)
PB = DevList(shapes=shapes, dtype=BF16)
flush = cu.Buf(256 << 20, np.uint8)
timer = cu.Timer()
print(f"ViT-B/16 list: {P.n} tensors, {P.total / 1e6:.1f} M elements", flush=True)
partials, out = grad_norm(G, 5.0)


def timed(name, bytes_per_elem, fn):
    ts = []
    for it in range(8):
        flush.zero()  # > L2
        us = timer.time(fn, n=1, warmup=0)
        if it >= 3:
            ts.append(us)
    us = sum(ts) / len(ts)
    nb = bytes_per_elem * P.total
    print(f"{name:34s} {us:8.1f} us  {nb / us / 1e3:7.0f} GB/s = {nb / us / 1e3 / PEAK_GB * 100:4.1f} % of {PEAK_GB:.0f}", flush=True)


timed("mt_ema            (12 B/elem)", 12, lambda: L.check(lib.vtb_mt_ema(M.ptrs, P.ptrs, P.numel, P.n, 0.996, None), lib))
timed("mt_grad_norm       (4 B/elem)", 4, lambda: L.check(lib.vtb_mt_grad_norm(G.ptrs, G.numel, G.n, 5.0, partials.ptr, out.ptr, None), lib))
timed("mt_scale           (8 B/elem)", 8, lambda: L.check(lib.vtb_mt_scale(G.ptrs, G.numel, G.n, C.c_void_p(out.addr + 4), None), lib))
timed("mt_adamw          (28 B/elem)", 28, lambda: L.check(lib.vtb_mt_adamw(P.ptrs, G.ptrs, M.ptrs, V2.ptrs, None, P.numel, P.n, 2.5e-4, 0.9, 0.999, 1e-8, 0.05, 1, C.c_void_p(out.addr + 4), None), lib))
timed("mt_adamw + bf16   (30 B/elem)", 30, lambda: L.check(lib.vtb_mt_adamw(P.ptrs, G.ptrs, M.ptrs, V2.ptrs, PB.ptrs, P.numel, P.n, 2.5e-4, 0.9, 0.999, 1e-8, 0.05, 2, C.c_void_p(out.addr + 4), None), lib))
timed("mt_cast_f32_bf16   (6 B/elem)", 6, lambda: L.check(lib.vtb_mt_cast_f32_bf16(P.ptrs, PB.ptrs, P.numel, P.n, None), lib))
print(f"cabi_step_bench: done in {time.time() - t0:.1f} s", flush=True)
