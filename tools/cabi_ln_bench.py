"""Torch-free check + micro-benchmark of vtb_layernorm_fwd / vtb_layernorm_bwd through the C-ABI (ctypes + numpy +
libcudart; see tools/cudart_ctypes.py): a float64 numpy LayerNorm + gradients on a small ragged problem first, then
CUDA-event timings against the HBM roofline on the ViT-B and Swin-S stage shapes (dense rows; fwd writes bf16,
bwd reads bf16 dy + the incoming residual gradient and also emits the bf16 copy of dx, as a transformer block uses it)."""
import json
import os
import sys
import time

import numpy as np

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vision-transformers-pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cudart_ctypes as cu  # noqa: E402
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


class LN:
    def __init__(self, rows, cols, seed=None, host=None):
        self.rows, self.cols = rows, cols
        mk = (lambda shape, dt, key: cu.Buf(shape, dt).upload(host[key])) if host else \
             (lambda shape, dt, key: cu.Buf(shape, dt).fill_from(seed[dt]))
        self.x, self.dx_in = mk((rows, cols), F32, "x"), mk((rows, cols), F32, "dx_in")
        self.gamma, self.beta = mk((cols,), F32, "gamma"), mk((cols,), F32, "beta")
        self.dy = mk((rows, cols), BF16, "dy")
        self.y, self.dxb = cu.Buf((rows, cols), BF16), cu.Buf((rows, cols), BF16)
        self.mean, self.rstd = cu.Buf(rows, F32), cu.Buf(rows, F32)
        self.dx = cu.Buf((rows, cols), F32)
        self.dgamma, self.dbeta = cu.Buf(cols, F32).zero(), cu.Buf(cols, F32).zero()

    def fwd(self, eps=1e-6):
        L.check(lib.vtb_layernorm_fwd(self.x.ptr, self.gamma.ptr, self.beta.ptr, eps, self.rows, self.cols, 0, 0, 0,
                                      self.y.ptr, 0, self.mean.ptr, self.rstd.ptr, None, 0, None), lib)

    def bwd(self):
        L.check(lib.vtb_layernorm_bwd(self.dy.ptr, 0, self.x.ptr, self.gamma.ptr, self.mean.ptr, self.rstd.ptr, self.rows,
                                      self.cols, 0, 0, 0, self.dx_in.ptr, self.dx.ptr, self.dxb.ptr, None, 0,
                                      self.dgamma.ptr, self.dbeta.ptr, None, None), lib)

    def free(self):
        for b in (self.x, self.dx_in, self.gamma, self.beta, self.dy, self.y, self.dxb, self.mean, self.rstd, self.dx,
                  self.dgamma, self.dbeta):
            b.free()


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30))


def self_check(rows, cols, eps=1e-6):
    rng = np.random.default_rng(rows + cols)
    host = {"x": (rng.standard_normal((rows, cols)) * 2 + 0.5).astype(F32), "dx_in": rng.standard_normal((rows, cols)).astype(F32),
            "gamma": (1 + 0.1 * rng.standard_normal(cols)).astype(F32), "beta": (0.1 * rng.standard_normal(cols)).astype(F32),
            "dy": cu.to_bf16_bits(rng.standard_normal((rows, cols)).astype(F32))}
    ln = LN(rows, cols, host=host)
    ln.fwd(eps)
    ln.bwd()
    cu.ck(cu.rt.cudaDeviceSynchronize(), "sync")
    x, g, b = host["x"].astype(np.float64), host["gamma"].astype(np.float64), host["beta"].astype(np.float64)
    dy = cu.from_bf16_bits(host["dy"]).astype(np.float64)
    mu, var = x.mean(1, keepdims=True), x.var(1, keepdims=True)
    rstd = 1 / np.sqrt(var + eps)
    xh = (x - mu) * rstd
    gd = dy * g
    dx = host["dx_in"] + rstd * (gd - gd.mean(1, keepdims=True) - xh * (gd * xh).mean(1, keepdims=True))
    errs = {"y": rel(cu.from_bf16_bits(ln.y.download()), xh * g + b), "dx": rel(ln.dx.download(), dx),
            "dx_bf16": rel(cu.from_bf16_bits(ln.dxb.download()), dx), "dgamma": rel(ln.dgamma.download(), (dy * xh).sum(0)),
            "dbeta": rel(ln.dbeta.download(), dy.sum(0)), "mean": rel(ln.mean.download(), mu[:, 0]),
            "rstd": rel(ln.rstd.download(), rstd[:, 0])}
    bars = {"y": 4e-3, "dx": 1e-5, "dx_bf16": 4e-3, "dgamma": 1e-4, "dbeta": 1e-4, "mean": 1e-5, "rstd": 1e-5}
    ok = all(errs[k] < bars[k] for k in errs)
    print(f"{'PASS' if ok else 'FAIL'} self-check [{rows}, {cols}]: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()), flush=True)
    ln.free()
    return ok


ok = self_check(333, 768) & self_check(1000, 96) & self_check(77, 384) & self_check(129, 192)
if not ok:
    raise SystemExit("FAIL: LayerNorm self-check")
rng = np.random.default_rng(1)
n_seed = 16 << 20
seed = {BF16: cu.Buf(n_seed, BF16).upload(cu.to_bf16_bits(rng.standard_normal(n_seed, F32))),
        F32: cu.Buf(n_seed, F32).upload(rng.standard_normal(n_seed, F32))}
timer = cu.Timer()
flush = cu.Buf(256 << 20, np.uint8)
for tag, rows, cols in (("ViT-B        ", 50432, 768), ("Swin-S stage1", 802816, 96), ("Swin-S stage2", 200704, 192),
                        ("Swin-S stage3", 50176, 384), ("Swin-S stage4", 12544, 768)):
    ln = LN(rows, cols, seed=seed)
    ln.fwd()
    line = f"{tag} [{rows:6d}, {cols:3d}]"
    for name, fn, bpe in (("fwd", ln.fwd, 6), ("bwd", ln.bwd, 16)):  # fwd: x f32 in, y bf16 out; bwd: dy bf16, x, dx_in, dx f32, dx bf16
        ts = []
        for it in range(7):
            flush.zero()
            us = timer.time(fn, n=1, warmup=0)
            if it >= 2:
                ts.append(us)
        us = sum(ts) / len(ts)
        nb = bpe * rows * cols
        line += f"  {name} {us:7.1f} us {nb / us / 1e3:6.0f} GB/s = {nb / us / 1e3 / PEAK_GB * 100:4.1f} %"
    print(line, flush=True)
    ln.free()
print(f"cabi_ln_bench: done in {time.time() - t0:.1f} s", flush=True)
