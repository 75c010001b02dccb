"""Torch-free micro-benchmark of vtb_gemm_bf16 through the C-ABI (ctypes + numpy + libcudart): the block GEMMs of ViT-B /
Swin-S stage 3 / stage 1 with the epilogues the models use — the same cases as tools/bench_gemm.py, but the process starts
in a second instead of a torch import, so one gpurun call costs ~25 s of budget.  A small product is checked against numpy
first (catches a wrong parameter block before any number is printed).
  GEMM_BLOCK=swin3,vitb,swin1   GEMM_ONLY=<substring of a case name>"""
import ctypes as C
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
from vtb200 import lib as L  # noqa: E402  (ctypes only; never imports torch unless get() is called)

cu.init()
lib = L.load()
L.check(lib.vtb_init(), lib)
PEAK_TF, PEAK_GB = 1392.6, 6543.1
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_TF, PEAK_GB = pk["bf16_tflops_sustained"], pk["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass
F32, BF16 = np.float32, np.uint16


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, out=None, out2=None, bias=None, resid=None, row_scale=None,
         rows_per_scale=0, epilogue=0, aux=None, accumulate=False, a_colsum=None):
    """Mirrors vtb200.ops.gemm for contiguous buffers: A [M,K] (or [K,M] if a_mn), B [N,K] (or [K,N] if b_mn)."""
    p = L.GemmParams()
    p.M, p.N, p.K = M, N, K
    p.A, p.lda, p.a_mn_major = a.addr, (M if a_mn else K), int(a_mn)
    p.B, p.ldb, p.b_mn_major = b.addr, (N if b_mn else K), int(b_mn)
    p.out, p.ldo, p.out_f32 = out.addr, N, int(out.dtype == np.float32)
    if out2 is not None:
        p.out2 = out2.addr
    if bias is not None:
        p.bias = bias.addr
    if resid is not None:
        p.resid, p.ldr = resid.addr, N
    if row_scale is not None:
        p.row_scale, p.rows_per_scale = row_scale.addr, rows_per_scale
    if aux is not None:
        p.aux, p.ldaux = aux.addr, N
    p.epilogue, p.splits, p.accumulate, p.alpha = epilogue, (int(os.environ.get("GEMM_SPLITS", 0)) if accumulate else 0), int(accumulate), 1.0
    if a_colsum is not None:
        p.a_colsum = a_colsum.addr
    L.check(lib.vtb_gemm_bf16(C.byref(p), None), lib)


def self_check():
    rng = np.random.default_rng(0)
    M, N, K = 256, 384, 128
    a, b = rng.standard_normal((M, K), F32), rng.standard_normal((N, K), F32)
    bias = rng.standard_normal(N).astype(F32)
    ab, bb = cu.to_bf16_bits(a), cu.to_bf16_bits(b)
    da, db, dbias = cu.Buf((M, K), BF16).upload(ab), cu.Buf((N, K), BF16).upload(bb), cu.Buf(N, F32).upload(bias)
    dout = cu.Buf((M, N), F32)
    gemm(da, db, M, N, K, out=dout, bias=dbias)
    cu.ck(cu.rt.cudaDeviceSynchronize(), "sync")
    want = cu.from_bf16_bits(ab).astype(np.float64) @ cu.from_bf16_bits(bb).astype(np.float64).T + bias
    err = np.abs(dout.download() - want).max() / np.abs(want).max()
    print(f"self-check y = x W^T + b ({M}x{N}x{K}, f32 out): max rel err {err:.2e}", flush=True)
    if not err < 1e-5:
        raise SystemExit("FAIL: the harness's parameter block does not reproduce a plain product")
    # dgrad / wgrad operand-major flags: dx = dy W (B MN-major), dW = dy^T x (A, B MN-major, accumulate into zeros)
    dy = rng.standard_normal((M, N), F32)
    dyb = cu.to_bf16_bits(dy)
    ddy, ddx, ddw = cu.Buf((M, N), BF16).upload(dyb), cu.Buf((M, K), F32), cu.Buf((N, K), F32).zero()
    gemm(ddy, db, M, K, N, b_mn=True, out=ddx)
    gemm(ddy, da, N, K, M, a_mn=True, b_mn=True, out=ddw, accumulate=True)
    cu.ck(cu.rt.cudaDeviceSynchronize(), "sync")
    dyf = cu.from_bf16_bits(dyb).astype(np.float64)
    e1 = np.abs(ddx.download() - dyf @ cu.from_bf16_bits(bb)).max() / np.abs(dyf @ cu.from_bf16_bits(bb)).max()
    e2 = np.abs(ddw.download() - dyf.T @ cu.from_bf16_bits(ab)).max() / np.abs(dyf.T @ cu.from_bf16_bits(ab)).max()
    print(f"self-check dgrad {e1:.2e}, wgrad {e2:.2e}", flush=True)
    if not (e1 < 1e-5 and e2 < 1e-5):
        raise SystemExit("FAIL: dgrad / wgrad operand flags")


def trace_report(name, fn):
    """Cycle counters of CTA 7 (library built with -DVTB_GEMM_TRACE, VTB_LIB=libvtb200_trace.so): per tile and per sub-tile step."""
    lib.vtb_debug_gemm_trace.argtypes = [C.c_void_p, C.c_int]
    buf = (C.c_ulonglong * 32)()
    lib.vtb_debug_gemm_trace(None, 1)
    n = 3
    for _ in range(n):
        fn()
    cu.ck(cu.rt.cudaDeviceSynchronize(), "sync")
    lib.vtb_debug_gemm_trace(buf, 1)
    v = list(buf)
    te, ti = max(v[18], 1), max(v[22], 1)
    names = ["t0", "t1", "tmem_ld_wait", "issue next ld + aux wait + free-buffer wait", "math+STS", "fence.proxy", "syncwarp+arrive", "bookkeeping"]
    print(f"   trace {name.strip()}: {te / n:.1f} tiles per launch on CTA 7; per tile: epilogue {v[17] / te:.0f} cycles of which "
          f"{v[16] / te:.0f} waiting for the accumulator; issuer {v[21] / ti:.0f} of which {v[20] / ti:.0f} waiting for a TMEM stage, "
          f"{v[19] / ti:.0f} for operands")
    for half in (0, 1):
        tot = sum(v[8 * half:8 * half + 8])
        print(f"     warp-half {half} (4 quarters summed / 4, per tile): " + ", ".join(
            f"{names[i]}={v[8 * half + i] / 4 / te:.0f}" for i in range(2, 8)) + f" | total {tot / 4 / te:.0f}")


def block(T, Cc, FF, tag, seed):
    QKV = 3 * Cc
    timer = cu.Timer()

    def rb(shape, dtype=BF16):  # random-filled device buffer
        return cu.Buf(shape, dtype).fill_from(seed[dtype])

    y, o, g = rb((T, Cc)), rb((T, Cc)), rb((T, Cc))
    x = rb((T, Cc), F32)
    wq, wo, w1, w2 = rb((QKV, Cc)), rb((Cc, Cc)), rb((FF, Cc)), rb((Cc, FF))
    bq, bo, b1, b2 = rb(QKV, F32), rb(Cc, F32), rb(FF, F32), rb(Cc, F32)
    qkv, u, h, du = rb((T, QKV)), rb((T, FF)), rb((T, FF)), cu.Buf((T, FF), BF16)
    out, dyb = cu.Buf((T, Cc), F32), cu.Buf((T, Cc), BF16)
    dw1, dw2, dwq = cu.Buf((FF, Cc), F32).zero(), cu.Buf((Cc, FF), F32).zero(), cu.Buf((QKV, Cc), F32).zero()
    db1, dbq = cu.Buf(FF, F32).zero(), cu.Buf(QKV, F32).zero()
    scale = cu.Buf(T // 196 + 1, F32).upload(np.ones(T // 196 + 1, F32))
    SILU_DUAL, SILU_GRAD = L.EPI_SILU_DUAL, L.EPI_SILU_GRAD
    cases = [
        ("qkv  fwd bf16+bias", 2 * T * QKV * Cc, (T * Cc + T * QKV) * 2, lambda: gemm(y, wq, T, QKV, Cc, out=qkv, bias=bq)),
        ("proj fwd f32+resid", 2 * T * Cc * Cc, T * Cc * 2 + T * Cc * 8,
         lambda: gemm(o, wo, T, Cc, Cc, out=out, bias=bo, resid=x, row_scale=scale, rows_per_scale=196)),
        ("fc1  fwd silu-dual", 2 * T * FF * Cc, T * Cc * 2 + T * FF * 4,
         lambda: gemm(y, w1, T, FF, Cc, out=u, out2=h, bias=b1, epilogue=SILU_DUAL)),
        ("fc2  fwd f32+resid", 2 * T * FF * Cc, T * FF * 2 + T * Cc * 8,
         lambda: gemm(h, w2, T, Cc, FF, out=out, bias=b2, resid=x, row_scale=scale, rows_per_scale=196)),
        ("fc2 dgrad silugrad", 2 * T * FF * Cc, T * Cc * 2 + T * FF * 4,
         lambda: gemm(g, w2, T, FF, Cc, b_mn=True, out=du, epilogue=SILU_GRAD, aux=u)),
        ("fc1 dgrad bf16    ", 2 * T * FF * Cc, T * FF * 2 + T * Cc * 2, lambda: gemm(du, w1, T, Cc, FF, b_mn=True, out=dyb)),
        ("qkv dgrad bf16    ", 2 * T * QKV * Cc, T * QKV * 2 + T * Cc * 2, lambda: gemm(qkv, wq, T, Cc, QKV, b_mn=True, out=dyb)),
        ("fc2 wgrad         ", 2 * T * FF * Cc, T * FF * 2 + T * Cc * 2,
         lambda: gemm(g, h, Cc, FF, T, a_mn=True, b_mn=True, out=dw2, accumulate=True)),
        ("fc1 wgrad         ", 2 * T * FF * Cc, T * FF * 2 + T * Cc * 2,
         lambda: gemm(du, y, FF, Cc, T, a_mn=True, b_mn=True, out=dw1, accumulate=True)),
        ("qkv wgrad         ", 2 * T * QKV * Cc, T * QKV * 2 + T * Cc * 2,
         lambda: gemm(qkv, y, QKV, Cc, T, a_mn=True, b_mn=True, out=dwq, accumulate=True)),
        ("fc1 wgrad + colsum", 2 * T * FF * Cc, T * FF * 2 + T * Cc * 2,
         lambda: gemm(du, y, FF, Cc, T, a_mn=True, b_mn=True, out=dw1, accumulate=True, a_colsum=db1)),
        ("qkv wgrad + colsum", 2 * T * QKV * Cc, T * QKV * 2 + T * Cc * 2,
         lambda: gemm(qkv, y, QKV, Cc, T, a_mn=True, b_mn=True, out=dwq, accumulate=True, a_colsum=dbq)),
        ("fc1 colsum pass   ", 0, T * FF * 2, lambda: L.check(lib.vtb_colsum_bf16(du.addr, T, FF, FF, db1.addr, None), lib)),
        ("qkv colsum pass   ", 0, T * QKV * 2, lambda: L.check(lib.vtb_colsum_bf16(qkv.addr, T, QKV, QKV, dbq.addr, None), lib)),
    ]
    only = os.environ.get("GEMM_ONLY")
    tot = 0.0
    for name, fl, by, fn in cases:
        if only and not any(o in name for o in only.split(",")):
            continue
        us = timer.time(fn)
        tot += us
        if os.environ.get("GEMM_TRACE"):
            trace_report(name, fn)
        print(f"{tag} {name}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s  "
              f"(floor: {fl / PEAK_TF / 1e6:6.1f} us tensor, {by / PEAK_GB / 1e3:6.1f} us hbm)", flush=True)
    print(f"{tag} total {tot:.1f} us", flush=True)
    for buf in (y, o, g, x, wq, wo, w1, w2, bq, bo, b1, b2, qkv, u, h, du, out, dyb, dw1, dw2, dwq, scale):
        buf.free()


for kv in filter(None, (os.environ.get("GEMM_OPTS") or os.environ.get("VTB_OPTS", "")).split(",")):  # e.g. GEMM_OPTS=gemm_colsum_pair=0
    k, v = kv.split("=")
    L.check(lib.vtb_set_option(k.encode(), int(v)), lib)
if not os.environ.get("GEMM_NOCHECK"):  # the knock-out probes of the debug build produce wrong results on purpose
    self_check()
rng = np.random.default_rng(1)
n_seed = 16 << 20
seed = {BF16: cu.Buf(n_seed, BF16).upload(cu.to_bf16_bits(rng.standard_normal(n_seed, F32))),
        F32: cu.Buf(n_seed, F32).upload(rng.standard_normal(n_seed, F32))}
which = os.environ.get("GEMM_BLOCK", "swin3,vitb").split(",")
if "swin3" in which:
    block(50176, 384, 1536, "swin-s3", seed)
if "vitb" in which:
    block(50432, 768, 3072, "vit-b  ", seed)
if "swin1" in which:
    block(802816, 96, 384, "swin-s1", seed)
print(f"cabi_gemm_bench: done in {time.time() - t0:.1f} s", flush=True)
