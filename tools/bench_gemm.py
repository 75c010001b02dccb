"""Micro-benchmark of the tcgen05 GEMM on the block shapes of Swin-S stage 3 and ViT-B (every epilogue variant)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-transformers-pytorch_b200")):
    sys.path.insert(0, p)
import torch
from vtb200 import lib, ops

F32, BF16 = torch.float32, torch.bfloat16
dev = "cuda"


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def block(T, C, FF, tag):
    QKV = 3 * C
    y = torch.randn(T, C, device=dev).to(BF16)
    x = torch.randn(T, C, device=dev)
    wq = torch.randn(QKV, C, device=dev).to(BF16)
    wo = torch.randn(C, C, device=dev).to(BF16)
    w1 = torch.randn(FF, C, device=dev).to(BF16)
    w2 = torch.randn(C, FF, device=dev).to(BF16)
    bq, bo, b1, b2 = (torch.randn(n, device=dev) for n in (QKV, C, FF, C))
    qkv = torch.empty(T, QKV, dtype=BF16, device=dev)
    o = torch.randn(T, C, device=dev).to(BF16)
    u = torch.empty(T, FF, dtype=BF16, device=dev)
    h = torch.empty(T, FF, dtype=BF16, device=dev)
    out = torch.empty(T, C, dtype=F32, device=dev)
    g = torch.randn(T, C, device=dev).to(BF16)
    du = torch.empty(T, FF, dtype=BF16, device=dev)
    dyb = torch.empty(T, C, dtype=BF16, device=dev)
    dw1 = torch.zeros(FF, C, device=dev)
    dw2 = torch.zeros(C, FF, device=dev)
    dwq = torch.zeros(QKV, C, device=dev)
    scale = torch.ones(T // 196 + 1, device=dev)
    cases = [
        ("qkv  fwd bf16+bias", 2 * T * QKV * C, (T * C + T * QKV) * 2, lambda: ops.gemm(y, wq, out=qkv, bias=bq)),
        ("proj fwd f32+resid", 2 * T * C * C, T * C * 2 + T * C * 8, lambda: ops.gemm(o, wo, out=out, bias=bo, resid=x, row_scale=scale, rows_per_scale=196)),
        ("fc1  fwd silu-dual", 2 * T * FF * C, T * C * 2 + T * FF * 4, lambda: ops.gemm(y, w1, out=u, out2=h, bias=b1, epilogue=lib.EPI_SILU_DUAL)),
        ("fc2  fwd f32+resid", 2 * T * FF * C, T * FF * 2 + T * C * 8, lambda: ops.gemm(h, w2, out=out, bias=b2, resid=x, row_scale=scale, rows_per_scale=196)),
        ("fc2 dgrad silugrad", 2 * T * FF * C, T * C * 2 + T * FF * 4, lambda: ops.gemm(g, w2, b_mn=True, out=du, epilogue=lib.EPI_SILU_GRAD, aux=u)),
        ("fc1 dgrad bf16    ", 2 * T * FF * C, T * FF * 2 + T * C * 2, lambda: ops.gemm(du, w1, b_mn=True, out=dyb)),
        ("qkv dgrad bf16    ", 2 * T * QKV * C, T * QKV * 2 + T * C * 2, lambda: ops.gemm(qkv, wq, b_mn=True, out=dyb)),
        ("fc2 wgrad         ", 2 * T * FF * C, T * FF * 2 + T * C * 2, lambda: ops.gemm(g, h, a_mn=True, b_mn=True, out=dw2, accumulate=True)),
        ("fc1 wgrad         ", 2 * T * FF * C, T * FF * 2 + T * C * 2, lambda: ops.gemm(du, y, a_mn=True, b_mn=True, out=dw1, accumulate=True)),
        ("qkv wgrad         ", 2 * T * QKV * C, T * QKV * 2 + T * C * 2, lambda: ops.gemm(qkv, y, a_mn=True, b_mn=True, out=dwq, accumulate=True)),
    ]
    only = os.environ.get("GEMM_ONLY")
    tot = 0.0
    for name, fl, by, fn in cases:
        if only and only not in name:
            continue
        us = timeit(fn)
        tot += us
        print(f"{tag} {name}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s  (floor: {fl / 1392.6e6:6.1f} us tensor, {by / 6.5431e6:6.1f} us hbm)", flush=True)
    print(f"{tag} total {tot:.1f} us")


which = os.environ.get("GEMM_BLOCK", "swin3,vitb").split(",")
if "swin3" in which:
    block(50176, 384, 1536, "swin-s3")
if "vitb" in which:
    block(50432, 768, 3072, "vit-b  ")
if "swin1" in which:
    block(802816, 96, 384, "swin-s1")
