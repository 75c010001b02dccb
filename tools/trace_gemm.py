"""Per-phase cycle breakdown of the staged GEMM epilogue (library built with -DVTB_GEMM_TRACE)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-transformers-pytorch_b200")):
    sys.path.insert(0, p)
import torch
from vtb200 import lib, ops
L = lib.get()
F32, BF16 = torch.float32, torch.bfloat16
names = ["-", "-", "tmem_ld_wait", "next ld / aux + free-buffer wait", "math+STS", "fence.proxy", "syncwarp+arrive", "bookkeeping"]
def run(tag, fn, subtiles_per_cta):
    fn(); torch.cuda.synchronize()
    L.vtb_debug_gemm_trace(None, 1)
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    buf = (C.c_ulonglong * 32)()
    L.vtb_debug_gemm_trace(buf, 1)
    v = list(buf)
    tot_i, tot_o = sum(v[:8]), sum(v[8:16])
    tiles_e, tiles_i = max(v[18], 1), max(v[22], 1)
    print(f"== {tag}: {us:.1f} us per launch with the trace build")
    print(f"== {tag}: per tile (CTA 7; {tiles_e / n:.1f} tiles / launch): epilogue warp waits {v[16] / tiles_e:.0f} cycles for the accumulator of "
          f"{v[17] / tiles_e:.0f}; MMA issuer waits {v[20] / tiles_i:.0f} for a TMEM stage + {v[19] / tiles_i:.0f} for operands of {v[21] / tiles_i:.0f}")
    print(f"== {tag}: issuer-lane cycles/sub-tile (4 quarters summed / 4): " + ", ".join(f"{names[i]}={v[i]/n/4/subtiles_per_cta:.0f}" for i in range(8)) + f" | total {tot_i/n/4/subtiles_per_cta:.0f}")
    print(f"   non-issuer warp: " + ", ".join(f"{names[i]}={v[8+i]/n/4/subtiles_per_cta:.0f}" for i in range(8)) + f" | total {tot_o/n/4/subtiles_per_cta:.0f}")
T, Cc, FF = 50176, 384, 1536
y = torch.randn(T, Cc, device="cuda").to(BF16); x = torch.randn(T, Cc, device="cuda")
w1 = torch.randn(FF, Cc, device="cuda").to(BF16); b1 = torch.randn(FF, device="cuda")
wq = torch.randn(3 * Cc, Cc, device="cuda").to(BF16); bq = torch.randn(3 * Cc, device="cuda")
wo = torch.randn(Cc, Cc, device="cuda").to(BF16); bo = torch.randn(Cc, device="cuda")
u = torch.empty(T, FF, dtype=BF16, device="cuda"); h = torch.empty_like(u)
qkv = torch.empty(T, 3 * Cc, dtype=BF16, device="cuda"); out = torch.empty(T, Cc, device="cuda")
o = torch.randn(T, Cc, device="cuda").to(BF16)
# sub-tiles per CTA: tiles per cluster * n_sub
run("fc1 silu-dual [50176x1536x384]", lambda: ops.gemm(y, w1, out=u, out2=h, bias=b1, epilogue=lib.EPI_SILU_DUAL), (196 * 6 / 74) * 4)
run("qkv bf16+bias [50176x1152x384]", lambda: ops.gemm(y, wq, out=qkv, bias=bq), (196 * 5 / 74) * 4)
run("proj f32+resid [50176x384x384]", lambda: ops.gemm(o, wo, out=out, bias=bo, resid=x), (196 * 3 / 74) * 4)
