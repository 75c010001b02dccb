"""Micro-benchmark of the window-attention kernels on the Swin-S stage shapes (B=256): tcgen05 tiles vs mma.sync."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-transformers-pytorch_b200")):
    sys.path.insert(0, p)
import torch
from oracle import restate as R
from vtb200 import lib, ops

B, W, dh = 256, 7, 32
SHAPES = ((14, 12),) if os.environ.get("WT_ONLY") else ((56, 3), (28, 6), (14, 12), (7, 24))
for Hs, H in SHAPES:
    for shift in ((True,) if os.environ.get("WT_ONLY") else (True, False)):
        HD, T = H * dh, B * Hs * Hs
        qkv = torch.randn(T, 3 * HD, device="cuda").bfloat16()
        do = torch.randn(T, HD, device="cuda").bfloat16()
        pos, mask = R.swin_tables(Hs, Hs, W, shift)
        table = 0.5 * torch.randn((2 * W - 1) ** 2, H, device="cuda")
        spec = ops.AttnSpec(lib.ATTN_WINDOW, B, H, dh, W * W, W * W, Hs=Hs, Ws=Hs, window=W, shift=(W // 2) if shift else 0,
                            rel_bias=table, pos=pos.to(torch.int32).cuda(), mask=mask.to(torch.uint8).cuda() if shift else None)
        line = f"Hs={Hs:2d} H={H:2d} shift={int(shift)}"
        for mode in ((1,) if os.environ.get("WT_ONLY") else (1, 0)):
            lib.set_option("attn_wt", mode)
            d = torch.empty_like(qkv)
            drel = torch.zeros_like(table)
            def fwd():
                return ops.attention_fwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:])
            o, lse = fwd()
            def bwd():
                ops.attention_bwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, d[:, :HD], d[:, HD:2 * HD], d[:, 2 * HD:], drel)
            ts = []
            for fn in (fwd, bwd):
                for _ in range(3):
                    fn()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 10)
            fb = T * H * dh * 2 * 4  # q,k,v read + o write
            bb = T * H * dh * 2 * 8  # q,k,v,o,do read + dq,dk,dv write
            line += f" | {'wt' if mode else 'wp'} fwd {ts[0]*1e3:7.1f} us ({fb/ts[0]/1e6:6.0f} GB/s) bwd {ts[1]*1e3:7.1f} us ({bb/ts[1]/1e6:6.0f} GB/s)"
        lib.set_option("attn_wt", 1)
        print(line, flush=True)
