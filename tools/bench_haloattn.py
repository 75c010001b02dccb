"""Micro-benchmark + cross-check of the halo-attention kernels on the Halo-T* stage shapes (B=128): tcgen05 tiles vs mma.sync."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-transformers-pytorch_b200")):
    sys.path.insert(0, p)
import torch
from oracle import restate as R
from vtb200 import lib, ops

B, W, hl, dh = int(os.environ.get("HB", 128)), 7, 3, 32
K = W + 2 * hl
ALL = ((56, 3), (28, 6), (14, 12), (7, 24))
SHAPES = tuple(s for s in ALL if str(s[0]) == os.environ["HT_ONLY"]) if os.environ.get("HT_ONLY") else ALL  # HT_ONLY=56: one shape, tcgen05 only


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


for Hs, H in SHAPES:
    HD, T = H * dh, B * Hs * Hs
    qkv = torch.randn(T, 3 * HD, device="cuda").bfloat16()
    do = torch.randn(T, HD, device="cuda").bfloat16()
    pos = R.halo_pos_table(W, hl)
    table = 0.5 * torch.randn(int(pos.max()) + 1, H, device="cuda")
    spec = ops.AttnSpec(lib.ATTN_HALO, B, H, dh, W * W, K * K, Hs=Hs, Ws=Hs, window=W, halo=hl, rel_bias=table,
                        pos=pos.to(torch.int32).cuda())
    line = f"Hs={Hs:2d} H={H:2d}"
    res = {}
    for mode in ((1,) if os.environ.get("HT_ONLY") else (1, 0)):
        lib.set_option("attn_ht", mode)
        d = torch.empty_like(qkv)
        drel = torch.zeros_like(table)
        def fwd():
            return ops.attention_fwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:])
        o, lse = fwd()
        def bwd():
            drel.zero_()
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
        res[mode] = (o.float(), lse.clone(), d.float(), drel.clone())
        fb = T * H * dh * 2 * 4  # q,k,v read + o write
        bb = T * H * dh * 2 * 8  # q,k,v,o,do read + dq,dk,dv write
        line += f" | {'ht' if mode else 'mma'} fwd {ts[0]*1e3:7.1f} us ({fb/ts[0]/1e6:6.0f} GB/s) bwd {ts[1]*1e3:7.1f} us ({bb/ts[1]/1e6:6.0f} GB/s)"
    lib.set_option("attn_ht", 1)
    if len(res) == 2:
        line += " | rel o %.1e lse %.1e dqkv %.1e drel %.1e" % tuple(rel(res[1][i], res[0][i]) for i in range(4))
    print(line, flush=True)
