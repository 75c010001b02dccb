"""Timing experiments on the tcgen05 halo backward kernel (vtb_set_option("attn_ht_dbg", bits)): which phase bounds a tile?
bits: 1 no global stores, 2 no K/V copies, 4 no softmax math, 8 no gradient MMAs, 16 no score MMAs, 32 no reduce kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-transformers-pytorch_b200")):
    sys.path.insert(0, p)
import torch
from oracle import restate as R
from vtb200 import lib, ops

B, W, hl, dh, Hs, H = 128, 7, 3, 32, 56, 3
K = W + 2 * hl
HD, T = H * dh, B * Hs * Hs
qkv = torch.randn(T, 3 * HD, device="cuda").bfloat16()
do = torch.randn(T, HD, device="cuda").bfloat16()
pos = R.halo_pos_table(W, hl)
table = 0.5 * torch.randn(int(pos.max()) + 1, H, device="cuda")
spec = ops.AttnSpec(lib.ATTN_HALO, B, H, dh, W * W, K * K, Hs=Hs, Ws=Hs, window=W, halo=hl, rel_bias=table, pos=pos.to(torch.int32).cuda())
o, lse = ops.attention_fwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:])
d = torch.empty_like(qkv)
drel = torch.zeros_like(table)
for bits in (0, 32, 32 + 1, 32 + 2, 32 + 4, 32 + 8, 32 + 16, 32 + 8 + 16, 32 + 1 + 2, 32 + 1 + 2 + 4, 63):
    lib.set_option("attn_ht_dbg", bits)
    def bwd():
        ops.attention_bwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, d[:, :HD], d[:, HD:2 * HD], d[:, 2 * HD:], drel)
    for _ in range(3):
        bwd()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        bwd()
    e1.record()
    torch.cuda.synchronize()
    print(f"dbg bits {bits:2d}: {e0.elapsed_time(e1) / 10 * 1e3:7.1f} us", flush=True)
lib.set_option("attn_ht_dbg", 0)
