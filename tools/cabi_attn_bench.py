"""Torch-free check + micro-benchmark of vtb_attention_fwd / vtb_attention_bwd through the C-ABI (ctypes + numpy +
libcudart; see tools/cudart_ctypes.py): global attention on the ViT-B / DeiT-S shapes and shifted-window attention on the
Swin-S stage shapes.  Each mode is first checked on a small problem against a float64 numpy restatement of the softmax
attention and its gradients, then timed with CUDA events at the BASELINE batch.
  ATTN_ONLY=global|window   ATTN_BATCH=<images>"""
import ctypes as C
import math
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
for kv in filter(None, os.environ.get("VTB_OPTS", "").split(",")):  # e.g. VTB_OPTS=attn_tc_fwd_version=1,attn_tc_bwd_version=1
    name, val = kv.split("=")
    L.check(lib.vtb_set_option(name.encode(), int(val)), lib)
F32, BF16 = np.float32, np.uint16
PEAK_GB = 6543.1


def swin_tables(Hs, Ws, W, shift):
    """pos int32 [W^2, W^2], mask uint8 [nW, W^2, W^2] (1 = masked) — SURVEY A2 (swin_transformer.py:50-101)."""
    s = W // 2 if shift else 0
    masks, pos = [], None
    for wy in range(Hs // W):
        for wx in range(Ws // W):
            ys = np.array([(wy * W + ty + s) % Hs for ty in range(W) for _ in range(W)])
            xs = np.array([(wx * W + tx + s) % Ws for _ in range(W) for tx in range(W)])
            dy, dx = ys[None, :] - ys[:, None], xs[None, :] - xs[:, None]  # [q, k] = k - q
            if shift:
                ok = (np.abs(dy) < W) & (np.abs(dx) < W)
                masks.append(~ok)
                dy, dx = dy * ok, dx * ok
            if pos is None:
                pos = (dy + W - 1) * (2 * W - 1) + (dx + W - 1)
    return pos.astype(np.int32), (np.stack(masks).astype(np.uint8) if shift else None)


def mask_bits(mask, n):
    """[n_mask, 2, 64] int64 row words (include/vtb200.h: mask_bits)."""
    m = (mask[:, :, :n] != 0).astype(np.uint64)
    sh = np.arange(n, dtype=np.uint64)
    bits = np.zeros((mask.shape[0], 2, 64), np.uint64)
    bits[:, 0, :n] = (m << sh[None, None, :]).sum(2)
    bits[:, 1, :n] = (m << sh[None, :, None]).sum(1)
    return bits.view(np.int64)


class Problem:
    """Device buffers + parameter block of one attention call (q, k, v = column slices of a fused [T, 3 H dh] buffer)."""

    def __init__(self, mode, B, H, dh, n, Hs=0, W=0, shift=False, seed=None, host=None, nkv=None):
        self.mode, self.B, self.H, self.dh, self.n, self.Hs, self.W, self.shift = mode, B, H, dh, n, Hs, W, shift
        self.nkv = nkv   # None: keys = queries (fused qkv buffer); else a separate [B * nkv, 2 HD] key / value buffer (PVT SRA)
        HD = H * dh
        self.T = B * n if mode == L.ATTN_GLOBAL else B * Hs * Hs
        self.groups = B if mode == L.ATTN_GLOBAL else B * (Hs // W) ** 2
        mk = (lambda shape, dt: cu.Buf(shape, dt).upload(host[0](shape, dt))) if host else \
             (lambda shape, dt: cu.Buf(shape, dt).fill_from(seed[dt]))
        self.qkv, self.do = mk((self.T, 3 * HD), BF16), mk((self.T, HD), BF16)
        self.o, self.dqkv = cu.Buf((self.T, HD), BF16), cu.Buf((self.T, 3 * HD), BF16)
        if nkv is not None:
            self.kv, self.dkv = mk((B * nkv, 2 * HD), BF16), cu.Buf((B * nkv, 2 * HD), BF16)
        self.lse, self.delta = cu.Buf((self.groups, H, n), F32), cu.Buf((self.groups, H, n), F32)
        self.keep = []
        p = self.p = L.AttnParams()
        p.mode, p.batch, p.heads, p.dh, p.nq, p.nkv = mode, B, H, dh, n, n
        p.scale = 1.0 / math.sqrt(dh)
        if mode == L.ATTN_WINDOW:
            p.Hs, p.Ws, p.window, p.shift = Hs, Hs, W, (W // 2 if shift else 0)
            pos, mask = swin_tables(Hs, Hs, W, shift)
            rng = np.random.default_rng(7)
            self.bias_host = (0.5 * rng.standard_normal(((2 * W - 1) ** 2, H))).astype(F32)
            self.pos_host, self.mask_host = pos, mask
            bias, dpos = cu.Buf(self.bias_host.shape, F32).upload(self.bias_host), cu.Buf(pos.shape, np.int32).upload(pos)
            self.dbias = cu.Buf(self.bias_host.shape, F32).zero()
            self.keep += [bias, dpos]
            p.rel_bias, p.pos, p.n_pos, p.drel_bias = bias.addr, dpos.addr, self.bias_host.shape[0], self.dbias.addr
            if mask is not None:
                dm = cu.Buf(mask.shape, np.uint8).upload(mask)
                db = cu.Buf((mask.shape[0], 2, 64), np.int64).upload(mask_bits(mask, n))
                self.keep += [dm, db]
                p.mask, p.n_mask, p.mask_ld, p.mask_bits = dm.addr, mask.shape[0], mask.shape[2], db.addr
        ld = 3 * HD
        p.q, p.ldq = self.qkv.addr, ld
        p.k, p.ldk = self.qkv.addr + HD * 2, ld
        p.v, p.ldv = self.qkv.addr + 2 * HD * 2, ld
        p.o, p.ldo, p.lse = self.o.addr, HD, self.lse.addr
        p.dout, p.lddo = self.do.addr, HD
        p.dq, p.lddq = self.dqkv.addr, ld
        p.dk, p.lddk = self.dqkv.addr + HD * 2, ld
        p.dv, p.lddv = self.dqkv.addr + 2 * HD * 2, ld
        p.dkv_f32, p.delta = 0, self.delta.addr
        if nkv is not None:
            p.nkv = nkv
            p.k, p.ldk, p.v, p.ldv = self.kv.addr, 2 * HD, self.kv.addr + HD * 2, 2 * HD
            p.dk, p.lddk, p.dv, p.lddv = self.dkv.addr, 2 * HD, self.dkv.addr + HD * 2, 2 * HD

    def fwd(self):
        L.check(lib.vtb_attention_fwd(C.byref(self.p), None), lib)

    def bwd(self):
        L.check(lib.vtb_attention_bwd(C.byref(self.p), None), lib)

    def free(self):
        for b in [self.qkv, self.do, self.o, self.dqkv, self.lse, self.delta] + ([self.kv, self.dkv] if self.nkv is not None else []) + self.keep + ([self.dbias] if hasattr(self, "dbias") else []):
            b.free()


def reference_sra(pr, qkv, kv, do):
    """Separate key / value tokens (PVT spatial-reduction attention): returns o, dq, dkv."""
    B, H, dh, n, m, HD = pr.B, pr.H, pr.dh, pr.n, pr.nkv, pr.H * pr.dh
    qkv, kv, do = qkv.astype(np.float64), kv.astype(np.float64), do.astype(np.float64)
    o, dq, dkv = np.zeros((B * n, HD)), np.zeros((B * n, HD)), np.zeros((B * m, 2 * HD))
    scale = 1.0 / math.sqrt(dh)
    for b in range(B):
        rq, rk = slice(b * n, (b + 1) * n), slice(b * m, (b + 1) * m)
        for h in range(H):
            c = slice(h * dh, (h + 1) * dh)
            q, k, v, g = qkv[rq, c], kv[rk, c], kv[rk, HD + h * dh: HD + (h + 1) * dh], do[rq, c]
            S = q @ k.T * scale
            P = np.exp(S - S.max(1, keepdims=True))
            P /= P.sum(1, keepdims=True)
            o[rq, c] = P @ v
            dP = g @ v.T
            dS = P * (dP - (dP * P).sum(1, keepdims=True))
            dq[rq, c] = dS @ k * scale
            dkv[rk, c] = dS.T @ q * scale
            dkv[rk, HD + h * dh: HD + (h + 1) * dh] = P.T @ g
    return o, dq, dkv


def reference(pr, qkv, do):
    """float64 numpy attention + gradients on the bf16-rounded inputs, in the layout of the device buffers."""
    B, H, dh, n, HD = pr.B, pr.H, pr.dh, pr.n, pr.H * pr.dh
    qkv, do = qkv.astype(np.float64), do.astype(np.float64)
    o, dqkv = np.zeros((pr.T, HD)), np.zeros((pr.T, 3 * HD))
    dbias = np.zeros_like(pr.bias_host, dtype=np.float64) if pr.mode == L.ATTN_WINDOW else None
    if pr.mode == L.ATTN_GLOBAL:
        groups = [(np.arange(b * n, (b + 1) * n), None) for b in range(B)]
    else:
        Hs, W = pr.Hs, pr.W
        s = W // 2 if pr.shift else 0
        groups = []
        for b in range(B):
            for wy in range(Hs // W):
                for wx in range(Hs // W):
                    ys = np.array([(wy * W + ty + s) % Hs for ty in range(W) for _ in range(W)])
                    xs = np.array([(wx * W + tx + s) % Hs for _ in range(W) for tx in range(W)])
                    groups.append(((b * Hs + ys) * Hs + xs, wy * (Hs // W) + wx))
    scale = 1.0 / math.sqrt(dh)
    for rows, wi in groups:
        for h in range(H):
            q, k, v = (qkv[rows, sel * HD + h * dh: sel * HD + (h + 1) * dh] for sel in range(3))
            g = do[rows, h * dh:(h + 1) * dh]
            S = q @ k.T * scale
            if pr.mode == L.ATTN_WINDOW:
                S = S + pr.bias_host[pr.pos_host, h]
                if pr.mask_host is not None:
                    S = np.where(pr.mask_host[wi] != 0, -np.inf, S)
            P = np.exp(S - S.max(1, keepdims=True))
            P /= P.sum(1, keepdims=True)
            o[rows, h * dh:(h + 1) * dh] = P @ v
            dP = g @ v.T
            dS = P * (dP - (dP * P).sum(1, keepdims=True))
            dqkv[rows, h * dh:(h + 1) * dh] = dS @ k * scale
            dqkv[rows, HD + h * dh: HD + (h + 1) * dh] = dS.T @ q * scale
            dqkv[rows, 2 * HD + h * dh: 2 * HD + (h + 1) * dh] = P.T @ g
            if dbias is not None:
                np.add.at(dbias[:, h], pr.pos_host, dS)
    return o, dqkv, dbias


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def mbar_debug(where):
    """Debug builds (-DVTB_MBAR_DEBUG): report the first mbarrier wait that timed out."""
    if hasattr(lib, "vtb_debug_mbar_timeout"):
        out = (C.c_uint32 * 4)()
        lib.vtb_debug_mbar_timeout(out)
        if out[0]:
            print(f"MBAR TIMEOUT after {where}: smem addr 0x{out[0]:x} parity {out[1]} block {out[2]} thread {out[3]}", flush=True)


def self_check(mode, **kw):
    rng = np.random.default_rng(3)
    host = {}

    def gen(shape, dt):
        x = cu.to_bf16_bits(rng.standard_normal(shape, F32))
        host[shape] = cu.from_bf16_bits(x)
        return x

    pr = Problem(mode, host=(gen,), **kw)
    pr.fwd()
    pr.bwd()
    cu.ck(cu.rt.cudaDeviceSynchronize(), "sync")
    mbar_debug("self-check")
    HD = pr.H * pr.dh
    if pr.nkv is not None:
        o, dq, dkv = reference_sra(pr, host[(pr.T, 3 * HD)][:, :HD], host[(pr.B * pr.nkv, 2 * HD)], host[(pr.T, HD)])
        e_o = rel(cu.from_bf16_bits(pr.o.download()), o)
        e_g = max(rel(cu.from_bf16_bits(pr.dqkv.download())[:, :HD], dq), rel(cu.from_bf16_bits(pr.dkv.download()), dkv))
        dbias = None
    else:
        o, dqkv, dbias = reference(pr, host[(pr.T, 3 * HD)], host[(pr.T, HD)])
        e_o = rel(cu.from_bf16_bits(pr.o.download()), o)
        e_g = rel(cu.from_bf16_bits(pr.dqkv.download()), dqkv)
    msg = f"self-check mode={mode} {kw}: o rel-L2 {e_o:.2e}, dqkv rel-L2 {e_g:.2e}"
    ok = e_o < 6e-3 and e_g < 1.5e-2  # the kernel-level tolerances of tests/test_kernels_gpu.py (bf16 P and outputs)
    if dbias is not None:
        e_b = rel(pr.dbias.download(), dbias)
        msg += f", drel_bias rel-L2 {e_b:.2e}"
        ok = ok and e_b < 1.5e-2
    print(("PASS " if ok else "FAIL ") + msg, flush=True)
    pr.free()
    return ok


def dump_trace(path):
    """Debug builds (-DVTB_ATTN_TRACE): timeline of block 0 of the last backward launch."""
    if not hasattr(lib, "vtb_debug_attn_trace"):
        return
    out = (C.c_uint32 * (16 * 2 * 1024))()
    lib.vtb_debug_attn_trace(out)
    ev = sorted((out[(w * 1024 + i) * 2 + 1], w, out[(w * 1024 + i) * 2]) for w in range(16) for i in range(1024)
                if out[(w * 1024 + i) * 2])
    with open(path, "w") as fh:
        t0 = ev[0][0] if ev else 0
        for t, w, e in ev:
            fh.write(f"{(t - t0) & 0xffffffff:10d} warp {w:2d} ev {e}\n")
    print(f"trace: {len(ev)} events -> {path}", flush=True)


def bench(tag, mode, seed, **kw):
    pr = Problem(mode, seed=seed, **kw)
    timer = cu.Timer()
    dump_trace("/dev/null")
    pr.fwd()
    cu.ck(cu.rt.cudaDeviceSynchronize(), "sync after fwd")
    dump_trace(os.path.join(ROOT, "gpurun_out", "attn_trace_fwd_" + tag.split()[1] + "_" + str(kw.get("n")) + ".txt"))
    mbar_debug(tag + " fwd")
    dump_trace("/dev/null")
    pr.bwd()
    cu.ck(cu.rt.cudaDeviceSynchronize(), "sync after bwd")
    dump_trace(os.path.join(ROOT, "gpurun_out", "attn_trace_" + tag.split()[1] + "_" + str(kw.get("n")) + ".txt"))
    mbar_debug(tag + " bwd")
    tf, tb = timer.time(pr.fwd), timer.time(pr.bwd)
    mbar_debug(tag + " timing loops")
    T, HD = pr.T, pr.H * pr.dh
    fb, bb = T * HD * 2 * 4, T * HD * 2 * 8  # q, k, v read + o written;  q, k, v, o, do read + dq, dk, dv written
    fl = 4.0 * pr.groups * pr.H * pr.n * pr.n * pr.dh
    print(f"{tag}: fwd {tf:7.1f} us ({fb / tf / 1e3:6.0f} GB/s = {fb / tf / 1e3 / PEAK_GB * 100:4.1f} %, {fl / tf / 1e6:6.1f} TFLOP/s)  "
          f"bwd {tb:7.1f} us ({bb / tb / 1e3:6.0f} GB/s = {bb / tb / 1e3 / PEAK_GB * 100:4.1f} %, {2.5 * fl / tb / 1e6:6.1f} TFLOP/s)", flush=True)
    pr.free()


only = os.environ.get("ATTN_ONLY", "")
Bn = int(os.environ.get("ATTN_BATCH", "256"))
ok = True
if os.environ.get("ATTN_SKIP_CHECK"):
    only_check, only = "none", only
else:
    only_check = only
if only_check in ("", "global"):
    ok &= self_check(L.ATTN_GLOBAL, B=2, H=2, dh=64, n=197)
    ok &= self_check(L.ATTN_GLOBAL, B=3, H=1, dh=64, n=37)
    for kw in (dict(B=5, H=3, n=197), dict(B=5, H=1, n=37), dict(B=2, H=2, n=50), dict(B=1, H=2, n=256), dict(B=2, H=1, n=128),
               dict(B=1, H=3, n=129), dict(B=1, H=1, n=16), dict(B=2, H=2, n=145), dict(B=150, H=2, n=197), dict(B=100, H=6, n=37), dict(B=70, H=5, n=128),
               dict(B=33, H=7, n=200), dict(B=64, H=8, n=50), dict(B=40, H=8, n=64), dict(B=90, H=5, n=130),
               dict(B=3, H=2, n=196, nkv=49), dict(B=64, H=5, n=196, nkv=49), dict(B=20, H=2, n=784, nkv=49), dict(B=70, H=3, n=100, nkv=200),
               dict(B=9, H=1, n=3136, nkv=49)):
        ok &= self_check(L.ATTN_GLOBAL, dh=64, **kw)
if only_check in ("", "window"):
    ok &= self_check(L.ATTN_WINDOW, B=1, H=2, dh=32, n=49, Hs=14, W=7, shift=True)
    ok &= self_check(L.ATTN_WINDOW, B=2, H=3, dh=32, n=49, Hs=14, W=7, shift=False)
if not ok:
    raise SystemExit("FAIL: attention self-check")
rng = np.random.default_rng(1)
n_seed = 16 << 20
seed = {BF16: cu.Buf(n_seed, BF16).upload(cu.to_bf16_bits(rng.standard_normal(n_seed, F32))),
        F32: cu.Buf(n_seed, F32).upload(rng.standard_normal(n_seed, F32))}
if only in ("", "global") and not os.environ.get("ATTN_N37_ONLY"):
    bench(f"global ViT-B   B={Bn} H=12 n=197", L.ATTN_GLOBAL, seed, B=Bn, H=12, dh=64, n=197)
    bench(f"global DeiT-S  B={Bn} H=6  n=197", L.ATTN_GLOBAL, seed, B=Bn, H=6, dh=64, n=197)
if only in ("", "global"):
    for Bx in ((Bn // 2, Bn, 2 * Bn, 4 * Bn) if os.environ.get("ATTN_N37_ONLY") else (4 * Bn,)):
        bench(f"global DeiT-S  B={Bx} H=6  n=37 ", L.ATTN_GLOBAL, seed, B=Bx, H=6, dh=64, n=37)
if only in ("", "window"):
    for Hs, H in ((56, 3), (28, 6), (14, 12), (7, 24)):
        for shift in (True, False):
            if Hs == 7 and shift:
                continue  # the 7 x 7 stage has one window: its shifted layer carries no mask, same kernel path as unshifted
            bench(f"window Swin-S  B={Bn} Hs={Hs:2d} H={H:2d} shift={int(shift)}", L.ATTN_WINDOW, seed, B=Bn, H=H, dh=32, n=49,
                  Hs=Hs, W=7, shift=shift)
print(f"cabi_attn_bench: done in {time.time() - t0:.1f} s", flush=True)
