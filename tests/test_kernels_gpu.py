"""Kernel-level parity on the B200: every C-ABI entry point against the oracle (oracle/restate.py) or a plain
fp32 torch restatement of the same op on identical seeded inputs.  Tolerances (stated per test) are for bf16
operands with fp32 accumulation: outputs stored as bf16 carry 2^-9 relative rounding."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16, F32 = torch.bfloat16, torch.float32


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from vtb200 import ops as o

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return o


def bf(t):
    return t.to(BF16)


# ----------------------------------------------------------------------------------------------- GEMM
@pytest.fixture(params=[0, 2], ids=["tiles_1cta", "tiles_pair"])
def gemm_mode(request):
    """Every GEMM test runs with independent 128-row tiles and with CTA pairs (cta_group::2, 256-row tiles) forced
    wherever legal; the library default (1) picks pairs only when they can fill the machine."""
    from vtb200 import lib

    lib.set_option("gemm_cluster", request.param)
    yield request.param
    lib.set_option("gemm_cluster", 1)


GEMM_SHAPES = [(128, 64, 64), (256, 256, 128), (300, 200, 72), (128, 96, 48), (1000, 768, 768), (197 * 4, 2304, 768),
               (64, 1000, 768), (513, 328, 1096)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("layout", ["nt", "nn", "tt", "tn"])
def test_gemm_layouts(ops, gemm_mode, M, N, K, layout):
    """C = A B^T for all four operand-major combinations (forward / dgrad / wgrad read layouts)."""
    if layout != "nt" and (M % 8 or N % 8):
        pytest.skip("MN-major operands need M, N multiples of 8 (16-byte TMA strides)")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = bf(torch.randn(M, K, device="cuda", generator=g))
    B = bf(torch.randn(N, K, device="cuda", generator=g))
    want = A.float() @ B.float().t()
    a_mn, b_mn = layout[0] == "t", layout[1] == "n"
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=F32)
    torch.cuda.synchronize()
    assert rel(got, want) < 1e-5, (layout, rel(got, want))
    got16 = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=BF16)
    assert rel(got16.float(), want) < 4e-3


@pytest.mark.parametrize("layout", ["nt", "nn", "tt", "tn"])
def test_gemm_cta_pairs_default_heuristic(ops, layout):
    """>= 74 pair tiles -> the default heuristic takes CTA pairs; odd number of row tiles -> one phantom half tile."""
    from vtb200 import lib

    _cluster_case(ops, lib, layout)


def _cluster_case(ops, lib, layout):
    g = torch.Generator(device="cuda").manual_seed(21)
    M, N, K = 128 * 65 - 24, 1024 - 8, 192
    A = bf(torch.randn(M, K, device="cuda", generator=g))
    B = bf(torch.randn(N, K, device="cuda", generator=g) * 0.2)
    want = A.float() @ B.float().t()
    a_mn, b_mn = layout[0] == "t", layout[1] == "n"
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=F32)
    assert rel(got, want) < 1e-5
    if layout == "nt":
        bias = torch.randn(N, device="cuda", generator=g)
        resid = torch.randn(M, N, device="cuda", generator=g)
        got = ops.gemm(a, b, out_dtype=F32, bias=bias, resid=resid)
        assert rel(got, want + bias + resid) < 1e-5
        u = torch.empty(M, N, dtype=BF16, device="cuda")
        h = torch.empty(M, N, dtype=BF16, device="cuda")
        ops.gemm(a, b, out=u, out2=h, bias=bias, epilogue=lib.EPI_SILU_DUAL)
        assert rel(u.float(), want + bias) < 4e-3
        assert rel(h.float(), torch.nn.functional.silu(u.float())) < 4e-3
    if layout in ("tt", "tn"):  # split-K reduce-add through the pair path
        got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=F32, accumulate=True, splits=2)
        assert rel(got, want) < 1e-5


def test_gemm_strided_views(ops, gemm_mode):
    """Operands that are column slices of wider buffers (q/k/v inside the fused qkv buffer)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    big = bf(torch.randn(400, 3 * 128, device="cuda", generator=g))
    W = bf(torch.randn(96, 128, device="cuda", generator=g))
    for s in range(3):
        a = big[:, s * 128:(s + 1) * 128]
        got = ops.gemm(a, W, out_dtype=F32)
        assert rel(got, a.float() @ W.float().t()) < 1e-5


def test_gemm_epilogue_bias_silu_dual(ops, gemm_mode):
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 392, 512, 192
    A, W = bf(torch.randn(M, K, device="cuda", generator=g)), bf(torch.randn(N, K, device="cuda", generator=g) * 0.1)
    bias = torch.randn(N, device="cuda", generator=g)
    u = torch.empty(M, N, dtype=BF16, device="cuda")
    h = torch.empty(M, N, dtype=BF16, device="cuda")
    from vtb200 import lib

    ops.gemm(A, W, out=u, out2=h, bias=bias, epilogue=lib.EPI_SILU_DUAL)
    want_u = (A.float() @ W.float().t() + bias)
    assert rel(u.float(), want_u) < 4e-3
    want_h = torch.nn.functional.silu(u.float())  # silu of the bf16-rounded pre-activation (layer.py:193 under autocast)
    assert rel(h.float(), want_h) < 4e-3


def test_gemm_epilogue_silu_grad(ops, gemm_mode):
    from vtb200 import lib

    g = torch.Generator(device="cuda").manual_seed(2)
    M, N, K = 264, 384, 128
    G, W = bf(torch.randn(M, K, device="cuda", generator=g)), bf(torch.randn(K, N, device="cuda", generator=g) * 0.1)
    u = bf(torch.randn(M, N, device="cuda", generator=g) * 2)
    got = ops.gemm(G, W, b_mn=True, epilogue=lib.EPI_SILU_GRAD, aux=u)
    uf = u.float()
    s = torch.sigmoid(uf)
    want = (G.float() @ W.float()) * (s * (1 + uf * (1 - s)))
    assert got.dtype == BF16 and rel(got.float(), want) < 4e-3  # bf16 store


def test_gemm_epilogue_residual_droppath(ops, gemm_mode):
    g = torch.Generator(device="cuda").manual_seed(3)
    B, n, N, K = 6, 50, 256, 320
    M = B * n
    A, W = bf(torch.randn(M, K, device="cuda", generator=g)), bf(torch.randn(N, K, device="cuda", generator=g) * 0.1)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    scale = torch.tensor([0., 1.25, 1.25, 0., 1.25, 1.25], device="cuda")
    got = ops.gemm(A, W, out_dtype=F32, bias=bias, resid=resid, row_scale=scale, rows_per_scale=n)
    want = resid + (A.float() @ W.float().t() + bias) * scale.repeat_interleave(n)[:, None]
    assert rel(got, want) < 1e-5
    assert torch.equal(got[:n], resid[:n])  # dropped sample: the branch contributes exactly zero


def test_gemm_splitk_accumulate_and_group_rows(ops, gemm_mode):
    g = torch.Generator(device="cuda").manual_seed(4)
    T, N, K = 4096 + 72, 256, 192
    Gd, X = bf(torch.randn(T, N, device="cuda", generator=g)), bf(torch.randn(T, K, device="cuda", generator=g))
    want = Gd.float().t() @ X.float()
    got = ops.gemm(Gd, X, a_mn=True, b_mn=True, out_dtype=F32, accumulate=True)  # auto split-K
    assert rel(got, want) < 1e-5
    got2 = ops.gemm(Gd, X, a_mn=True, b_mn=True, out=got.clone(), accumulate=True, splits=3)
    assert rel(got2, 2 * want) < 1e-5


@pytest.mark.parametrize("T,N,K", [(1000, 384, 96), (50432 // 8, 3072, 768), (197 * 3, 200, 64), (4096, 2304, 768)])
def test_gemm_wgrad_with_fused_bias_gradient(ops, gemm_mode, T, N, K):
    """dW = g^T x and db = column sums of g from ONE launch (a_colsum: the epilogue warps add up the A tiles in smem).
    The last shape gives every CTA several (tile, k-split) work items: the ring is then served between the sub-tile steps
    of a drain (non-blocking state machine) and the partial sums are flushed per tile."""
    gen = torch.Generator(device="cuda").manual_seed(T + N)
    g = bf(torch.randn(T, N, device="cuda", generator=gen))
    x = bf(torch.randn(T, K, device="cuda", generator=gen))
    db = torch.zeros(N, device="cuda")
    dw = ops.gemm(g, x, a_mn=True, b_mn=True, out_dtype=F32, accumulate=True, a_colsum=db)
    assert rel(dw, g.float().t() @ x.float()) < 1e-5
    assert rel(db, g.float().sum(0)) < 1e-5
    # accumulates into the caller's buffer
    ops.gemm(g, x, a_mn=True, b_mn=True, out_dtype=F32, accumulate=True, a_colsum=db)
    assert rel(db, 2 * g.float().sum(0)) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dropout_kernel(ops, dtype):
    """vtb_dropout: out = resid + row_scale[row] * (keep ? x * scale : 0), in place or not, f32 / bf16."""
    g = torch.Generator(device="cuda").manual_seed(21)
    rows, cols, rps = 6 * 17, 40, 17
    x = torch.randn(rows, cols, device="cuda", generator=g).to(dtype)
    keep = torch.rand(rows, cols, device="cuda", generator=g) > 0.3
    scale = 1 / 0.7
    want = torch.where(keep, x.float() * scale, torch.zeros((), device="cuda")).to(dtype)
    assert torch.equal(ops.dropout(x, keep, scale), want)
    y = x.clone()
    assert ops.dropout(y, keep, scale, out=y) is y and torch.equal(y, want)
    if dtype == torch.float32:
        resid = torch.randn(rows, cols, device="cuda", generator=g)
        rs = torch.tensor([0., 2., 2., 0., 2., 2.], device="cuda")
        got = ops.dropout(x, keep, scale, resid=resid, row_scale=rs, rows_per_scale=rps)
        ref = resid + rs.repeat_interleave(rps)[:, None] * torch.where(keep, x * scale, torch.zeros((), device="cuda"))
        assert rel(got, ref) < 1e-6
    else:
        with pytest.raises(ValueError):
            ops.dropout(x, keep, scale, resid=x)


def test_gemm_unaligned_output_falls_back_to_direct_path(ops, gemm_mode):
    """N=10 / N=50 heads: rows are not 16-byte multiples -> per-thread epilogue instead of TMA stores."""
    g = torch.Generator(device="cuda").manual_seed(8)
    for N in (10, 50, 1000):
        A, W = bf(torch.randn(77, 64, device="cuda", generator=g)), bf(torch.randn(N, 64, device="cuda", generator=g))
        bias = torch.randn(N, device="cuda", generator=g)
        got = ops.gemm(A, W, out_dtype=F32, bias=bias)
        assert rel(got, A.float() @ W.float().t() + bias) < 1e-5


def test_vit_assemble_tokens(ops):
    g = torch.Generator(device="cuda").manual_seed(14)
    B, n, D = 3, 16, 64
    tok, cls, pos = (torch.randn(s, device="cuda", generator=g) for s in ((B * n, D), (D,), (n + 1, D)))
    x = ops.vit_assemble_tokens(tok, cls, pos, B, n, D)
    want = torch.cat((cls.expand(B, 1, D), tok.view(B, n, D)), 1) + pos
    assert torch.equal(x, want)


def test_gemm_rejects_bad_arguments(ops):
    A = torch.zeros(16, 12, dtype=BF16, device="cuda")  # K=12 -> lda not a multiple of 8
    with pytest.raises(RuntimeError, match="multiples of 8"):
        ops.gemm(A, A)
    with pytest.raises(ValueError):
        ops.gemm(A.float(), A)


# ----------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,cols", [(197 * 3, 768), (1000, 96), (77, 1536), (64, 32), (5, 384), (4099, 768),
                                       (20001, 96), (3001, 320), (2500, 512), (1003, 192), (9, 64)])
@pytest.mark.parametrize("eps", [1e-6, 1e-5])
def test_layernorm_fwd_bwd(ops, rows, cols, eps):
    from oracle import restate as R

    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    x = (torch.randn(rows, cols, device="cuda", generator=g) * 2 + 0.5).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(cols, device="cuda", generator=g)).requires_grad_(True)
    b = (0.1 * torch.randn(cols, device="cuda", generator=g)).requires_grad_(True)
    y, mean, rstd = ops.layernorm_fwd(x.detach(), w.detach(), b.detach(), eps, out_dtype=F32)
    want = R.layer_norm(x, w, b, eps)
    assert rel(y, want) < 2e-6
    y16, _, _ = ops.layernorm_fwd(x.detach(), w.detach(), b.detach(), eps)
    assert rel(y16.float(), want) < 4e-3
    dy = torch.randn(rows, cols, device="cuda", generator=g)
    want.backward(dy)
    dx_in = torch.randn(rows, cols, device="cuda", generator=g)
    scale = torch.rand(rows, device="cuda", generator=g)
    dx, dxb, dg, db = ops.layernorm_bwd(dy, x.detach(), w.detach(), mean, rstd, dx_in=dx_in, want_bf16=True,
                                        row_scale=scale, rows_per_scale=1)
    assert rel(dx, x.grad + dx_in) < 1e-5
    assert rel(dxb.float(), (x.grad + dx_in) * scale[:, None]) < 4e-3
    assert rel(dg, w.grad) < 1e-4 and rel(db, b.grad) < 1e-4  # fp32 atomics: order-dependent
    dx2, _, _, _ = ops.layernorm_bwd(dy.to(BF16), x.detach(), w.detach(), mean, rstd)
    assert rel(dx2, x.grad) < 5e-3
    if cols <= 768:  # fused column sums of the scaled bf16 copy (the producer Linear's bias gradient)
        cs = torch.zeros(cols, device="cuda")
        _, dxb2, _, _ = ops.layernorm_bwd(dy.to(BF16), x.detach(), w.detach(), mean, rstd, dx_in=dx_in, want_bf16=True,
                                          row_scale=scale, rows_per_scale=1, colsum_out=cs)
        assert rel(cs, dxb2.float().sum(0)) < 1e-4


def test_layernorm_stream_and_register_kernels_agree(ops):
    """The bulk-copy streaming kernels against the register-resident ones (vtb_set_option ln_stream)."""
    from vtb200 import lib as L

    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(5000, 384, device="cuda", generator=g) * 3 + 1
    w = 1 + 0.1 * torch.randn(384, device="cuda", generator=g)
    b = 0.1 * torch.randn(384, device="cuda", generator=g)
    dy = torch.randn(5000, 384, device="cuda", generator=g).to(BF16)
    res = []
    for on in (1, 0):
        L.set_option("ln_stream", on)
        y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6)
        dx, _, dg, db = ops.layernorm_bwd(dy, x, w, mean, rstd, dx_in=x)
        res.append((y.float(), mean, rstd, dx, dg, db))
    L.set_option("ln_stream", 1)
    for a, c in zip(*res):
        assert rel(a, c) < 1e-5


def test_layernorm_patchify_prologue(ops):
    """PatchMerge: patchify(2) folded into the LN row addressing (swin:224-229)."""
    from oracle import restate as R

    g = torch.Generator(device="cuda").manual_seed(9)
    B, H, W, C = 3, 8, 12, 32
    x = torch.randn(B, H, W, C, device="cuda", generator=g).requires_grad_(True)
    w = (1 + 0.1 * torch.randn(4 * C, device="cuda", generator=g))
    b = 0.1 * torch.randn(4 * C, device="cuda", generator=g)
    y, mean, rstd = ops.layernorm_fwd(x.detach(), w, b, 1e-5, out_dtype=F32, patchify=(2, H, W))
    want = R.layer_norm(R.patchify(x, 2), w, b, 1e-5)
    assert rel(y, want.reshape(-1, 4 * C)) < 2e-6
    dy = torch.randn_like(want)
    want.backward(dy)
    dx, _, _, _ = ops.layernorm_bwd(dy.reshape(-1, 4 * C).contiguous(), x.detach(), w, mean, rstd, patchify=(2, H, W))
    assert rel(dx, x.grad) < 1e-5


# ----------------------------------------------------------------------------------------------- attention
def _attn_reference(q, k, v, bias, mask, do):
    """fp32 autograd reference on [G, H, N, dh] tensors."""
    from oracle import restate as R

    q, k, v = (t.float().requires_grad_(True) for t in (q, k, v))
    bias = bias.float().requires_grad_(True) if bias is not None else None
    o = R.softmax_attention(q, k, v, bias, mask)
    o.backward(do.float())
    return o.detach(), q.grad, k.grad, v.grad, (bias.grad if bias is not None else None)


# the last five shapes give every persistent CTA (148 of them) several problems: they exercise the tile ring, the
# prefetch and the chunk pipeline of the tcgen05 kernels for each tiles-per-problem geometry (ViT-B, DeiT 96^2 crops, PVT
# stages 3 / 4, keys > queries) — the single-problem shapes above them never wrap the ring
@pytest.mark.parametrize("B,H,dh,Nq,Nkv", [(2, 3, 64, 197, 197), (3, 2, 32, 37, 37), (2, 1, 64, 300, 49), (2, 5, 64, 50, 50),
                                            (1, 2, 64, 64, 128), (150, 3, 64, 197, 197), (120, 6, 64, 37, 37),
                                            (64, 5, 64, 196, 49), (64, 8, 64, 50, 50), (70, 3, 64, 100, 200),
                                            (9, 1, 64, 3136, 49)])
def test_attention_global(ops, B, H, dh, Nq, Nkv):
    from vtb200 import lib

    g = torch.Generator(device="cuda").manual_seed(Nq * 13 + Nkv)
    HD = H * dh
    qbuf = bf(torch.randn(B * Nq, HD, device="cuda", generator=g))
    kvbuf = bf(torch.randn(B * Nkv, 2 * HD, device="cuda", generator=g))
    spec = ops.AttnSpec(lib.ATTN_GLOBAL, B, H, dh, Nq, Nkv)
    o, lse = ops.attention_fwd(spec, qbuf, kvbuf[:, :HD], kvbuf[:, HD:])
    q4 = qbuf.view(B, Nq, H, dh).permute(0, 2, 1, 3)
    k4 = kvbuf[:, :HD].reshape(B, Nkv, H, dh).permute(0, 2, 1, 3)
    v4 = kvbuf[:, HD:].reshape(B, Nkv, H, dh).permute(0, 2, 1, 3)
    do = bf(torch.randn(B * Nq, HD, device="cuda", generator=g))
    do4 = do.view(B, Nq, H, dh).permute(0, 2, 1, 3)
    wo, wdq, wdk, wdv, _ = _attn_reference(q4, k4, v4, None, None, do4)
    got_o = o.view(B, Nq, H, dh).permute(0, 2, 1, 3).float()
    assert rel(got_o, wo) < 6e-3, rel(got_o, wo)
    want_lse = torch.logsumexp((q4.float() @ k4.float().transpose(-1, -2)) / math.sqrt(dh), -1)
    assert rel(lse, want_lse) < 1e-4
    dq = torch.empty_like(qbuf)
    dkv = torch.empty_like(kvbuf)
    ops.attention_bwd(spec, qbuf, kvbuf[:, :HD], kvbuf[:, HD:], o, lse, do, dq, dkv[:, :HD], dkv[:, HD:])
    assert rel(dq.view(B, Nq, H, dh).permute(0, 2, 1, 3).float(), wdq) < 1.5e-2
    assert rel(dkv[:, :HD].reshape(B, Nkv, H, dh).permute(0, 2, 1, 3).float(), wdk) < 1.5e-2
    assert rel(dkv[:, HD:].reshape(B, Nkv, H, dh).permute(0, 2, 1, 3).float(), wdv) < 1.5e-2


def test_attention_tcgen05_and_mma_paths_agree(ops):
    """The tcgen05/TMEM kernels and the mma.sync kernels implement the same contract (A/B via vtb_set_option)."""
    from vtb200 import lib

    B, H, dh, N = 3, 4, 64, 197
    HD = H * dh
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = bf(torch.randn(B * N, 3 * HD, device="cuda", generator=g))
    do = bf(torch.randn(B * N, HD, device="cuda", generator=g))
    spec = ops.AttnSpec(lib.ATTN_GLOBAL, B, H, dh, N, N)
    res = {}
    for mode in (1, 0):
        lib.set_option("attn_tc", mode)
        try:
            o, lse = ops.attention_fwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:])
            d = torch.empty_like(qkv)
            ops.attention_bwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, d[:, :HD], d[:, HD:2 * HD],
                              d[:, 2 * HD:])
            res[mode] = (o.float(), lse, d.float())
        finally:
            lib.set_option("attn_tc", 1)
    assert rel(res[1][0], res[0][0]) < 6e-3 and rel(res[1][1], res[0][1]) < 1e-4 and rel(res[1][2], res[0][2]) < 1.5e-2


def _window_case(ops, Hs, W, shift, H, dh, B, seed, use_bias=True):
    """Window attention on a fused qkv buffer vs the oracle's gather-form restatement."""
    from oracle import restate as R
    from vtb200 import lib

    g = torch.Generator(device="cuda").manual_seed(seed)
    HD = H * dh
    T = B * Hs * Hs
    qkv = bf(torch.randn(T, 3 * HD, device="cuda", generator=g))
    pos, mask = R.swin_tables(Hs, Hs, W, shift)
    table = (0.5 * torch.randn((2 * W - 1) ** 2, H, device="cuda", generator=g)) if use_bias else None
    spec = ops.AttnSpec(lib.ATTN_WINDOW, B, H, dh, W * W, W * W, Hs=Hs, Ws=Hs, window=W, shift=(W // 2) if shift else 0,
                        rel_bias=table, pos=pos.to(torch.int32).cuda() if use_bias else None,
                        mask=mask.to(torch.uint8).cuda() if shift else None)
    o, lse = ops.attention_fwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:])
    # oracle: identity projections so that swin_window_attention returns the raw attention output
    C3 = 3 * HD
    x = qkv.float().view(B, Hs, Hs, C3).requires_grad_(True)
    tab = (table.clone() if use_bias else torch.zeros((2 * W - 1) ** 2, H, device="cuda")).requires_grad_(True)
    sd = {"a.weight.weight": torch.eye(C3, device="cuda"), "a.weight.bias": torch.zeros(C3, device="cuda"),
          "a.linear.weight": torch.eye(HD, device="cuda"), "a.linear.bias": torch.zeros(HD, device="cuda"),
          "a.rel_pos.weight": tab, "a.pos": pos.cuda()}
    if shift:
        sd["a.local_mask"] = mask.cuda()
    want = R.swin_window_attention(x, sd, "a.", H, dh, W, shift)
    assert rel(o.float().view(B, Hs, Hs, HD), want) < 6e-3
    do = bf(torch.randn(T, HD, device="cuda", generator=g))
    want.backward(do.float().view(B, Hs, Hs, HD))
    dqkv = torch.empty_like(qkv)
    drel = torch.zeros_like(table) if use_bias else None
    ops.attention_bwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, dqkv[:, :HD], dqkv[:, HD:2 * HD],
                      dqkv[:, 2 * HD:], drel)
    assert rel(dqkv.float().view(B, Hs, Hs, C3), x.grad) < 1.5e-2
    if use_bias:
        assert rel(drel, tab.grad) < 1e-2


@pytest.mark.parametrize("Hs,shift", [(14, True), (14, False), (7, True), (28, True)])
def test_attention_window_swin(ops, Hs, shift):
    _window_case(ops, Hs, 7, shift, H=3, dh=32, B=2, seed=Hs + int(shift))


@pytest.mark.parametrize("B,Hs,shift", [(3, 7, True), (2, 14, True), (1, 21, False), (5, 7, False)])
def test_attention_window_tcgen05_and_mma_paths_agree(ops, B, Hs, shift):
    """The tcgen05 window kernels (two windows per 128-row tile) and the mma.sync warp-per-window kernels implement the
    same contract; odd group counts leave half a tile empty."""
    from oracle import restate as R
    from vtb200 import lib

    H, dh, W = 3, 32, 7
    HD, T = H * dh, B * Hs * Hs
    g = torch.Generator(device="cuda").manual_seed(100 + Hs + B)
    qkv = bf(torch.randn(T, 3 * HD, device="cuda", generator=g))
    do = bf(torch.randn(T, HD, device="cuda", generator=g))
    pos, mask = R.swin_tables(Hs, Hs, W, shift)
    table = 0.5 * torch.randn((2 * W - 1) ** 2, H, device="cuda", generator=g)
    spec = ops.AttnSpec(lib.ATTN_WINDOW, B, H, dh, W * W, W * W, Hs=Hs, Ws=Hs, window=W, shift=(W // 2) if shift else 0,
                        rel_bias=table, pos=pos.to(torch.int32).cuda(), mask=mask.to(torch.uint8).cuda() if shift else None)
    res = {}
    for mode in (1, 0):
        lib.set_option("attn_wt", mode)
        try:
            o, lse = ops.attention_fwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:])
            d = torch.empty_like(qkv)
            drel = torch.zeros_like(table)
            ops.attention_bwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, d[:, :HD], d[:, HD:2 * HD],
                              d[:, 2 * HD:], drel)
            res[mode] = (o.float(), lse, d.float(), drel)
        finally:
            lib.set_option("attn_wt", 1)
    assert rel(res[1][0], res[0][0]) < 6e-3 and rel(res[1][1], res[0][1]) < 1e-4
    assert rel(res[1][2], res[0][2]) < 1.5e-2 and rel(res[1][3], res[0][3]) < 1e-2


def test_attention_window_plain_twins(ops):
    _window_case(ops, 14, 7, False, H=2, dh=64, B=2, seed=77, use_bias=False)


@pytest.mark.parametrize("Hs,W,hl,B", [(14, 7, 3, 2), (7, 7, 3, 2), (8, 2, 1, 2), (21, 7, 3, 1), (28, 7, 3, 5), (12, 4, 2, 3),
                                       ((14, 35), 7, 3, 3), ((6, 2), 2, 1, 1)])
@pytest.mark.parametrize("ht", [1, 0])
def test_attention_halo(ops, Hs, W, hl, B, ht):
    """ht=1: tcgen05 halo tiles (forward; backward with per-block partial dK / dV rows + the per-token sum);
    ht=0: the mma.sync kernels.  B=1 x 9 blocks: an odd number of blocks (the last tile holds one); (14, 35) / (6, 2):
    non-square token maps (blocks per row != blocks per column, a map narrower than the halo window)."""
    from vtb200 import lib

    lib.set_option("attn_ht", ht)
    try:
        _halo_case(ops, Hs, W, hl, B)
    finally:
        lib.set_option("attn_ht", 1)


def _halo_case(ops, Hs, W, hl, B):
    from oracle import restate as R
    from vtb200 import lib

    Hs, Ws = Hs if isinstance(Hs, tuple) else (Hs, Hs)
    g = torch.Generator(device="cuda").manual_seed(Hs * 31 + W + Ws)
    H, dh = 3, 32
    HD, T, K = H * dh, B * Hs * Ws, W + 2 * hl
    qkv = bf(torch.randn(T, 3 * HD, device="cuda", generator=g))
    pos = R.halo_pos_table(W, hl)
    table = 0.5 * torch.randn(int(pos.max()) + 1, H, device="cuda", generator=g)
    spec = ops.AttnSpec(lib.ATTN_HALO, B, H, dh, W * W, K * K, Hs=Hs, Ws=Ws, window=W, halo=hl, rel_bias=table,
                        pos=pos.to(torch.int32).cuda())
    o, lse = ops.attention_fwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:])
    C3 = 3 * HD
    x = qkv.float().view(B, Hs, Ws, C3).requires_grad_(True)
    tab = table.clone().requires_grad_(True)
    sd = {"a.weight.weight": torch.eye(C3, device="cuda"), "a.linear.weight": torch.eye(HD, device="cuda"),
          "a.linear.bias": torch.zeros(HD, device="cuda"), "a.rel_pos.weight": tab, "a.pos": pos.cuda()}
    want = R.halo_attention(x, sd, "a.", H, dh, W, hl)
    assert rel(o.float().view(B, Hs, Ws, HD), want) < 6e-3
    do = bf(torch.randn(T, HD, device="cuda", generator=g))
    want.backward(do.float().view(B, Hs, Ws, HD))
    dq = torch.empty(T, HD, dtype=BF16, device="cuda")
    dkv = torch.zeros(T, 2 * HD, dtype=F32, device="cuda")
    drel = torch.zeros_like(table)
    ops.attention_bwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, dq, dkv[:, :HD], dkv[:, HD:], drel,
                      dkv_f32=True)  # query-centric path: fp32 atomics
    xg = x.grad.view(T, C3)
    assert rel(dq.float(), xg[:, :HD]) < 1.5e-2
    assert rel(dkv, xg[:, HD:]) < 1.5e-2
    assert rel(drel, tab.grad) < 1e-2
    # key-centric path: bf16 dK / dV written without atomics by walking the neighbouring query blocks
    d2 = torch.empty(T, 3 * HD, dtype=BF16, device="cuda")
    drel2 = torch.zeros_like(table)
    ops.attention_bwd(spec, qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, d2[:, :HD], d2[:, HD:2 * HD],
                      d2[:, 2 * HD:], drel2)
    assert rel(d2.float(), xg) < 1.5e-2, rel(d2.float(), xg)
    assert rel(drel2, tab.grad) < 1e-2


def test_attention_fully_masked_rows_do_not_nan(ops):
    """Edge case: key padding beyond nkv and a mask that removes most keys."""
    from vtb200 import lib

    B, H, dh, N = 1, 1, 32, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = bf(torch.randn(B * 4, 3 * dh, device="cuda", generator=g))
    mask = torch.ones(1, N, N, dtype=torch.uint8, device="cuda")
    mask[0].fill_diagonal_(0)  # each token only sees itself -> output == v
    pos = torch.zeros(N, N, dtype=torch.int32, device="cuda")
    table = torch.zeros(9, H, device="cuda")
    spec = ops.AttnSpec(lib.ATTN_WINDOW, B, H, dh, N, N, Hs=2, Ws=2, window=2, shift=0, rel_bias=table, pos=pos, mask=mask)
    o, _ = ops.attention_fwd(spec, qkv[:, :dh], qkv[:, dh:2 * dh], qkv[:, 2 * dh:])
    assert torch.isfinite(o.float()).all()
    assert rel(o.float(), qkv[:, 2 * dh:].float()) < 1e-6


# ----------------------------------------------------------------------------------------------- helpers
def test_cast_colsum_scalecast(ops):
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(1003, 328, device="cuda", generator=g)
    assert torch.equal(ops.cast_bf16(x), x.to(BF16))
    s = torch.rand(17, device="cuda", generator=g)
    got = ops.scale_cast_bf16(x[:1003 - 1003 % 59], s, 59)
    want = (x[:1003 - 1003 % 59] * s.repeat_interleave(59)[:, None]).to(BF16)
    assert torch.equal(got, want)
    got2, cs = ops.scale_cast_colsum_bf16(x[:1003 - 1003 % 59], s, 59)
    assert torch.equal(got2, want) and rel(cs, want.float().sum(0)) < 1e-5
    xb = x.to(BF16)
    acc = torch.ones(328, device="cuda")
    ops.colsum(xb, acc)
    assert rel(acc, 1 + xb.float().sum(0)) < 1e-5


@pytest.mark.parametrize("nchw,c_major", [(True, True), (True, False), (False, False), (False, True)])
def test_patch_gather_scatter(ops, nchw, c_major):
    from einops import rearrange

    g = torch.Generator(device="cuda").manual_seed(12)
    B, C, H, W, p = 2, 6, 8, 12, 2
    nhwc = torch.randn(B, H, W, C, device="cuda", generator=g)
    src = nhwc.permute(0, 3, 1, 2).contiguous() if nchw else nhwc
    got = ops.patch_gather(src, nchw=nchw, c_major=c_major, B=B, Cc=C, H=H, W=W, p=p)
    pat = "b (h py) (w px) c -> (b h w) (c py px)" if c_major else "b (h py) (w px) c -> (b h w) (py px c)"
    want = rearrange(nhwc, pat, py=p, px=p)
    assert torch.equal(got, want.to(BF16))
    back = ops.patch_scatter(want.contiguous(), c_major=c_major, B=B, Cc=C, H=H, W=W, p=p)
    assert torch.equal(back, nhwc)  # scatter is the exact inverse permutation


@pytest.mark.parametrize("src_bf16", [True, False])
@pytest.mark.parametrize("B,C,H,W,p", [(2, 64, 16, 24, 8), (3, 320, 14, 14, 2), (1, 8, 4, 4, 4)])
def test_patch_gather_scatter_vectorised_nhwc(ops, src_bf16, B, C, H, W, p):
    """The 16-byte-vector pos-major path (NHWC source, C a multiple of 8 / 4) is an exact permutation both ways."""
    from einops import rearrange

    g = torch.Generator(device="cuda").manual_seed(C + p)
    nhwc = bf(torch.randn(B, H, W, C, device="cuda", generator=g)).float()  # bf16-representable values
    src = nhwc.to(BF16) if src_bf16 else nhwc
    got = ops.patch_gather(src, nchw=False, c_major=False, B=B, Cc=C, H=H, W=W, p=p)
    want = rearrange(nhwc, "b (h py) (w px) c -> (b h w) (py px c)", py=p, px=p)
    assert torch.equal(got, want.to(BF16))
    dA = want.to(BF16).contiguous() if src_bf16 else want.contiguous()
    back = ops.patch_scatter(dA, c_major=False, B=B, Cc=C, H=H, W=W, p=p)
    assert torch.equal(back, nhwc)
    acc = ops.patch_scatter(dA, c_major=False, B=B, Cc=C, H=H, W=W, p=p, dx=back.clone(), accumulate=True)
    assert torch.equal(acc, 2 * nhwc)


def test_pool_fill_rowsum_silu(ops):
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(5, 49, 96, device="cuda", generator=g)
    assert rel(ops.mean_rows_fwd(x, 5, 49, 96), x.mean(1)) < 1e-6
    dy = torch.randn(5, 96, device="cuda", generator=g)
    assert rel(ops.mean_rows_bwd(dy, 5, 49, 96).view(5, 49, 96), (dy / 49)[:, None].expand(5, 49, 96)) < 1e-6
    out = torch.zeros(49, 96, device="cuda")
    ops.rowgroup_sum(x, 49 * 96, 5, 49, 96, out)
    assert rel(out, x.sum(0)) < 1e-6
    buf = torch.zeros(5, 49, 96, device="cuda")
    a, b = torch.randn(96, device="cuda", generator=g), torch.randn(96, device="cuda", generator=g)
    ops.fill_rows(buf, 49 * 96, 5, 96, a, b)
    assert torch.equal(buf[:, 0], (a + b).expand(5, 96)) and torch.all(buf[:, 1:] == 0)
    xs = torch.randn(1000, device="cuda", generator=g).requires_grad_(True)
    ys = torch.nn.functional.silu(xs)
    assert rel(ops.silu_fwd(xs.detach()), ys) < 1e-5
    ys.backward(torch.ones_like(ys))
    assert rel(ops.silu_bwd(xs.detach(), torch.ones_like(xs)), xs.grad) < 1e-4


def test_transpose_hw_and_nchw_scatter(ops):
    g = torch.Generator(device="cuda").manual_seed(15)
    B, H, W, C = 2, 6, 10, 8
    x = torch.randn(B, H, W, C, device="cuda", generator=g)
    assert torch.equal(ops.transpose_hw(x, B, H, W, C), x.transpose(1, 2).contiguous())
    xb = x.to(BF16)
    assert torch.equal(ops.transpose_hw(xb, B, H, W, C), xb.transpose(1, 2).contiguous())
    # Twins scramble (twins.py:70): transposed copy read as NCHW, gathered for a k=s=2 conv
    from oracle import restate as R
    from einops import rearrange

    img = R.twins_scrambled_image(x)  # [B, C, H, W]
    got = ops.patch_gather(ops.transpose_hw(x, B, H, W, C), nchw=True, c_major=True, B=B, Cc=C, H=H, W=W, p=2)
    want = rearrange(img, "b c (h py) (w px) -> (b h w) (c py px)", py=2, px=2)
    assert torch.equal(got, want.to(BF16))
    back = ops.patch_scatter(want.contiguous(), c_major=True, B=B, Cc=C, H=H, W=W, p=2, dst_nchw=True)
    assert torch.equal(back, img)


def test_dwconv3x3_peg(ops):
    from oracle import restate as R

    g = torch.Generator(device="cuda").manual_seed(16)
    B, H, W, C = 2, 7, 9, 32
    x = torch.randn(B, H, W, C, device="cuda", generator=g).requires_grad_(True)
    w = (0.3 * torch.randn(C, 1, 3, 3, device="cuda", generator=g)).requires_grad_(True)
    y = ops.dwconv3x3_fwd(x.detach(), w.detach())
    want = R.peg(x, w)
    assert rel(y, want) < 1e-6
    dy = torch.randn(B, H, W, C, device="cuda", generator=g)
    want.backward(dy)
    dx, dw = ops.dwconv3x3_bwd(x.detach(), w.detach(), dy)
    assert rel(dx, x.grad) < 1e-6 and rel(dw, w.grad) < 1e-5


# ----------------------------------------------------------------------------------------------- DINO loss (§8f)
@pytest.mark.parametrize("B,K,n_crops", [(3, 256, 10), (4, 65536, 10), (2, 1000, 4), (1, 64, 2)])
def test_dino_loss_fused_vs_oracle(ops, B, K, n_crops):
    """vtb_dino_loss (one kernel: loss + gradient) vs the oracle restatement of DINOLoss.forward (loss.py:119-142) and
    torch autograd through it; the drop-in `loss.DINOLoss` module also updates the centre like the reference."""
    from oracle import restate as R
    from vtb200.blocks import DINOLossFn

    g = torch.Generator(device="cuda").manual_seed(B * 7 + n_crops)
    student = (3.0 * torch.randn(n_crops * B, K, device="cuda", generator=g)).requires_grad_(True)
    teacher = 3.0 * torch.randn(2 * B, K, device="cuda", generator=g)
    center = 0.5 * torch.randn(1, K, device="cuda", generator=g)
    want = R.dino_loss(student, teacher, center, n_crops, 0.1, 0.04)
    (gw,) = torch.autograd.grad(want * 1.7, student)
    s2 = student.detach().clone().requires_grad_(True)
    got = DINOLossFn.apply(s2, teacher, center, n_crops, 0.1, 0.04)
    (gg,) = torch.autograd.grad(got * 1.7, s2)
    assert abs(got.item() - want.item()) < 2e-5 * max(1.0, abs(want.item())), (got.item(), want.item())
    assert rel(gg, gw) < 2e-5, rel(gg, gw)
    # module-level drop-in: same constructor as the reference, centre EMA (loss.py:144-152)
    import loss as L

    mod = L.DINOLoss(K, n_crops, 0.04, 0.04, 0, 10).cuda()
    mod.center.copy_(center)
    out = mod(student.detach(), teacher, 0)
    assert abs(out.item() - want.item()) < 2e-5 * max(1.0, abs(want.item()))
    want_center = center * 0.9 + teacher.sum(0, keepdim=True) / teacher.shape[0] * 0.1
    assert rel(mod.center, want_center) < 1e-6


# ----------------------------------------------------------------------------------------------- DINO head rows (a7)
@pytest.mark.parametrize("rows,cols", [(37, 256), (5, 33), (1, 8), (300, 64)])
def test_dino_head_row_kernels_vs_torch(ops, rows, cols):
    """vtb_l2norm_*, vtb_weight_norm_*, vtb_gelu_* against F.normalize / nn.utils.weight_norm / nn.GELU in fp32 autograd."""
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    x = torch.randn(rows, cols, device="cuda", generator=g).requires_grad_()
    dy = torch.randn(rows, cols, device="cuda", generator=g)
    # L2 normalisation (vit.py:259); a zero row exercises the eps clamp
    with torch.no_grad():
        x[0].mul_(0.0 if rows > 1 else 1.0)
    want = torch.nn.functional.normalize(x, dim=-1, p=2)
    (dx_w,) = torch.autograd.grad(want, x, dy)
    yb, inv = ops.l2norm_fwd(x.detach())
    sl = slice(1, None) if rows > 1 else slice(None)
    assert rel(yb.float(), want.detach()) < 3e-3 and (rows == 1 or not yb[0].any())
    dx = ops.l2norm_bwd(dy, x.detach(), inv)
    assert rel(dx[sl], dx_w[sl]) < 1e-5
    # weight norm (vit.py:244-248)
    v = torch.randn(rows, cols, device="cuda", generator=g).requires_grad_()
    gg = (torch.rand(rows, 1, device="cuda", generator=g) + 0.5).requires_grad_()
    w_want = v * (gg / v.norm(dim=1, keepdim=True))
    dv_w, dg_w = torch.autograd.grad(w_want, (v, gg), dy)
    wb, inv_w = ops.weight_norm_fwd(v.detach(), gg.detach())
    assert rel(wb.float(), w_want.detach()) < 3e-3
    dv, dg = ops.weight_norm_bwd(dy, v.detach(), gg.detach(), inv_w)
    assert dg.shape == gg.shape and rel(dv, dv_w) < 1e-5 and rel(dg, dg_w) < 1e-5
    assert ops.weight_norm_bwd(dy, v.detach(), gg.detach(), inv_w, want_dg=False)[1] is None
    # GELU (exact)
    u = (3 * torch.randn(rows, cols, device="cuda", generator=g)).requires_grad_()
    a_want = torch.nn.functional.gelu(u)
    (du_w,) = torch.autograd.grad(a_want, u, dy)
    ab, a = ops.gelu_fwd(u.detach(), want_f32=True)
    assert rel(a, a_want.detach()) < 1e-6 and torch.equal(ab, a.to(torch.bfloat16))
    assert rel(ops.gelu_bwd(u.detach(), dy), du_w) < 1e-5


def test_dino_head_module_vs_oracle(ops):
    """DINOHead forward + every gradient (vit.py:206-262) against the oracle restatement in fp32 autograd."""
    from models.dino import DINOHead
    from oracle import restate as R

    torch.manual_seed(5)
    head = DINOHead(64, 1024, norm_last_layer=False, depth=3, dim_ff=128, dim_bottleneck=32).cuda()
    with torch.no_grad():
        for p in head.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    x = torch.randn(3, 8, 64, device="cuda").requires_grad_()
    dy = torch.randn(3, 8, 1024, device="cuda")
    y = head(x)
    y.backward(dy)
    got = {n: p.grad.clone() for n, p in head.named_parameters()}
    sd = {k: v.detach().clone().requires_grad_() for k, v in head.state_dict().items()}
    xr = x.detach().clone().requires_grad_()
    want = R.dino_head(sd, xr, pre="")
    want.backward(dy)
    assert y.shape == want.shape and rel(y, want) < 1e-2
    assert rel(x.grad, xr.grad) < 3e-2
    for n, gr in got.items():
        assert rel(gr, sd[n].grad) < 3e-2, n
    # norm_last_layer=True (config/dino_deit-s-16.conf keeps weight_g trainable = False): no gradient for it
    head2 = DINOHead(64, 1024, norm_last_layer=True, depth=1, dim_bottleneck=32).cuda()
    head2(x.detach()).sum().backward()
    assert head2.last.weight_g.grad is None and head2.last.weight_v.grad is not None


# ----------------------------------------------------------------------------------------------- DINO vs the reference (a7, f1)
def test_dino_head_module_vs_reference_golden(ops):
    """models.dino.DINOHead (tcgen05 Linears + row kernels) against the reference's OWN DINOHead outputs / gradients
    (tests/golden/dino_ops.pt, oracle/make_dino_golden.py): bf16 operand tolerances."""
    from conftest import load_golden
    from models.dino import DINOHead

    for name, h in load_golden("dino_ops")["heads"].items():
        head = DINOHead(**h["ctor"]).cuda()
        head.load_state_dict(h["state_dict"])
        x = h["x"].cuda().requires_grad_()
        y = head(x)
        assert y.shape == h["output"].shape and rel(y, h["output"].cuda()) < 1e-2, (name, rel(y, h["output"].cuda()))
        (y * h["probe"].cuda()).sum().backward()
        assert rel(x.grad, h["dx"].cuda()) < 3e-2, name
        for k, p in head.named_parameters():
            g = h["grads"][k]
            if g is None:
                assert p.grad is None, (name, k)
            else:
                assert rel(p.grad, g.cuda()) < 3e-2, (name, k, rel(p.grad, g.cuda()))


def test_dino_loss_module_vs_reference_golden(ops):
    """loss.DINOLoss (one fused kernel + centre EMA) against the reference's OWN DINOLoss: loss, student gradient and the
    centre buffer after one and two calls."""
    import loss as L
    from conftest import load_golden

    for name, c in load_golden("dino_ops")["losses"].items():
        mod = L.DINOLoss(*c["ctor"]).cuda()
        mod.center.copy_(c["center0"].cuda())
        s = c["student"].cuda().requires_grad_()
        out = mod(s, c["teacher"].cuda(), c["epoch"])
        assert abs(out.item() - c["loss"].item()) < 2e-5 * abs(c["loss"].item()), (name, out.item(), c["loss"].item())
        out.backward()
        assert rel(s.grad, c["dstudent"].cuda()) < 2e-5, name
        assert rel(mod.center, c["center1"].cuda()) < 1e-6, name
        out2 = mod(c["student2"].cuda(), c["teacher2"].cuda(), c["epoch"])
        assert abs(out2.item() - c["loss2"].item()) < 2e-5 * abs(c["loss2"].item()), name
        assert rel(mod.center, c["center2"].cuda()) < 1e-6, name
