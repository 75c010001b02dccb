"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — CPU/GPU-agnostic restatement, in plain PyTorch
floating point, of the reference's transformer-block path.  It is written from the index maps of
SURVEY.md Appendix A (explicit gathers / einsums over a reference-layout state_dict), NOT from the
reference's module code, so that agreement with the reference (tests/test_oracle_vs_reference.py, run in
the build container) and with the committed golden vectors (tests/golden/, generated from the real
reference by oracle/make_golden.py) is a genuine cross-check.  The oracle's backward is torch autograd of
these functions.

Parity status: PINNED — every family below is checked against (a) the real reference modules imported by
oracle/ref_loader.py and (b) the golden fixtures.  The reference itself ships no tests or vectors
(SURVEY §4), so the pins are generated, not inherited.

Each function cites the reference lines it restates (paths relative to the reference root).
"""
import contextlib
import math

import torch
import torch.nn.functional as F
from einops import rearrange


# ------------------------------------------------------------------------------------------ primitives
def layer_norm(x, w, b, eps):
    """nn.LayerNorm over the last dim (vit.py:13 etc.)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def silu(x):
    return x * torch.sigmoid(x)


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


class _ElementDropout:
    masks, sites, i = None, frozenset(), 0


_ED = _ElementDropout()


@contextlib.contextmanager
def element_dropout(masks, sites):
    """Replays nn.Dropout keep masks — (keep, 1 / (1 - p)) pairs in the reference's call order — at the named sites:
    "ffn" (layer.py:194), "branch" (vit.py:60-61), "pos" (vit.py:146), "patch" (pvt.py:141).  Outside this context every
    Dropout is the identity (p = 0 / eval mode)."""
    _ED.masks, _ED.sites, _ED.i = list(masks), frozenset(sites), 0
    try:
        yield
        assert _ED.i == len(_ED.masks), f"{len(_ED.masks) - _ED.i} dropout masks were never consumed"
    finally:
        _ED.masks, _ED.sites, _ED.i = None, frozenset(), 0


def _drop(x, site):
    if _ED.masks is None or site not in _ED.sites:
        return x
    keep, scale = _ED.masks[_ED.i]
    _ED.i += 1
    return x * keep.reshape(x.shape).to(x.dtype) * scale


def ffn(x, sd, pre):
    """PositionwiseFeedForward: Linear(idx 0) - SiLU - Dropout - Linear(idx 3)  (layer.py:186-196)."""
    hidden = _drop(silu(linear(x, sd[pre + "0.weight"], sd[pre + "0.bias"])), "ffn")
    return linear(hidden, sd[pre + "3.weight"], sd[pre + "3.bias"])


def _dp(branch, scale):
    """DropPath with a given per-sample scale = mask/keep (layer.py:172-178); None = identity."""
    if scale is None:
        return branch
    return branch * scale.view(-1, *([1] * (branch.dim() - 1))).to(branch.dtype)


class DropPathScales:
    """Hands out per-branch DropPath scales in call order (so product and oracle can share masks)."""

    def __init__(self, scales=None):
        self.scales = list(scales) if scales is not None else None
        self.i = 0

    def next(self):
        if self.scales is None:
            return None
        s = self.scales[self.i]
        self.i += 1
        return s


def softmax_attention(q, k, v, bias=None, mask=None):
    """q [..., Nq, dh], k/v [..., Nk, dh]; S = q k^T / sqrt(dh) (+bias) (masked -> -inf); softmax; @ v.
    Scale is applied AFTER the product (vit.py:37, swin:134, pvt:56, halo:94)."""
    s = (q @ k.transpose(-2, -1)) / math.sqrt(q.shape[-1])
    if bias is not None:
        s = s + bias
    if mask is not None:
        s = s.masked_fill(mask, float("-inf"))
    return torch.softmax(s, -1) @ v


# ------------------------------------------------------------------------------------------ ViT
def vit_patch_tokens(img, w, b, p):
    """Conv2d(k=s=p) as a GEMM over (c, py, px)-ordered patch vectors (vit.py:73-76)."""
    a = rearrange(img, "b c (h py) (w px) -> b (h w) (c py px)", py=p, px=p)
    return a @ w.reshape(w.shape[0], -1).t() + b


def vit_pos_embed(pos_embed, n_patch):
    """Bicubic resample of the grid part when #patches differs (vit.py:153-175)."""
    n_pos = pos_embed.shape[1] - 1
    if n_patch == n_pos:
        return pos_embed
    side, new = int(math.sqrt(n_pos)), int(round(math.sqrt(n_patch)))
    D = pos_embed.shape[-1]
    grid = pos_embed[:, 1:].reshape(1, side, side, D).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, scale_factor=math.sqrt(n_patch / n_pos), mode="bicubic", align_corners=False,
                         recompute_scale_factor=False)
    assert grid.shape[-1] == new
    return torch.cat((pos_embed[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, D)), 1)


def mhsa_global(x, sd, pre, heads):
    """vit.py:27-45: fused qkv columns ordered sel*(H*dh) + h*dh + d."""
    B, N, D = x.shape
    qkv = linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"])
    q, k, v = rearrange(qkv, "b n (s h d) -> s b h n d", s=3, h=heads)
    o = rearrange(softmax_attention(q, k, v), "b h n d -> b n (h d)")
    return linear(o, sd[pre + "linear.weight"], sd[pre + "linear.bias"])


def vit_layer(x, sd, pre, heads, dps):
    """vit.py:59-63."""
    x = x + _dp(_drop(mhsa_global(layer_norm(x, sd[pre + "norm_attn.weight"], sd[pre + "norm_attn.bias"], 1e-6), sd,
                                  pre + "attn.", heads), "branch"), dps.next())
    x = x + _dp(_drop(ffn(layer_norm(x, sd[pre + "norm_ff.weight"], sd[pre + "norm_ff.bias"], 1e-6), sd, pre + "ff."),
                      "branch"), dps.next())
    return x


def vit_features(sd, img, *, patch, depth, heads, dp_scales=None):
    """VisionTransformer.forward_feature (vit.py:139-151) for one resolution group."""
    dps = DropPathScales(dp_scales)
    tok = vit_patch_tokens(img, sd["patch_embedding.linear.weight"], sd["patch_embedding.linear.bias"], patch)
    x = torch.cat((sd["cls_token"].expand(tok.shape[0], -1, -1), tok), 1)
    x = _drop(x + vit_pos_embed(sd["pos_embed"], tok.shape[1]), "pos")
    for i in range(depth):
        x = vit_layer(x, sd, f"layers.{i}.", heads, dps)
    x = layer_norm(x, sd["norm.weight"], sd["norm.bias"], 1e-6)
    return x[:, 0]


def vit_forward(sd, inputs, *, patch, depth, heads, head_fn=None, dp_scales=None):
    """VisionTransformer.forward incl. multi-crop grouping by consecutive equal width (vit.py:177-203)."""
    if not isinstance(inputs, (list, tuple)):
        inputs = [inputs]
    groups, start = [], 0
    for i in range(1, len(inputs) + 1):
        if i == len(inputs) or inputs[i].shape[-1] != inputs[start].shape[-1]:
            groups.append(torch.cat(inputs[start:i]))
            start = i
    assert dp_scales is None or len(groups) == 1, "shared DropPath scales are per resolution group"
    out = torch.cat([vit_features(sd, g, patch=patch, depth=depth, heads=heads, dp_scales=dp_scales)
                     for g in groups])
    return head_fn(out) if head_fn is not None else out


def dino_head(sd, x, pre="head."):
    """DINOHead.forward (vit.py:257-262): MLP with exact GELU, L2-normalise, weight-normed Linear (no bias)."""
    if f"{pre}mlp.weight" in sd:   # depth 1: the MLP is a single Linear (vit.py:219-220)
        x = linear(x, sd[f"{pre}mlp.weight"], sd[f"{pre}mlp.bias"])
    i = 0
    while f"{pre}mlp.{i}.weight" in sd:
        x = linear(x, sd[f"{pre}mlp.{i}.weight"], sd[f"{pre}mlp.{i}.bias"])
        if f"{pre}mlp.{i + 2}.weight" in sd:
            x = 0.5 * x * (1 + torch.erf(x / math.sqrt(2.0)))
        i += 2
    x = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    v, g = sd[pre + "last.weight_v"], sd[pre + "last.weight_g"]
    return x @ (v * (g / v.norm(dim=1, keepdim=True))).t()


def dino_loss(student, teacher, center, n_crops, t_student=0.1, t_teacher=0.04):
    """DINOLoss.forward (loss.py:119-142): cross-entropy between the centred/sharpened teacher distribution of each
    global crop and the student distribution of every OTHER crop, averaged over the pairs."""
    s = (student / t_student).chunk(n_crops)
    t = torch.softmax((teacher - center) / t_teacher, -1).detach().chunk(2)
    total, n = 0.0, 0
    for iq, q in enumerate(t):
        for v in range(n_crops):
            if v == iq:
                continue
            total = total + (-q * torch.log_softmax(s[v], -1)).sum(-1).mean()
            n += 1
    return total / n


# ------------------------------------------------------------------------------------------ patchify
def patchify(x, s):
    """NHWC block gather, feature order (sy, sx, c) (swin:15-22)."""
    return rearrange(x, "b (h sy) (w sx) c -> b h w (sy sx c)", sy=s, sx=s)


# ------------------------------------------------------------------------------------------ Swin
def swin_tables(Hs, Ws, W, shift):
    """pos [W^2, W^2] and mask [nW, W^2, W^2] (True = masked) from the formulas of SURVEY A2
    (swin_transformer.py:50-101), built with explicit loops over window 0 / all windows."""
    s = W // 2 if shift else 0
    ny, nx = Hs // W, Ws // W

    def coords(wy, wx):
        ys = [(wy * W + ty + s) % Hs for ty in range(W) for _ in range(W)]
        xs = [(wx * W + tx + s) % Ws for _ in range(W) for tx in range(W)]
        return torch.tensor(ys), torch.tensor(xs)

    masks = []
    pos = None
    for wy in range(ny):
        for wx in range(nx):
            ys, xs = coords(wy, wx)
            dy = ys[None, :] - ys[:, None]  # [q, k] = k - q
            dx = xs[None, :] - xs[:, None]
            if shift:
                ok = (dy.abs() < W) & (dx.abs() < W)
                masks.append(~ok)
                dy, dx = dy * ok, dx * ok
            if pos is None:
                pos = (dy + W - 1) * (2 * W - 1) + (dx + W - 1)
    return pos, (torch.stack(masks) if shift else None)


def swin_window_attention(x, sd, pre, heads, dh, W, shift):
    """swin_transformer.py:103-160 in gather form: for window (wy,wx), token (ty,tx) lives at
    pixel ((wy*W+ty+s) % Hs, (wx*W+tx+s) % Ws) of the UN-rolled map, for reading and for writing back."""
    B, Hs, Ws, C = x.shape
    s = W // 2 if shift else 0
    ny, nx = Hs // W, Ws // W
    qkv = linear(x, sd[pre + "weight.weight"], sd[pre + "weight.bias"])
    iy = ((torch.arange(ny, device=x.device)[:, None] * W + torch.arange(W, device=x.device)[None, :] + s) % Hs)
    ix = ((torch.arange(nx, device=x.device)[:, None] * W + torch.arange(W, device=x.device)[None, :] + s) % Ws)
    Y = iy[:, None, :, None].expand(ny, nx, W, W)
    X = ix[None, :, None, :].expand(ny, nx, W, W)
    win = qkv[:, Y, X]  # [B, ny, nx, W, W, 3*H*dh]
    q, k, v = rearrange(win, "b wy wx ty tx (s h d) -> s b (wy wx) h (ty tx) d", s=3, h=heads)
    bias = sd[pre + "rel_pos.weight"][sd[pre + "pos"]]  # [W^2, W^2, H]
    bias = bias.permute(2, 0, 1)[None, None]
    mask = sd[pre + "local_mask"][None, :, None] if shift else None
    o = softmax_attention(q, k, v, bias, mask)  # [B, nW, H, W^2, dh]
    o = rearrange(o, "b (wy wx) h (ty tx) d -> b wy wx ty tx (h d)", wy=ny, ty=W)
    out = torch.zeros(B, Hs, Ws, heads * dh, dtype=o.dtype, device=o.device)
    out = out.index_put((torch.arange(B, device=x.device)[:, None, None, None, None], Y[None], X[None]), o)
    return linear(out, sd[pre + "linear.weight"], sd[pre + "linear.bias"])


def swin_layer(x, sd, pre, heads, dh, W, shift, dps):
    """swin_transformer.py:193-197."""
    x = x + _dp(swin_window_attention(layer_norm(x, sd[pre + "norm_attn.weight"], sd[pre + "norm_attn.bias"], 1e-6),
                                      sd, pre + "attn.", heads, dh, W, shift), dps.next())
    x = x + _dp(ffn(layer_norm(x, sd[pre + "norm_ff.weight"], sd[pre + "norm_ff.bias"], 1e-6), sd, pre + "ff."),
                dps.next())
    return x


def swin_forward(sd, img, *, depths, n_heads, dim_head, window, dp_scales=None):
    """SwinTransformer.forward (swin_transformer.py:370-379); shift on even layer index (:362)."""
    dps = DropPathScales(dp_scales)
    x = patchify(img.permute(0, 2, 3, 1), 4)
    x = linear(x, sd["patch_embedding.linear.weight"], sd["patch_embedding.linear.bias"])
    x = layer_norm(x, sd["patch_embedding.norm.weight"], sd["patch_embedding.norm.bias"], 1e-5)
    for st in range(4):
        pre = f"block{st + 1}."
        off = 0
        if st > 0:
            x = patchify(x, 2)
            x = layer_norm(x, sd[pre + "0.norm.weight"], sd[pre + "0.norm.bias"], 1e-5)
            x = linear(x, sd[pre + "0.linear.weight"])
            off = 1
        for i in range(depths[st]):
            x = swin_layer(x, sd, f"{pre}{i + off}.", n_heads[st], dim_head, window, i % 2 == 0, dps)
    x = layer_norm(x, sd["final_linear.0.weight"], sd["final_linear.0.bias"], 1e-5)
    x = x.mean((1, 2))
    return linear(x, sd["classifier.2.weight"], sd["classifier.2.bias"])


# ------------------------------------------------------------------------------------------ PVT
def conv_patch_nhwc(x, w, b, p):
    """Conv2d(k=s=p) applied to the NCHW view of NHWC data, result as NHWC tokens (pvt.py:44-45,132)."""
    a = rearrange(x, "b (h py) (w px) c -> b h w c py px", py=p, px=p)
    return torch.einsum("bhwcyx,ocyx->bhwo", a, w) + b


def pvt_attention(x, sd, pre, heads, R, Hs, Ws):
    """pvt.py:32-69 (returns only the projected output; callers use [0], pvt.py:98)."""
    B, N, C = x.shape
    q = rearrange(linear(x, sd[pre + "linear_q.weight"]), "b n (h d) -> b h n d", h=heads)
    if R > 1:
        red = conv_patch_nhwc(x.reshape(B, Hs, Ws, C), sd[pre + "reduce_conv.weight"], sd[pre + "reduce_conv.bias"], R)
        kvin = layer_norm(red.reshape(B, -1, C), sd[pre + "reduce_norm.weight"], sd[pre + "reduce_norm.bias"], 1e-6)
    else:
        kvin = x
    k, v = rearrange(linear(kvin, sd[pre + "linear_kv.weight"]), "b n (s h d) -> s b h n d", s=2, h=heads)
    o = rearrange(softmax_attention(q, k, v), "b h n d -> b n (h d)")
    return linear(o, sd[pre + "linear.weight"], sd[pre + "linear.bias"])


def pvt_forward(sd, img, *, depths, n_heads, reductions, dp_scales=None):
    """PyramidVisionTransformer.forward (pvt.py:255-280)."""
    dps = DropPathScales(dp_scales)
    B = img.shape[0]
    x = img.permute(0, 2, 3, 1)  # NHWC view of the stage input
    for st in range(4):
        pe = f"patch_embedding.{st}."
        p = (4, 2, 2, 2)[st]
        x = conv_patch_nhwc(x, sd[pe + "conv.weight"], sd[pe + "conv.bias"], p)
        Hs, Ws = x.shape[1], x.shape[2]
        x = layer_norm(x.reshape(B, Hs * Ws, -1), sd[pe + "norm.weight"], sd[pe + "norm.bias"], 1e-6)
        if pe + "cls_token" in sd:
            x = torch.cat((sd[pe + "cls_token"].view(1, 1, -1).expand(B, -1, -1), x), 1)
        x = _drop(x + sd[pe + "pos"][None], "patch")
        for i in range(depths[st]):
            pre = f"block{st + 1}.{i}."
            x = x + _dp(pvt_attention(layer_norm(x, sd[pre + "norm_attn.weight"], sd[pre + "norm_attn.bias"], 1e-6),
                                      sd, pre + "attn.", n_heads[st], reductions[st], Hs, Ws), dps.next())
            x = x + _dp(ffn(layer_norm(x, sd[pre + "norm_ff.weight"], sd[pre + "norm_ff.bias"], 1e-6), sd,
                            pre + "ff."), dps.next())
        if st < 3:
            x = x.reshape(B, Hs, Ws, -1)
    x = layer_norm(x[:, 0], sd["norm.weight"], sd["norm.bias"], 1e-6)
    return linear(x, sd["classifier.weight"], sd["classifier.bias"])


# ------------------------------------------------------------------------------------------ Halo
def halo_pos_table(W, hl):
    """pos[t, j] per SURVEY A3 (halo_transformer.py:41-55), explicit loops."""
    K = W + 2 * hl
    off = W + hl - 1
    rows = []
    for ty in range(W):
        for tx in range(W):
            rows.append([(ky - (ty + hl) + off) * K + (kx - (tx + hl) + off) for ky in range(K) for kx in range(K)])
    return torch.tensor(rows)


def halo_attention(x, sd, pre, heads, dh, W, hl):
    """halo_transformer.py:57-114 in gather form: key slot (ky,kx) of block (by,bx) is pixel
    (by*W-hl+ky, bx*W-hl+kx); outside the map k = v = 0 but the slot stays in the softmax (logit = bias)."""
    B, Hs, Ws, C = x.shape
    ny, nx = Hs // W, Ws // W
    K = W + 2 * hl
    qkv = linear(x, sd[pre + "weight.weight"])
    HD = heads * dh
    qm, km, vm = qkv[..., :HD], qkv[..., HD:2 * HD], qkv[..., 2 * HD:]
    q = rearrange(qm, "b (by ty) (bx tx) (h d) -> b (by bx) h (ty tx) d", ty=W, tx=W, h=heads)
    dev = x.device
    ky = torch.arange(ny, device=dev)[:, None] * W + torch.arange(K, device=dev)[None, :]  # index into padded map
    kx = torch.arange(nx, device=dev)[:, None] * W + torch.arange(K, device=dev)[None, :]
    Y = ky[:, None, :, None].expand(ny, nx, K, K)
    X = kx[None, :, None, :].expand(ny, nx, K, K)

    def gather(m):
        mp = F.pad(m, (0, 0, hl, hl, hl, hl))  # zero-pad W then H
        return rearrange(mp[:, Y, X], "b by bx ky kx (h d) -> b (by bx) h (ky kx) d", h=heads)

    bias = sd[pre + "rel_pos.weight"][sd[pre + "pos"]].permute(2, 0, 1)[None, None]  # [1,1,H,W^2,K^2]
    o = softmax_attention(q, gather(km), gather(vm), bias)
    o = rearrange(o, "b (by bx) h (ty tx) d -> b (by ty) (bx tx) (h d)", by=ny, ty=W)
    return linear(o, sd[pre + "linear.weight"], sd[pre + "linear.bias"])


def halo_forward(sd, img, *, depths, n_heads, dim_head, window, halo, dp_scales=None):
    """HaloTransformer.forward (halo_transformer.py:271-280) with the out-of-place residual restatement of
    :147-148 (the reference's in-place `+=` makes its own autograd raise)."""
    dps = DropPathScales(dp_scales)
    x = img.permute(0, 2, 3, 1)
    for st in range(4):
        pre = f"block{st + 1}."
        x = patchify(x, (4, 2, 2, 2)[st])
        x = linear(x, sd[pre + "0.linear.weight"], sd[pre + "0.linear.bias"])
        x = layer_norm(x, sd[pre + "0.norm.weight"], sd[pre + "0.norm.bias"], 1e-5)
        for i in range(depths[st]):
            lp = f"{pre}{i + 1}."
            x = x + _dp(halo_attention(layer_norm(x, sd[lp + "norm_attn.weight"], sd[lp + "norm_attn.bias"], 1e-6),
                                       sd, lp + "attn.", n_heads[st], dim_head, window, halo), dps.next())
            x = x + _dp(ffn(layer_norm(x, sd[lp + "norm_ff.weight"], sd[lp + "norm_ff.bias"], 1e-6), sd, lp + "ff."),
                        dps.next())
    x = layer_norm(x, sd["final_linear.0.weight"], sd["final_linear.0.bias"], 1e-5)
    x = linear(x, sd["final_linear.1.weight"], sd["final_linear.1.bias"])
    x = silu(layer_norm(x, sd["final_linear.2.weight"], sd["final_linear.2.bias"], 1e-5))
    x = x.mean((1, 2))
    return linear(x, sd["classifier.2.weight"], sd["classifier.2.bias"])


# ------------------------------------------------------------------------------------------ Twins
def twins_scrambled_image(x):
    """twins.py:70 as an explicit index map: the [B,H,W,C] map is copied as [B,W,H,C] and the flat
    buffer is then read as [B,C,H,W]:  img[b,c',y',x'] = flat[(c'*H + y')*W + x'],  flat[(w*H + h)*C + c] = x[b,h,w,c]."""
    B, H, W, C = x.shape
    pos = torch.arange(C * H * W, device=x.device)  # flat position of img[c', y', x'] in row-major (C,H,W)
    c = pos % C
    h = (pos // C) % H
    w = pos // (C * H)
    return x[:, h, w, c].view(B, C, H, W)


def twins_global_attention(x, sd, pre, heads, R):
    """twins.py:58-93: q from every token; K/V from a k=s=R conv over the scrambled image (no LayerNorm)."""
    B, H, W, C = x.shape
    q = rearrange(linear(x, sd[pre + "linear_q.weight"]), "b h w (n d) -> b n (h w) d", n=heads)
    if R > 1:
        img = twins_scrambled_image(x)
        a = rearrange(img, "b c (h py) (w px) -> b (h w) c py px", py=R, px=R)
        kvin = torch.einsum("bncyx,ocyx->bno", a, sd[pre + "reduce_conv.weight"]) + sd[pre + "reduce_conv.bias"]
    else:
        kvin = x.reshape(B, H * W, C)
    k, v = rearrange(linear(kvin, sd[pre + "linear_kv.weight"]), "b n (s h d) -> s b h n d", s=2, h=heads)
    o = rearrange(softmax_attention(q, k, v), "b n (h w) d -> b h w (n d)", h=H)
    return linear(o, sd[pre + "linear.weight"], sd[pre + "linear.bias"])


def twins_local_attention(x, sd, pre, heads, dh, W):
    """twins.py:109-152: plain window attention (no shift, bias or mask)."""
    B, Hs, Ws, _ = x.shape
    qkv = linear(x, sd[pre + "weight.weight"], sd[pre + "weight.bias"])
    q, k, v = rearrange(qkv, "b (wy ty) (wx tx) (s h d) -> s b (wy wx) h (ty tx) d", ty=W, tx=W, s=3, h=heads)
    o = softmax_attention(q, k, v)
    o = rearrange(o, "b (wy wx) h (ty tx) d -> b (wy ty) (wx tx) (h d)", wy=Hs // W, ty=W)
    return linear(o, sd[pre + "linear.weight"], sd[pre + "linear.bias"])


def peg(x, w):
    """twins.py:31-36: depthwise 3x3 conv (zero padding 1, no bias) + identity, NHWC."""
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))
    B, H, W, C = x.shape
    out = x
    for ky in range(3):
        for kx in range(3):
            out = out + xp[:, ky:ky + H, kx:kx + W, :] * w[:, 0, ky, kx]
    return out


def twins_forward(sd, img, *, depths, n_heads, dim_head, window, dp_scales=None):
    """TwinsSVT.forward (twins.py:347-356)."""
    dps = DropPathScales(dp_scales)
    x = img.permute(0, 2, 3, 1)
    for st in range(4):
        pre = f"block{st + 1}."
        x = patchify(x, (4, 2, 2, 2)[st])
        x = linear(x, sd[pre + "0.linear.weight"], sd[pre + "0.linear.bias"])
        x = layer_norm(x, sd[pre + "0.norm.weight"], sd[pre + "0.norm.bias"], 1e-5)
        idx = 1
        for i in range(depths[st]):
            lp = f"{pre}{idx}."
            ln = lambda t, nm: layer_norm(t, sd[lp + nm + ".weight"], sd[lp + nm + ".bias"], 1e-6)  # noqa: E731
            x = x + _dp(twins_local_attention(ln(x, "norm_attn_local"), sd, lp + "attn_local.", n_heads[st], dim_head,
                                              window), dps.next())
            x = x + _dp(ffn(ln(x, "norm_ff_local"), sd, lp + "ff_local."), dps.next())
            x = x + _dp(twins_global_attention(ln(x, "norm_attn_global"), sd, lp + "attn_global.", n_heads[st], window),
                        dps.next())
            x = x + _dp(ffn(ln(x, "norm_ff_global"), sd, lp + "ff_global."), dps.next())
            idx += 1
            if i == 0:
                x = peg(x, sd[f"{pre}{idx}.proj.weight"])
                idx += 1
    x = layer_norm(x, sd["final_linear.0.weight"], sd["final_linear.0.bias"], 1e-5)
    x = x.mean((1, 2))
    return linear(x, sd["classifier.2.weight"], sd["classifier.2.bias"])


# ------------------------------------------------------------------------------------------ helpers
def randomize_(module, seed):
    """Seeded re-randomisation of every parameter so zero-initialised tables (rel_pos) and unit LayerNorm
    affines are exercised (SURVEY §8c pins): LN gamma ~ 1+0.1N, beta ~ 0.1N; rel_pos ~ N(0,0.5);
    everything else ~ N(0, 0.02) except biases ~ N(0, 0.02) too."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if "norm" in name or name.startswith("final_linear.0") or name.startswith("final_linear.2"):
                if name.endswith("weight"):
                    p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif "rel_pos" in name:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            elif name.endswith("proj.weight") and p.dim() == 4 and p.shape[1] == 1:  # PEG depthwise kernel
                p.copy_(0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return module
