"""HaloNet-style transformer — drop-in for the reference's models/halo_transformer.py.

Blocked local attention with a zero-padded halo (halo_transformer.py:57-114): the F.unfold gather, the
relative-position bias and the softmax run inside vtb_attention_fwd/bwd (HALO mode); padded key slots
keep their softmax mass (logit = bias) exactly like the reference.  The reference TransformerLayer adds
its branches in place (`input += ...`, :147-148), which makes ITS autograd raise; the maths is the
ordinary out-of-place residual, which is what runs here.
"""
import torch
from torch import nn

from .layer import (DropPath, PositionwiseFeedForward, assign_drop_path, check_no_dropout, ffn_branch,
                    init_transformer_weights, make_classifier, ramp_rates, transformer_layers)


def halo_pos(window, halo):
    """pos[t, j] = (ky - (ty+halo) + window+halo-1) * K + (kx - (tx+halo) + window+halo-1), K = window+2*halo
    (halo_transformer.py:41-55).  Returns (pos int64 [W^2, K^2], max_pos)."""
    K = window + 2 * halo
    k = torch.arange(K)
    q = torch.arange(window) + halo
    off = window + halo - 1
    dy = (k[None, :] - q[:, None]) + off  # [W, K]
    pos = (dy[:, None, :, None] * K + dy[None, :, None, :]).reshape(window * window, K * K)
    max_pos = off * 2 * K + off * 2
    return pos.contiguous(), max_pos


class MultiHeadedHaloAttention(nn.Module):
    """Parameter / buffer holder (halo_transformer.py:22-55)."""

    def __init__(self, dim, n_head, dim_head, window_size, halo_size, dropout=0):
        super().__init__()
        self.dim_head = dim_head
        self.n_head = n_head
        self.weight = nn.Linear(dim, n_head * dim_head * 3, bias=False)
        self.linear = nn.Linear(n_head * dim_head, dim)
        self.window_size = window_size
        self.halo_size = halo_size
        self.dropout = dropout
        rel_pos, max_pos = halo_pos(window_size, halo_size)
        self.register_buffer("pos", rel_pos)
        self.rel_pos = nn.Embedding(max_pos + 1, n_head)
        self.rel_pos.weight.detach().zero_()
        self._tab = None

    def tables(self):
        key = (self.pos.device, self.pos._version, self.pos.data_ptr())
        if self._tab is None or self._tab[0] != key:
            self._tab = (key, self.pos.to(torch.int32).contiguous())
        return self._tab[1]


class TransformerLayer(nn.Module):
    """halo_transformer.py:117-150."""

    def __init__(self, dim, n_head, dim_head, dim_ff, window_size, halo_size, activation=nn.SiLU, drop_ff=0,
                 drop_attn=0, drop_path=0):
        super().__init__()
        self.norm_attn = nn.LayerNorm(dim, eps=1e-6)
        self.attn = MultiHeadedHaloAttention(dim, n_head, dim_head, window_size, halo_size, drop_attn)
        self.drop_path = DropPath(drop_path)
        self.norm_ff = nn.LayerNorm(dim, eps=1e-6)
        self.ff = PositionwiseFeedForward(dim, dim_ff, activation=activation, dropout=drop_ff)

    def set_drop_path(self, p):
        self.drop_path.p = p

    def forward(self, input):
        from vtb200 import lib as _l
        from vtb200.blocks import AttnBranchFn

        a = self.attn
        check_no_dropout(self, a.dropout)
        B, H, W, _ = input.shape
        w, hl = a.window_size, a.halo_size
        geom = dict(mode=_l.ATTN_HALO, batch=B, heads=a.n_head, dh=a.dim_head, nq=w * w,
                    nkv=(w + 2 * hl) ** 2, Hs=H, Ws=W, window=w, shift=0, halo=hl)
        out = AttnBranchFn.apply(input, self.drop_path.scale(B), self.norm_attn.eps, H * W, geom, a.tables(),
                                 None, self.norm_attn.weight, self.norm_attn.bias, a.weight.weight, None,
                                 a.linear.weight, a.linear.bias, a.rel_pos.weight)
        return ffn_branch(out, self.drop_path, self.norm_ff, self.ff, H * W)


class PatchEmbedding(nn.Module):
    """patchify(s) -> Linear -> LayerNorm(1e-5) (halo_transformer.py:153-166)."""

    def __init__(self, in_dim, out_dim, window_size):
        super().__init__()
        self.window_size = window_size
        self.linear = nn.Linear(in_dim * window_size * window_size, out_dim)
        self.norm = nn.LayerNorm(out_dim)

    def forward(self, input):
        from vtb200.blocks import PatchLinearFn

        # the first stage receives the NCHW image as an NHWC permuted view (halo_transformer.py:272)
        nchw = input.dim() == 4 and not input.is_contiguous() and input.permute(0, 3, 1, 2).is_contiguous()
        src = input.permute(0, 3, 1, 2) if nchw else input
        return PatchLinearFn.apply(src, self.window_size, nchw, self.norm.eps, self.linear.weight,
                                   self.linear.bias, self.norm.weight, self.norm.bias)


def reduce_size(size, reduction):
    return (size[0] // reduction, size[1] // reduction)


class HaloTransformer(nn.Module):
    """halo_transformer.py:173-280."""

    def __init__(self, image_size, n_class, depths, dims, dim_head, n_heads, dim_ffs, window_size, halo_size,
                 drop_ff=0, drop_attn=0, drop_path=0):
        super().__init__()
        self.depths = depths
        in_dims = (3, dims[0], dims[1], dims[2])
        reductions = (4, 2, 2, 2)
        for i in range(4):
            setattr(self, f"block{i + 1}", self.make_block(
                depths[i], in_dims[i], dims[i], n_heads[i], dim_head, dim_ffs[i], window_size, halo_size,
                reductions[i], drop_ff, drop_attn, drop_path))
        self.final_linear = nn.Sequential(
            nn.LayerNorm(dims[-1]),
            nn.Linear(dims[-1], dims[-1] * 2),
            nn.LayerNorm(dims[-1] * 2),
            nn.SiLU(inplace=True),
        )
        self.classifier = make_classifier(dims[-1] * 2, n_class, std=0.01)
        self.apply(self.init_weights)

    init_weights = staticmethod(init_transformer_weights)

    def make_block(self, depth, in_dim, dim, n_head, dim_head, dim_ff, window_size, halo_size, reduction,
                   drop_ff, drop_attn, drop_path):
        block = [PatchEmbedding(in_dim, dim, reduction)]
        for _ in range(depth):
            block.append(TransformerLayer(dim, n_head, dim_head, dim_ff, window_size, halo_size,
                                          drop_ff=drop_ff, drop_attn=drop_attn, drop_path=drop_path))
        return nn.Sequential(*block)

    def forward(self, input):
        from vtb200.blocks import LayerNormFn, LinearFn, MeanRowsFn, SiLUFn

        out = self.block1(input.permute(0, 2, 3, 1))
        out = self.block2(out)
        out = self.block3(out)
        out = self.block4(out)
        B, H, W, C = out.shape
        n0, lin, n1, _ = self.final_linear
        out = LayerNormFn.apply(out, n0.weight, n0.bias, n0.eps)
        out = LinearFn.apply(out, lin.weight, lin.bias)
        out = LayerNormFn.apply(out, n1.weight, n1.bias, n1.eps)
        out = SiLUFn.apply(out)
        out = MeanRowsFn.apply(out.view(B, H * W, -1))
        head = self.classifier[2]
        return LinearFn.apply(out, head.weight, head.bias)
