"""Twins-SVT — drop-in for the reference's models/twins.py (not exported by the reference's models/__init__,
imported as `models.twins.TwinsSVT`).  Each layer = locally-grouped attention (plain 7x7 windows, no bias /
mask / shift) + FFN + global sub-sampled attention (K/V from a k=s=window conv of a scrambled view, twins.py:70)
+ FFN; a depthwise-conv positional encoding generator follows the first layer of every stage."""
from typing import Tuple

from pydantic import StrictFloat, StrictInt
from torch import nn

from ._compat import config_model
from .layer import (DropPath, PositionwiseFeedForward, assign_drop_path, check_no_dropout, ffn_branch,
                    init_transformer_weights, make_classifier, ramp_rates, transformer_layers)

LayerNorm = lambda x: nn.LayerNorm(x, eps=1e-6)  # noqa: E731  (twins.py:12)


class PositionalEncodingGenerator(nn.Module):
    """twins.py:25-36."""

    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, 3, padding=1, bias=False, groups=dim)

    def forward(self, input):
        from vtb200.blocks import PEGFn

        return PEGFn.apply(input, self.proj.weight)


class MultiHeadedAttention(nn.Module):
    """Global sub-sampled attention parameter holder (twins.py:39-93)."""

    def __init__(self, dim, n_head, reduction=1, dropout=0):
        super().__init__()
        self.dim_head = dim // n_head
        self.n_head = n_head
        self.linear_q = nn.Linear(dim, dim, bias=False)
        self.linear_kv = nn.Linear(dim, dim * 2, bias=False)
        self.linear = nn.Linear(dim, dim)
        self.dropout = dropout
        self.reduction = reduction
        if self.reduction > 1:
            self.reduce_conv = nn.Conv2d(dim, dim, self.reduction, stride=self.reduction)


class MultiHeadedLocalAttention(nn.Module):
    """Locally-grouped (window) attention parameter holder (twins.py:96-152)."""

    def __init__(self, dim, n_head, dim_head, window_size, dropout=0):
        super().__init__()
        self.dim_head = dim_head
        self.n_head = n_head
        self.weight = nn.Linear(dim, n_head * dim_head * 3, bias=True)
        self.linear = nn.Linear(n_head * dim_head, dim)
        self.window_size = window_size
        self.dropout = dropout


class TransformerLayer(nn.Module):
    """twins.py:155-197: LSA, FFN, GSA, FFN — one shared DropPath module, four draws per layer."""

    def __init__(self, dim, n_head, dim_head, dim_ff, window_size, activation=nn.SiLU, drop_ff=0, drop_attn=0,
                 drop_path=0):
        super().__init__()
        self.norm_attn_local = LayerNorm(dim)
        self.attn_local = MultiHeadedLocalAttention(dim, n_head, dim_head, window_size, drop_attn)
        self.norm_ff_local = LayerNorm(dim)
        self.ff_local = PositionwiseFeedForward(dim, dim_ff, activation=activation, dropout=drop_ff)
        self.norm_attn_global = LayerNorm(dim)
        self.attn_global = MultiHeadedAttention(dim, n_head, window_size, drop_attn)
        self.norm_ff_global = LayerNorm(dim)
        self.ff_global = PositionwiseFeedForward(dim, dim_ff, activation=activation, dropout=drop_ff)
        self.drop_path = DropPath(drop_path)

    def set_drop_path(self, p):
        self.drop_path.p = p

    def forward(self, input):
        from vtb200 import lib as _l
        from vtb200.blocks import AttnBranchFn, SRABranchFn

        la, ga = self.attn_local, self.attn_global
        check_no_dropout(self, la.dropout, ga.dropout)
        B, H, W, _ = input.shape
        w = la.window_size
        geom = dict(mode=_l.ATTN_WINDOW, batch=B, heads=la.n_head, dh=la.dim_head, nq=w * w, nkv=w * w, Hs=H, Ws=W,
                    window=w, shift=0, halo=0)
        n = self.norm_attn_local
        out = AttnBranchFn.apply(input, self.drop_path.scale(B), n.eps, H * W, geom, None, None, n.weight, n.bias,
                                 la.weight.weight, la.weight.bias, la.linear.weight, la.linear.bias, None)
        out = ffn_branch(out, self.drop_path, self.norm_ff_local, self.ff_local, H * W)
        n = self.norm_attn_global
        red = ga.reduction > 1
        cfg = dict(heads=ga.n_head, reduction=ga.reduction, height=H, width=W, kv_norm=False, scramble=True)
        out = SRABranchFn.apply(out, self.drop_path.scale(B), n.eps, H * W, cfg, n.weight, n.bias,
                                ga.linear_q.weight, ga.linear_kv.weight, ga.linear.weight, ga.linear.bias,
                                ga.reduce_conv.weight if red else None, ga.reduce_conv.bias if red else None,
                                None, None)
        return ffn_branch(out, self.drop_path, self.norm_ff_global, self.ff_global, H * W)


class PatchEmbedding(nn.Module):
    """patchify(s) -> Linear -> LayerNorm(1e-5) (twins.py:200-213)."""

    def __init__(self, in_dim, out_dim, window_size):
        super().__init__()
        self.window_size = window_size
        self.linear = nn.Linear(in_dim * window_size * window_size, out_dim)
        self.norm = nn.LayerNorm(out_dim)

    def forward(self, input):
        from vtb200.blocks import PatchLinearFn

        nchw = input.dim() == 4 and not input.is_contiguous() and input.permute(0, 3, 1, 2).is_contiguous()
        src = input.permute(0, 3, 1, 2) if nchw else input
        return PatchLinearFn.apply(src, self.window_size, nchw, self.norm.eps, self.linear.weight,
                                   self.linear.bias, self.norm.weight, self.norm.bias)


@config_model(name="twins_svt", namespace="model", use_type=True)
class TwinsSVT(nn.Module):
    """twins.py:220-356."""

    def __init__(
        self,
        n_class: StrictInt,
        depths: Tuple[StrictInt, StrictInt, StrictInt, StrictInt],
        dims: Tuple[StrictInt, StrictInt, StrictInt, StrictInt],
        dim_head: StrictInt,
        n_heads: Tuple[StrictInt, StrictInt, StrictInt, StrictInt],
        dim_ffs: Tuple[StrictInt, StrictInt, StrictInt, StrictInt],
        window_size: StrictInt,
        drop_ff: StrictFloat = 0.0,
        drop_attn: StrictFloat = 0.0,
        drop_path: StrictFloat = 0.0,
    ):
        super().__init__()
        self.depths = depths
        in_dims = (3, dims[0], dims[1], dims[2])
        for i, red in enumerate((4, 2, 2, 2)):
            setattr(self, f"block{i + 1}", self.make_block(depths[i], in_dims[i], dims[i], n_heads[i], dim_head,
                                                           dim_ffs[i], window_size, red, drop_ff, drop_attn))
        self.final_linear = nn.Sequential(nn.LayerNorm(dims[-1]))
        self.classifier = make_classifier(dims[-1], n_class)
        self.apply(self.init_weights)
        self.set_dropout(None, drop_path)

    def blocks(self):
        return (self.block1, self.block2, self.block3, self.block4)

    def set_dropout(self, dropout, drop_path):
        """Linear ramp over transformer layers in stage order; patch embeds / PEGs are skipped (twins.py:267-304)."""
        layers = transformer_layers(self.blocks())
        assign_drop_path(layers, ramp_rates(drop_path, sum(self.depths)))

    init_weights = staticmethod(init_transformer_weights)

    def make_block(self, depth, in_dim, dim, n_head, dim_head, dim_ff, window_size, reduction, drop_ff, drop_attn):
        block = [PatchEmbedding(in_dim, dim, reduction)]
        for i in range(depth):
            block.append(TransformerLayer(dim, n_head, dim_head, dim_ff, window_size, drop_ff=drop_ff,
                                          drop_attn=drop_attn))
            if i == 0:
                block.append(PositionalEncodingGenerator(dim))
        return nn.Sequential(*block)

    def forward(self, input):
        from vtb200.blocks import LayerNormFn, LinearFn, MeanRowsFn

        out = input.permute(0, 2, 3, 1)
        for stage in self.blocks():
            out = stage(out)
        B, H, W, C = out.shape
        norm = self.final_linear[0]
        out = LayerNormFn.apply(out, norm.weight, norm.bias, norm.eps)
        out = MeanRowsFn.apply(out.view(B, H * W, C))
        lin = self.classifier[2]
        return LinearFn.apply(out, lin.weight, lin.bias)
