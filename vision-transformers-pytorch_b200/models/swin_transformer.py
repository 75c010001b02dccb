"""Swin Transformer — drop-in for the reference's models/swin_transformer.py (class names, ctor
signatures/annotations, parameter + buffer names and shapes are the reference's; lines cited per class).
The roll, window partition, relative-position bias gather, shift mask, softmax and un-partition of
swin_transformer.py:103-160 all happen inside one attention kernel (vtb_attention_fwd/bwd, WINDOW mode).
"""
import math
from typing import Tuple

import torch
from pydantic import StrictFloat, StrictInt
from torch import nn

from ._compat import config_model
from .layer import (DropPath, PositionwiseFeedForward, assign_drop_path, check_no_dropout, ffn_branch,
                    init_transformer_weights, make_classifier, ramp_rates, transformer_layers)

LayerNorm = lambda x: nn.LayerNorm(x, eps=1e-6)  # noqa: E731  (swin_transformer.py:12)


def window_tables(input_size, window_size, shift):
    """Integer tables of swin_transformer.py:50-101, derived independently:

    pos[t, u]  = (ky - qy + W-1) * (2W-1) + (kx - qx + W-1) on the coordinates of window 0 of the
                 (rolled, if shift) grid, with offsets zeroed where the pair is masked;
    mask[w, t, u] = True where the two tokens' original coordinates differ by >= W in y or x.
    Returns (pos int64 [W^2, W^2], mask bool [nW, W^2, W^2] or None).
    """
    W = window_size
    nh, nw = input_size[0] // W, input_size[1] // W
    ys = torch.arange(nh * W)
    xs = torch.arange(nw * W)
    if shift:
        s = W // 2
        ys = (ys + s) % (nh * W)  # torch.roll(a, -s)[i] = a[(i + s) % n]
        xs = (xs + s) % (nw * W)
    # coordinates of every window's tokens: [nW, W*W]
    wy = ys.view(nh, W)[:, None, :, None].expand(nh, nw, W, W).reshape(nh * nw, W * W)
    wx = xs.view(nw, W)[None, :, None, :].expand(nh, nw, W, W).reshape(nh * nw, W * W)
    dy = wy[:, None, :] - wy[:, :, None]  # [nW, q, k] = k - q
    dx = wx[:, None, :] - wx[:, :, None]
    if shift:
        ok = (dy.abs() < W) & (dx.abs() < W)
        dy, dx = dy * ok, dx * ok
        mask = ~ok
    else:
        mask = None
    pos = (dy + (W - 1)) * (2 * W - 1) + (dx + (W - 1))
    return pos[0].contiguous(), mask


class MultiHeadedLocalAttention(nn.Module):
    """Parameter / buffer holder (swin_transformer.py:25-101)."""

    def __init__(self, dim, n_head, dim_head, input_size, window_size, shift, dropout=0):
        super().__init__()
        self.dim_head = dim_head
        self.n_head = n_head
        self.weight = nn.Linear(dim, n_head * dim_head * 3, bias=True)
        self.linear = nn.Linear(n_head * dim_head, dim)
        self.input_size = input_size
        self.window_size = window_size
        self.dropout = dropout
        self.shift = shift
        pos, mask = window_tables(input_size, window_size, shift)
        self.register_buffer("pos", pos)
        self.rel_pos = nn.Embedding((2 * window_size - 1) ** 2, n_head)
        self.rel_pos.weight.detach().zero_()
        if shift:
            self.register_buffer("local_mask", mask)
        self._tab = None

    def tables(self):
        """int32 / uint8 device copies of the buffers for the kernel (buffers stay int64/bool for
        state_dict compatibility); rebuilt if the buffers moved or were reloaded."""
        key = (self.pos.device, self.pos._version, self.pos.data_ptr())
        if self._tab is None or self._tab[0] != key:
            pos32 = self.pos.to(torch.int32).contiguous()
            mask8 = None
            if self.shift:
                # rows padded to a 64-byte pitch when they fit: the window kernels fetch them with 16-byte loads
                m = self.local_mask.to(torch.uint8)
                pitch = 64 if m.shape[2] <= 64 else m.shape[2]
                mask8 = torch.zeros((m.shape[0], m.shape[1], pitch), dtype=torch.uint8, device=m.device)
                mask8[:, :, :m.shape[2]] = m
            self._tab = (key, pos32, mask8)
        return self._tab[1], self._tab[2]


class TransformerLayer(nn.Module):
    """swin_transformer.py:163-197."""

    def __init__(self, dim, n_head, dim_head, dim_ff, input_size, window_size, shift, activation=nn.SiLU,
                 drop_ff=0, drop_attn=0, drop_path=0):
        super().__init__()
        self.norm_attn = LayerNorm(dim)
        self.attn = MultiHeadedLocalAttention(dim, n_head, dim_head, input_size, window_size, shift, drop_attn)
        self.drop_path = DropPath(drop_path)
        self.norm_ff = LayerNorm(dim)
        self.ff = PositionwiseFeedForward(dim, dim_ff, activation=activation, dropout=drop_ff)

    def set_drop_path(self, p):
        self.drop_path.p = p

    def forward(self, input):
        from vtb200 import lib as _l
        from vtb200.blocks import AttnBranchFn

        a = self.attn
        check_no_dropout(self, a.dropout)
        B, H, W, _ = input.shape
        w = a.window_size
        pos, mask = a.tables()
        geom = dict(mode=_l.ATTN_WINDOW, batch=B, heads=a.n_head, dh=a.dim_head, nq=w * w, nkv=w * w, Hs=H,
                    Ws=W, window=w, shift=(w // 2) if a.shift else 0, halo=0)
        out = AttnBranchFn.apply(input, self.drop_path.scale(B), self.norm_attn.eps, H * W, geom, pos, mask,
                                 self.norm_attn.weight, self.norm_attn.bias, a.weight.weight, a.weight.bias,
                                 a.linear.weight, a.linear.bias, a.rel_pos.weight)
        return ffn_branch(out, self.drop_path, self.norm_ff, self.ff, H * W)


class PatchEmbedding(nn.Module):
    """patchify(4) -> Linear -> LayerNorm(1e-5) (swin_transformer.py:200-213)."""

    def __init__(self, in_dim, out_dim, window_size):
        super().__init__()
        self.window_size = window_size
        self.linear = nn.Linear(in_dim * window_size * window_size, out_dim)
        self.norm = nn.LayerNorm(out_dim)

    def forward(self, input, nchw=False):
        from vtb200.blocks import PatchLinearFn

        return PatchLinearFn.apply(input, self.window_size, nchw, self.norm.eps, self.linear.weight,
                                   self.linear.bias, self.norm.weight, self.norm.bias)


class PatchMerge(nn.Module):
    """patchify(2) -> LayerNorm(4C, 1e-5) -> Linear(no bias) (swin_transformer.py:216-229)."""

    def __init__(self, in_dim, out_dim, window_size):
        super().__init__()
        self.window_size = window_size
        self.norm = nn.LayerNorm(in_dim * window_size * window_size)
        self.linear = nn.Linear(in_dim * window_size * window_size, out_dim, bias=False)

    def forward(self, input):
        from vtb200.blocks import PatchMergeFn

        return PatchMergeFn.apply(input, self.window_size, self.norm.eps, self.norm.weight, self.norm.bias,
                                  self.linear.weight)


def reduce_size(size, reduction):
    return (size[0] // reduction, size[1] // reduction)


@config_model(name="swin_transformer", namespace="model", use_type=True)
class SwinTransformer(nn.Module):
    """swin_transformer.py:236-379."""

    def __init__(
        self,
        image_size: Tuple[StrictInt, StrictInt],
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
        self.patch_embedding = PatchEmbedding(3, dims[0], 4)
        stage_in = (3, dims[0], dims[1], dims[2])
        stage_red = (1, 2, 2, 2)
        size = reduce_size(image_size, 4)
        for i in range(4):
            size = reduce_size(size, stage_red[i])
            block = self.make_block(depths[i], stage_in[i], dims[i], n_heads[i], dim_head, dim_ffs[i], size,
                                    window_size, stage_red[i], drop_ff, drop_attn)
            setattr(self, f"block{i + 1}", block)
        self.final_linear = nn.Sequential(nn.LayerNorm(dims[-1]))
        self.classifier = make_classifier(dims[-1], n_class)
        self.apply(self.init_weights)
        self.set_dropout(None, drop_path)

    def blocks(self):
        return (self.block1, self.block2, self.block3, self.block4)

    def set_dropout(self, dropout, drop_path):
        """Linear drop-path ramp over the transformer layers in stage order (swin:286-321)."""
        layers = transformer_layers(self.blocks())
        assign_drop_path(layers, ramp_rates(drop_path, sum(self.depths)))

    init_weights = staticmethod(init_transformer_weights)

    def make_block(self, depth, in_dim, dim, n_head, dim_head, dim_ff, input_size, window_size, reduction,
                   drop_ff, drop_attn):
        block = []
        if reduction > 1:
            block.append(PatchMerge(in_dim, dim, reduction))
        for i in range(depth):
            block.append(TransformerLayer(dim, n_head, dim_head, dim_ff, input_size, window_size,
                                          shift=i % 2 == 0, drop_ff=drop_ff, drop_attn=drop_attn))
        return nn.Sequential(*block)

    def forward(self, input):
        from vtb200.blocks import LayerNormFn, LinearFn, MeanRowsFn

        out = self.patch_embedding(input, nchw=True)  # NCHW image -> NHWC tokens (swin:371)
        for stage in self.blocks():
            out = stage(out)
        B, H, W, C = out.shape
        norm = self.final_linear[0]
        out = LayerNormFn.apply(out, norm.weight, norm.bias, norm.eps)
        out = MeanRowsFn.apply(out.view(B, H * W, C))
        lin = self.classifier[2]
        return LinearFn.apply(out, lin.weight, lin.bias)
