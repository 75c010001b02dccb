"""ViT — drop-in for the reference's models/vit.py: same classes, constructor signatures, parameter names and
state_dict (reference lines cited per class); forward() runs on libvtb200 kernels only.  DINOHead / dino live in
models/dino.py and are re-exported here under their reference names."""
import itertools
import math

import torch
from torch import nn
from torch.nn import functional as F

from .dino import DINOHead, dino  # noqa: F401  (vit.py:206-307 of the reference)
from .layer import (DropPath, PositionwiseFeedForward, assign_drop_path, autocast_dtype, check_no_dropout, ffn_branch,
                    init_transformer_weights, linspace_rates, tuple2)

LN_EPS = 1e-6  # vit.py:13


def LayerNorm(dim):
    return nn.LayerNorm(dim, eps=LN_EPS)


class MultiHeadedAttention(nn.Module):
    """Parameter holder of global MHSA (vit.py:16-45): `qkv` rows ordered (q|k|v) x head x dh, then `linear`."""

    def __init__(self, dim, n_head, bias=True, dropout=0):
        super().__init__()
        self.n_head, self.dim_head = n_head, dim // n_head
        self.qkv = nn.Linear(dim, 3 * dim, bias=bias)
        self.dropout = nn.Dropout(dropout)
        self.linear = nn.Linear(dim, dim)


class TransformerLayer(nn.Module):
    """vit.py:48-66.  One DropPath module shared by both branches (two RNG draws per layer, attention first)."""

    def __init__(self, dim, n_head, dim_ff, dropout, drop_attn, drop_ff, drop_path):
        super().__init__()
        self.norm_attn = LayerNorm(dim)
        self.attn = MultiHeadedAttention(dim, n_head, dropout=drop_attn)
        self.norm_ff = LayerNorm(dim)
        self.ff = PositionwiseFeedForward(dim, dim_ff, dropout=drop_ff)
        self.dropout = nn.Dropout(dropout)
        self.drop_path = DropPath(drop_path)

    def set_drop_path(self, p):
        self.drop_path.p = p

    def forward(self, input):
        from vtb200 import lib as _l
        from vtb200.blocks import AttnBranchFn, make_dropout_keep

        check_no_dropout(self, self.attn.dropout.p)
        batch, tokens = input.shape[0], input.shape[1]
        att, ln = self.attn, self.norm_attn
        geom = dict(mode=_l.ATTN_GLOBAL, batch=batch, heads=att.n_head, dh=att.dim_head, nq=tokens, nkv=tokens)
        # vit.py:60: drop_path(dropout(attn(...))) — the Dropout mask is drawn before the DropPath mask
        out_drop = make_dropout_keep(self.dropout.training, self.dropout.p, input.shape, autocast_dtype(input), input.device)
        hidden = AttnBranchFn.apply(input, self.drop_path.scale(batch), ln.eps, tokens, geom, None, None, ln.weight,
                                    ln.bias, att.qkv.weight, att.qkv.bias, att.linear.weight, att.linear.bias, None,
                                    out_drop)
        return ffn_branch(hidden, self.drop_path, self.norm_ff, self.ff, tokens, out_dropout=self.dropout)


class PatchEmbedding(nn.Module):
    """Holder of the k = s = patch conv (vit.py:69-76); executed as patch gather + tcgen05 GEMM by the model."""

    def __init__(self, in_dim, out_dim, window_size):
        super().__init__()
        self.window_size = window_size
        self.linear = nn.Conv2d(in_dim, out_dim, window_size, stride=window_size)


class VisionTransformer(nn.Module):
    """vit.py:79-203."""

    def __init__(self, head, image_size, window_size, depth, dim, n_head, dim_ff, dropout, drop_attn, drop_ff,
                 drop_path):
        super().__init__()
        height, width = tuple2(image_size)
        n_patch = (height // window_size) * (width // window_size)
        self.patch_embedding = PatchEmbedding(3, dim, window_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patch + 1, dim))
        self.pos_drop = nn.Dropout(dropout)
        self.layers = nn.ModuleList(TransformerLayer(dim, n_head, dim_ff, dropout, drop_attn, drop_ff, rate)
                                    for rate in linspace_rates(drop_path, depth))
        self.norm = LayerNorm(dim)
        self.apply(self.init_weights)
        for table in (self.pos_embed, self.cls_token):
            nn.init.normal_(table, std=0.02)
        self.head = head  # attached after apply(): a head keeps its own initialisation
        self.depth = depth

    init_weights = staticmethod(init_transformer_weights)

    def set_drop_path(self, drop_path):
        assign_drop_path(self.layers, linspace_rates(drop_path, self.depth))

    def interpolate_pos_embedding(self, n_patch, pos_embed):
        """Positional table for an input with `n_patch` patches: identity at the training resolution, bicubic
        resample of the grid part otherwise (vit.py:153-175).  Left to ATen — one [1, 196, D] op per step."""
        n_pos = pos_embed.shape[1] - 1
        if n_patch == n_pos:
            return pos_embed
        side, dim = int(math.sqrt(n_pos)), pos_embed.shape[-1]
        grid = pos_embed[:, 1:].reshape(1, side, side, dim).permute(0, 3, 1, 2)
        grid = F.interpolate(grid, scale_factor=math.sqrt(n_patch / n_pos), mode="bicubic", align_corners=False,
                             recompute_scale_factor=False)
        return torch.cat((pos_embed[:, :1], grid.permute(0, 2, 3, 1).reshape(1, -1, dim)), 1)

    def forward_feature(self, input):
        from vtb200.blocks import LayerNormFn, ViTPatchEmbedFn, dropout

        patch = self.patch_embedding.window_size
        n_patch = (input.shape[-2] // patch) * (input.shape[-1] // patch)
        conv = self.patch_embedding.linear
        tokens = ViTPatchEmbedFn.apply(input, conv.weight, conv.bias, self.cls_token,
                                       self.interpolate_pos_embedding(n_patch, self.pos_embed), patch)
        tokens = dropout(tokens, self.pos_drop)  # vit.py:146
        for layer in self.layers:
            tokens = layer(tokens)
        # the final LayerNorm is row-wise, so only the cls rows that are returned need it (vit.py:149-151)
        return LayerNormFn.apply(tokens[:, 0], self.norm.weight, self.norm.bias, self.norm.eps)

    def forward(self, input):
        """A tensor, or a list of crops: consecutive crops of equal width are batched together (vit.py:177-198)."""
        crops = list(input) if isinstance(input, (list, tuple)) else [input]
        feats = [self.forward_feature(torch.cat(list(group)))
                 for _, group in itertools.groupby(crops, key=lambda c: c.shape[-1])]
        output = feats[0] if len(feats) == 1 else torch.cat(feats)
        return output if self.head is None else self.head(output)


class FusedLinear(nn.Linear):
    """nn.Linear whose forward/backward run on the tcgen05 GEMM (a drop-in `head` for VisionTransformer)."""

    def forward(self, input):
        from vtb200.blocks import LinearFn

        return LinearFn.apply(input, self.weight, self.bias)
