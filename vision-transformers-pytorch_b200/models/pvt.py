"""Pyramid Vision Transformer — drop-in for the reference's models/pvt.py (lines cited per class)."""
import torch
from torch import nn

from .layer import (DropPath, PositionwiseFeedForward, assign_drop_path, check_no_dropout, ffn_branch,
                    init_transformer_weights, linspace_rates, transformer_layers, tuple2)

def LayerNorm(dim):  # pvt.py:9
    return nn.LayerNorm(dim, eps=1e-6)


class MultiHeadedAttention(nn.Module):
    """Spatial-reduction attention parameter holder (pvt.py:12-69)."""

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
            self.reduce_norm = LayerNorm(dim)


class TransformerLayer(nn.Module):
    """pvt.py:72-101."""

    def __init__(self, dim, n_head, dim_ff, activation=nn.SiLU, reduction=1, drop_ff=0, drop_attn=0,
                 drop_path=0):
        super().__init__()
        self.norm_attn = LayerNorm(dim)
        self.attn = MultiHeadedAttention(dim, n_head, reduction, drop_attn)
        self.drop_path = DropPath(drop_path)
        self.norm_ff = LayerNorm(dim)
        self.ff = PositionwiseFeedForward(dim, dim_ff, activation=activation, dropout=drop_ff)

    def set_drop_path(self, p):
        self.drop_path.p = p

    def forward(self, input, height, width):
        from vtb200.blocks import SRABranchFn

        a = self.attn
        check_no_dropout(self, a.dropout)
        B, N, _ = input.shape
        cfg = dict(heads=a.n_head, reduction=a.reduction, height=height, width=width)
        red = a.reduction > 1
        out = SRABranchFn.apply(
            input, self.drop_path.scale(B), self.norm_attn.eps, N, cfg, self.norm_attn.weight,
            self.norm_attn.bias, a.linear_q.weight, a.linear_kv.weight, a.linear.weight, a.linear.bias,
            a.reduce_conv.weight if red else None, a.reduce_conv.bias if red else None,
            a.reduce_norm.weight if red else None, a.reduce_norm.bias if red else None)
        return ffn_branch(out, self.drop_path, self.norm_ff, self.ff, N)


class PatchEmbedding(nn.Module):
    """conv k=s=p -> LN -> [cls] -> + pos (pvt.py:104-143)."""

    def __init__(self, image_size, in_dim, dim, patch_size, cls_token=False, dropout=0):
        super().__init__()
        size = tuple2(patch_size)
        img_size = tuple2(image_size)
        self.conv = nn.Conv2d(in_dim, dim, size, stride=size)
        self.norm = LayerNorm(dim)
        height, width = img_size[0] // size[0], img_size[1] // size[1]
        n_patch = height * width
        if cls_token:
            n_patch += 1
        self.pos = nn.Parameter(torch.randn(n_patch, dim) * 0.02)
        self.cls_token = None
        if cls_token:
            self.cls_token = nn.Parameter(torch.randn(dim) * 0.02)
        self.dim = dim
        self.dropout = nn.Dropout(dropout)
        self.patch = size[0]

    def forward(self, input):
        from vtb200.blocks import PVTPatchEmbedFn, dropout

        height, width = input.shape[2] // self.patch, input.shape[3] // self.patch
        out = PVTPatchEmbedFn.apply(input, self.patch, self.norm.eps, self.conv.weight, self.conv.bias,
                                    self.norm.weight, self.norm.bias, self.pos, self.cls_token)
        return dropout(out, self.dropout), (height, width)  # pvt.py:141


class PyramidVisionTransformer(nn.Module):
    """pvt.py:146-280: four stages of (conv patch embedding, SRA transformer layers); cls token at the last stage."""

    PATCH_SIZES = (4, 2, 2, 2)

    def __init__(self, image_size, n_class, in_dim, depths, patch_embed_dims, n_heads, dim_ffs, reductions,
                 drop_ff=0, drop_attn=0, drop_path=0):
        super().__init__()
        self.depths = depths
        dims = list(patch_embed_dims)
        size = tuple2(image_size)
        embeds = []
        for stage, (width_in, width_out, patch) in enumerate(zip([in_dim] + dims[:-1], dims, self.PATCH_SIZES)):
            embeds.append(PatchEmbedding(size, width_in, width_out, patch, cls_token=(stage == len(dims) - 1),
                                         dropout=drop_ff))
            size = (size[0] // patch, size[1] // patch)
        self.patch_embedding = nn.ModuleList(embeds)
        for stage in range(4):
            layers = self.make_block(depths[stage], dims[stage], n_heads[stage], dim_ffs[stage], reductions[stage],
                                     drop_ff, drop_attn)
            setattr(self, f"block{stage + 1}", layers)
        self.norm = LayerNorm(dims[-1])
        self.classifier = nn.Linear(dims[-1], n_class)
        self.apply(self.init_weights)
        self.set_drop_path(drop_path)

    init_weights = staticmethod(init_transformer_weights)

    def blocks(self):
        return (self.block1, self.block2, self.block3, self.block4)

    def set_drop_path(self, drop_path):
        layers = transformer_layers(self.blocks())
        assign_drop_path(layers, linspace_rates(drop_path, len(layers)))

    def make_block(self, depth, dim, n_head, dim_ff, reduction, drop_ff, drop_attn):
        return nn.ModuleList(TransformerLayer(dim, n_head, dim_ff, reduction=reduction, drop_ff=drop_ff,
                                              drop_attn=drop_attn) for _ in range(depth))

    def forward(self, input):
        from vtb200.blocks import LayerNormFn, LinearFn

        batch, out = input.shape[0], input
        for stage, layers in enumerate(self.blocks()):
            out, (height, width) = self.patch_embedding[stage](out)
            for layer in layers:
                out = layer(out, height, width)
            if stage < 3:
                # tokens -> NCHW *view* for the next stage's conv (pvt.py:261): no copy, the patch gather reads
                # the NHWC memory directly
                out = out.reshape(batch, height, width, -1).permute(0, 3, 1, 2)
        cls = LayerNormFn.apply(out[:, 0], self.norm.weight, self.norm.bias, self.norm.eps)
        return LinearFn.apply(cls, self.classifier.weight, self.classifier.bias)
