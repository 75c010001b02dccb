"""Pyramid Vision Transformer — drop-in for the reference's models/pvt.py (lines cited per class)."""
import torch
from torch import nn

from .layer import DropPath, PositionwiseFeedForward, check_no_dropout, ffn_branch, tuple2

LayerNorm = lambda x: nn.LayerNorm(x, eps=1e-6)  # noqa: E731  (pvt.py:9)


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
        check_no_dropout(self, a.dropout, self.ff[2].p)
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
        from vtb200.blocks import PVTPatchEmbedFn

        check_no_dropout(self, self.dropout.p)
        height, width = input.shape[2] // self.patch, input.shape[3] // self.patch
        out = PVTPatchEmbedFn.apply(input, self.patch, self.norm.eps, self.conv.weight, self.conv.bias,
                                    self.norm.weight, self.norm.bias, self.pos, self.cls_token)
        return out, (height, width)


class PyramidVisionTransformer(nn.Module):
    """pvt.py:146-280."""

    def __init__(self, image_size, n_class, in_dim, depths, patch_embed_dims, n_heads, dim_ffs, reductions,
                 drop_ff=0, drop_attn=0, drop_path=0):
        super().__init__()
        self.depths = depths
        self.patch_embedding = nn.ModuleList()
        patch_embed_dims = list(patch_embed_dims)
        patch_sizes = (4, 2, 2, 2)
        img_size = tuple2(image_size)
        in_dims = [in_dim] + patch_embed_dims[:-1]
        for i, (p_in, p_out, p_size) in enumerate(zip(in_dims, patch_embed_dims, patch_sizes)):
            last = i == len(patch_embed_dims) - 1
            self.patch_embedding.append(
                PatchEmbedding(img_size, p_in, p_out, p_size, cls_token=last, dropout=drop_ff))
            img_size = (img_size[0] // p_size, img_size[1] // p_size)
        for i in range(4):
            setattr(self, f"block{i + 1}", self.make_block(depths[i], patch_embed_dims[i], n_heads[i],
                                                           dim_ffs[i], reductions[i], drop_ff, drop_attn))
        self.norm = LayerNorm(patch_embed_dims[-1])
        self.classifier = nn.Linear(patch_embed_dims[-1], n_class)
        self.apply(self.init_weights)
        self.set_drop_path(drop_path)

    def blocks(self):
        return (self.block1, self.block2, self.block3, self.block4)

    def set_drop_path(self, drop_path):
        p = torch.linspace(0, drop_path, sum(self.depths)).tolist()
        i = 0
        for stage in self.blocks():
            for layer in stage:
                layer.set_drop_path(p[i])
                i += 1

    def init_weights(self, module):
        if isinstance(module, nn.Linear):
            nn.init.normal_(module.weight, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.LayerNorm):
            nn.init.ones_(module.weight)
            nn.init.zeros_(module.bias)

    def make_block(self, depth, dim, n_head, dim_ff, reduction, drop_ff, drop_attn):
        return nn.ModuleList(
            [TransformerLayer(dim, n_head, dim_ff, reduction=reduction, drop_ff=drop_ff, drop_attn=drop_attn)
             for _ in range(depth)])

    def forward(self, input):
        from vtb200.blocks import LayerNormFn, LinearFn

        batch = input.shape[0]
        out = input
        for i, stage in enumerate(self.blocks()):
            out, (height, width) = self.patch_embedding[i](out)
            for layer in stage:
                out = layer(out, height, width)
            if i < 3:
                # tokens -> NCHW view for the next stage's conv (pvt.py:261); stays a view, the patch
                # gather reads the NHWC memory directly
                out = out.reshape(batch, height, width, -1).permute(0, 3, 1, 2)
        out = LayerNormFn.apply(out[:, 0], self.norm.weight, self.norm.bias, self.norm.eps)
        return LinearFn.apply(out, self.classifier.weight, self.classifier.bias)
