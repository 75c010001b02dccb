"""ViT / DINO — drop-in for the reference's models/vit.py (same classes, ctor signatures, parameter names
and state_dict; reference lines cited per class).  forward() runs on libvtb200 kernels only."""
import math
from typing import Tuple, Union

import torch
from pydantic import StrictBool, StrictFloat, StrictInt
from torch import nn
from torch.nn import functional as F

from ._compat import config_model
from .layer import DropPath, PositionwiseFeedForward, check_no_dropout, ffn_branch, tuple2

LayerNorm = lambda x: nn.LayerNorm(x, eps=1e-6)  # noqa: E731  (vit.py:13)


class MultiHeadedAttention(nn.Module):
    """Parameter holder for global MHSA (vit.py:16-45): qkv Linear rows ordered (q|k|v) x head x dh."""

    def __init__(self, dim, n_head, bias=True, dropout=0):
        super().__init__()
        self.dim_head = dim // n_head
        self.n_head = n_head
        self.qkv = nn.Linear(dim, dim * 3, bias=bias)
        self.dropout = nn.Dropout(dropout)
        self.linear = nn.Linear(dim, dim)


class TransformerLayer(nn.Module):
    """vit.py:48-66.  One shared DropPath module, called once per branch (two RNG draws, same order)."""

    def __init__(self, dim, n_head, dim_ff, dropout, drop_attn, drop_ff, drop_path):
        super().__init__()
        self.norm_attn = LayerNorm(dim)
        self.attn = MultiHeadedAttention(dim, n_head, dropout=drop_attn)
        self.norm_ff = LayerNorm(dim)
        self.ff = PositionwiseFeedForward(dim, dim_ff, dropout=drop_ff)
        self.dropout = nn.Dropout(dropout)
        self.drop_path = DropPath(drop_path)

    def forward(self, input):
        from vtb200 import lib as _l
        from vtb200.blocks import AttnBranchFn

        check_no_dropout(self, self.dropout.p, self.attn.dropout.p, self.ff[2].p)
        B, N, _ = input.shape
        a = self.attn
        geom = dict(mode=_l.ATTN_GLOBAL, batch=B, heads=a.n_head, dh=a.dim_head, nq=N, nkv=N)
        out = AttnBranchFn.apply(input, self.drop_path.scale(B), self.norm_attn.eps, N, geom, None, None,
                                 self.norm_attn.weight, self.norm_attn.bias, a.qkv.weight, a.qkv.bias,
                                 a.linear.weight, a.linear.bias, None)
        out = ffn_branch(out, self.drop_path, self.norm_ff, self.ff, N)
        return out

    def set_drop_path(self, p):
        self.drop_path.p = p


class PatchEmbedding(nn.Module):
    """Holder of the k=s conv (vit.py:69-76); runs as gather + tcgen05 GEMM in VisionTransformer."""

    def __init__(self, in_dim, out_dim, window_size):
        super().__init__()
        self.window_size = window_size
        self.linear = nn.Conv2d(in_dim, out_dim, window_size, stride=window_size)


class VisionTransformer(nn.Module):
    """vit.py:79-203."""

    def __init__(self, head, image_size, window_size, depth, dim, n_head, dim_ff, dropout, drop_attn,
                 drop_ff, drop_path):
        super().__init__()
        image_size = tuple2(image_size)
        n_patch = (image_size[0] // window_size) * (image_size[1] // window_size)
        self.patch_embedding = PatchEmbedding(3, dim, window_size)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_patch + 1, dim))
        self.pos_drop = nn.Dropout(dropout)
        drop_path_rate = torch.linspace(0, drop_path, depth).tolist()
        self.layers = nn.ModuleList(
            [TransformerLayer(dim, n_head, dim_ff, dropout, drop_attn, drop_ff, dpr) for dpr in drop_path_rate]
        )
        self.norm = LayerNorm(dim)
        self.apply(self.init_weights)
        nn.init.normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=0.02)
        self.head = head
        self.depth = depth

    def set_drop_path(self, drop_path):
        drop_path_rate = torch.linspace(0, drop_path, self.depth).tolist()
        for layer, p in zip(self.layers, drop_path_rate):
            layer.set_drop_path(p)

    def init_weights(self, module):
        if isinstance(module, nn.Linear):
            nn.init.normal_(module.weight, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.LayerNorm):
            nn.init.ones_(module.weight)
            nn.init.zeros_(module.bias)

    def interpolate_pos_embedding(self, n_patch, pos_embed):
        """Bicubic resample of the grid part for multi-crop inputs (vit.py:153-175); identity at 224^2.
        Left to ATen: it touches [1, 196, D] once per step."""
        n_pos = pos_embed.shape[1] - 1
        if n_patch == n_pos:
            return pos_embed
        cls_embed = pos_embed[:, 0]
        dim = pos_embed.shape[-1]
        side = int(math.sqrt(n_pos))
        grid = F.interpolate(
            pos_embed[:, 1:].reshape(1, side, side, dim).permute(0, 3, 1, 2),
            scale_factor=math.sqrt(n_patch / n_pos), mode="bicubic", align_corners=False,
            recompute_scale_factor=False,
        )
        grid = grid.permute(0, 2, 3, 1).reshape(1, -1, dim)
        return torch.cat((cls_embed.unsqueeze(0), grid), 1)

    def forward_feature(self, input):
        from vtb200.blocks import LayerNormFn, ViTPatchEmbedFn

        check_no_dropout(self, self.pos_drop.p)
        pe = self.patch_embedding
        p = pe.window_size
        n_patch = (input.shape[-2] // p) * (input.shape[-1] // p)
        pos = self.interpolate_pos_embedding(n_patch, self.pos_embed)
        out = ViTPatchEmbedFn.apply(input, pe.linear.weight, pe.linear.bias, self.cls_token, pos, p)
        for layer in self.layers:
            out = layer(out)
        # LayerNorm is row-wise, so norm(out)[:, 0] == norm(out[:, 0]) (vit.py:149-151): only cls rows run.
        return LayerNormFn.apply(out[:, 0], self.norm.weight, self.norm.bias, self.norm.eps)

    def forward(self, input):
        if not isinstance(input, (list, tuple)):
            input = [input]
        crops = torch.cumsum(
            torch.unique_consecutive(torch.tensor([i.shape[-1] for i in input]), return_counts=True)[1], 0
        )
        start = 0
        output = None
        for end in crops:
            out = self.forward_feature(torch.cat(input[start:end]))
            output = out if start == 0 else torch.cat((output, out))
            start = end
        if self.head is not None:
            output = self.head(output)
        return output


class FusedLinear(nn.Linear):
    """nn.Linear whose forward/backward run on the tcgen05 GEMM (drop-in head for VisionTransformer)."""

    def forward(self, input):
        from vtb200.blocks import LinearFn

        return LinearFn.apply(input, self.weight, self.bias)


class DINOHead(nn.Module):
    """vit.py:206-262.  Linears run on the tcgen05 GEMM; GELU / L2-normalise / weight-norm stay in ATen for
    now ("next" row, not part of the block hot path)."""

    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, depth=3, dim_ff=2048,
                 dim_bottleneck=256):
        super().__init__()
        if depth == 1:
            self.mlp = nn.Linear(in_dim, dim_bottleneck)
        else:
            layers = [nn.Linear(in_dim, dim_ff)]
            if use_bn:
                layers.append(nn.BatchNorm1d(dim_ff))
            layers.append(nn.GELU())
            for _ in range(depth - 2):
                layers.append(nn.Linear(dim_ff, dim_ff))
                if use_bn:
                    layers.append(nn.BatchNorm1d(dim_ff))
                layers.append(nn.GELU())
            layers.append(nn.Linear(dim_ff, dim_bottleneck))
            self.mlp = nn.Sequential(*layers)
        self.apply(self.init_weights)
        self.last = nn.utils.weight_norm(nn.Linear(dim_bottleneck, out_dim, bias=False))
        self.last.weight_g.detach().fill_(1)
        if norm_last_layer:
            self.last.weight_g.requires_grad = False

    def init_weights(self, module):
        if isinstance(module, nn.Linear):
            nn.init.normal_(module.weight, std=0.02)
            if module.bias is not None:
                nn.init.zeros_(module.bias)

    def forward(self, input):
        from vtb200.blocks import LinearFn

        out = input
        mods = [self.mlp] if isinstance(self.mlp, nn.Linear) else list(self.mlp)
        for m in mods:
            out = LinearFn.apply(out, m.weight, m.bias) if isinstance(m, nn.Linear) else m(out)
        out = F.normalize(out, dim=-1, p=2)
        g, v = self.last.weight_g, self.last.weight_v
        w = v * (g / v.norm(dim=1, keepdim=True))
        return LinearFn.apply(out, w, None)


@config_model(name="dino", namespace="model", use_type=True)
def dino(
    image_size: Union[StrictInt, Tuple[StrictInt, StrictInt]],
    window_size: StrictInt,
    depth: StrictInt,
    dim: StrictInt,
    n_head: StrictInt,
    dim_ff: StrictInt,
    dropout: StrictFloat,
    drop_attn: StrictFloat,
    drop_ff: StrictFloat,
    drop_path: StrictFloat,
    dim_head_out: StrictInt,
    use_bn: StrictBool = False,
    norm_last_layer: StrictBool = True,
    depth_head: StrictInt = 3,
    dim_head_ff: StrictInt = 2048,
    dim_head_bottleneck: StrictInt = 256,
):
    head = DINOHead(dim, dim_head_out, use_bn, norm_last_layer, depth_head, dim_head_ff, dim_head_bottleneck)
    return VisionTransformer(head, image_size, window_size, depth, dim, n_head, dim_ff, dropout, drop_attn,
                             drop_ff, drop_path)
