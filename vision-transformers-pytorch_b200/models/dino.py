"""DINO projection head and the `dino` model factory — drop-in for vit.py:206-307 of the reference.

DINOHead is a "next" row of the hot-path table: its Linears run on the tcgen05 GEMM, GELU, the L2
normalisation and the weight-norm reparametrisation on row kernels that emit the bf16 GEMM operands directly
(vtb_gelu_*, vtb_l2norm_*, vtb_weight_norm_*).  BatchNorm1d (use_bn=True; no reference config sets it) stays ATen."""
from typing import Tuple, Union

from pydantic import StrictBool, StrictFloat, StrictInt
from torch import nn

from ._compat import config_model


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, depth=3, dim_ff=2048, dim_bottleneck=256):
        super().__init__()
        if depth == 1:
            self.mlp = nn.Linear(in_dim, dim_bottleneck)
        else:
            widths = [in_dim] + [dim_ff] * (depth - 1)
            blocks = []
            for w_in, w_out in zip(widths[:-1], widths[1:]):
                blocks.append(nn.Linear(w_in, w_out))
                if use_bn:
                    blocks.append(nn.BatchNorm1d(w_out))
                blocks.append(nn.GELU())
            blocks.append(nn.Linear(dim_ff, dim_bottleneck))
            self.mlp = nn.Sequential(*blocks)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        # weight-normed output layer: parameters `last.weight_g` / `last.weight_v` (train_util.py:29-31 keys on "last")
        self.last = nn.utils.weight_norm(nn.Linear(dim_bottleneck, out_dim, bias=False))
        self.last.weight_g.detach().fill_(1)
        self.last.weight_g.requires_grad = not norm_last_layer

    def forward(self, input):
        from vtb200.blocks import GeluFn, LinearFn, NormLinearFn

        out = input
        for m in ([self.mlp] if isinstance(self.mlp, nn.Linear) else self.mlp):
            if isinstance(m, nn.Linear):
                out = LinearFn.apply(out, m.weight, m.bias)
            elif isinstance(m, nn.GELU):
                out = GeluFn.apply(out)
            else:
                out = m(out)
        return NormLinearFn.apply(out, self.last.weight_v, self.last.weight_g)


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
    from .vit import VisionTransformer

    head = DINOHead(dim, dim_head_out, use_bn=use_bn, norm_last_layer=norm_last_layer, depth=depth_head,
                    dim_ff=dim_head_ff, dim_bottleneck=dim_head_bottleneck)
    return VisionTransformer(head, image_size, window_size, depth, dim, n_head, dim_ff, dropout, drop_attn, drop_ff,
                             drop_path)
