"""Shared layer holders (reference: models/layer.py:25,166-196).

These modules only HOLD parameters / hyper-parameters under the reference's names; the arithmetic of a
block runs in vtb200.blocks (one autograd.Function per pre-LN branch)."""
from collections import abc
from itertools import repeat

import torch
from torch import nn


def ensure_tuple(x, n_item):
    if isinstance(x, abc.Iterable):
        try:
            if len(x) != n_item:
                raise ValueError(f"length of {x} (length: {len(x)}) does not match n_item={n_item}")
        except TypeError:
            pass
        return x
    return tuple(repeat(x, n_item))


def tuple2(x):
    return ensure_tuple(x, 2)


class DropPath(nn.Module):
    """Holder of the stochastic-depth rate `p` (layer.py:166-183).  The per-sample keep mask is drawn by
    vtb200.blocks.make_drop_path_scale and applied inside the residual GEMM epilogue."""

    def __init__(self, p=0):
        super().__init__()
        self.p = p

    def scale(self, batch, like=torch.bfloat16):
        from vtb200.blocks import make_drop_path_scale

        return make_drop_path_scale(self.training, self.p, batch, like)

    def forward(self, input):
        """Stand-alone use (the CNN zoo calls `DropPath` as a module: efficientnet.py imports it from models.layer):
        keep-mask per sample / keep probability, identity in eval mode or at p = 0 (layer.py:171-181).  The transformer
        blocks never call this: there the mask rides in the residual GEMM epilogue (`scale`)."""
        if not self.training or self.p == 0:
            return input
        keep = 1 - self.p
        mask = input.new_empty([input.shape[0]] + [1] * (input.ndim - 1)).bernoulli_(keep)
        return input / keep * mask

    def __repr__(self):
        return f"{self.__class__.__name__}(p={self.p})"


class PositionwiseFeedForward(nn.Sequential):
    """Linear - SiLU - Dropout - Linear parameter holder (layer.py:186-196); indices 0 and 3 are the
    Linears, exactly as in the reference state_dict."""

    def __init__(self, in_dim, dim=None, out_dim=None, activation=nn.SiLU, dropout=0):
        dim = in_dim if dim is None else dim
        out_dim = in_dim if out_dim is None else out_dim
        if activation is not nn.SiLU:
            raise NotImplementedError("vtb200: the fused FFN kernel implements SiLU (the only activation "
                                      "any reference block uses, layer.py:187)")
        super().__init__(nn.Linear(in_dim, dim), activation(), nn.Dropout(dropout), nn.Linear(dim, out_dim))

    def params(self):
        return self[0].weight, self[0].bias, self[3].weight, self[3].bias


def check_no_dropout(module, *ps):
    """Dropout on the attention PROBABILITIES (vit.py:39, swin:144, pvt.py:62, halo:101, twins) with p > 0 in training mode
    is rejected: the probabilities never leave the attention kernels' tensor memory, a mask there would have to be drawn
    inside every kernel (every reference config uses drop_attn = 0).  The other nn.Dropout sites — FFN hidden, ViT branch
    outputs / token embedding, PVT patch embedding — are implemented (vtb200.blocks.make_dropout_keep / vtb_dropout)."""
    if module.training and any(float(p) > 0 for p in ps):
        raise NotImplementedError("vtb200: dropout on the attention probabilities (drop_attn > 0) is not implemented in the "
                                  "tcgen05 attention kernels (all reference configs use drop_attn = 0); dropout / drop_ff "
                                  "are supported")


def autocast_dtype(x):
    """dtype an activation has at this point of the reference's forward: the autocast dtype inside torch.autocast
    (train.py:273), the input's own dtype otherwise — torch's dropout kernel draws its mask per dtype-sized vector."""
    return torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype


def ffn_branch(x, drop_path, norm, ff, rows_per_sample, out_dropout=None):
    """x + drop_path(out_dropout(ff(norm(x)))).  RNG draws in the reference's order: the Dropout inside the FFN
    (layer.py:194), the Dropout on the branch output (ViT only, vit.py:61), then the DropPath mask."""
    from vtb200.blocks import FFNBranchFn, make_dropout_keep

    w1, b1, w2, b2 = ff.params()
    dt = autocast_dtype(x)
    ff_drop = make_dropout_keep(ff.training, ff[2].p, (*x.shape[:-1], w1.shape[0]), dt, x.device)
    out_drop = None if out_dropout is None else make_dropout_keep(out_dropout.training, out_dropout.p, x.shape, dt, x.device)
    return FFNBranchFn.apply(x, drop_path.scale(x.shape[0]), norm.eps, rows_per_sample, norm.weight,
                             norm.bias, w1, b1, w2, b2, ff_drop, out_drop)


def init_transformer_weights(module, std=0.02):
    """The initialisation every transformer family of the zoo applies through `self.apply(...)`
    (vit.py:128-137, swin:323-333, pvt.py:227-237, halo:225-235, twins:306-316): N(0, std) Linear weights,
    zero Linear biases, unit LayerNorm."""
    if isinstance(module, nn.Linear):
        nn.init.normal_(module.weight, std=std)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.LayerNorm):
        nn.init.ones_(module.weight)
        nn.init.zeros_(module.bias)


def transformer_layers(stages):
    """Every sub-module with a `set_drop_path` method, in stage order (patch embeds / merges / PEGs are skipped)."""
    return [m for stage in stages for m in stage if hasattr(m, "set_drop_path")]


def assign_drop_path(layers, rates):
    for layer, p in zip(layers, rates):
        layer.set_drop_path(float(p))


def linspace_rates(drop_path, n):
    """torch.linspace(0, drop_path, n) as python floats (vit.py:104, pvt.py:207)."""
    return torch.linspace(0, drop_path, n).tolist()


def ramp_rates(drop_path, n):
    """drop_path * i / n for i < n (swin:287-288, twins:268-269)."""
    return [drop_path * float(i) / n for i in range(n)]


def make_classifier(dim, n_class, std=0.02):
    """AdaptiveAvgPool2d(1) - Flatten - Linear holder with the reference's names/indices (swin:278-281)."""
    linear = nn.Linear(dim, n_class)
    nn.init.normal_(linear.weight, std=std)
    nn.init.zeros_(linear.bias)
    return nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(1), linear)


def __getattr__(name):
    """CNN-only helpers of the reference's models/layer.py (ScaledActivation, WSConv2d, StochasticDepth, SqueezeExcite,
    GlobalContext: used by nfnet.py / nfefficientnet.py, outside the transformer hot path) resolve lazily to the
    reference's own definitions when its tree is importable (models/_compat.py: reference_file)."""
    if name.startswith("__"):
        raise AttributeError(name)
    from ._compat import load_reference_file

    ref = load_reference_file("layer", as_name="models._reference_layer")
    try:
        return getattr(ref, name)
    except AttributeError:
        raise AttributeError(f"module 'models.layer' has no attribute {name!r}") from None
