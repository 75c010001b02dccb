"""Drop-in `models` package: same exports as the reference's models/__init__.py:1-7.  The transformer families (the
hot path) are implemented here; the CNN zoo (NFNet / efficientnet / efficientnetv2 / nfefficientnetv2: cuDNN conv nets,
SURVEY §2 row 8) is outside the transformer-block path and resolves LAZILY to the reference's own files when its checkout
is importable (models/_compat.py), so that `config/efficientnetv2-s.conf` keeps working under the documented install
(this package in front of the reference on PYTHONPATH).  Without the checkout those four names raise ImportError."""
from .halo_transformer import HaloTransformer
from .pvt import PyramidVisionTransformer
from .swin_transformer import SwinTransformer
from .twins import TwinsSVT  # noqa: F401  (not exported by the reference's __init__, registered as "twins_svt")
from .dino import DINOHead, dino
from .vit import FusedLinear, VisionTransformer

__all__ = ["HaloTransformer", "PyramidVisionTransformer", "SwinTransformer", "VisionTransformer",
           "DINOHead", "FusedLinear", "dino"]

_CNN_ZOO = {"NFNet": "nfnet", "efficientnet": "efficientnet", "efficientnetv2": "efficientnet",
            "nfefficientnetv2": "nfefficientnet"}
__all__ += list(_CNN_ZOO)


def __getattr__(name):
    if name in _CNN_ZOO:
        from ._compat import load_reference_file

        return getattr(load_reference_file(_CNN_ZOO[name]), name)
    raise AttributeError(f"module 'models' has no attribute {name!r}")
