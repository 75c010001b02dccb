"""Drop-in `models` package: same exports as the reference's models/__init__.py:1-7 for the transformer
families (the hot path).  The CNN zoo (NFNet / EfficientNet / NF-EfficientNetV2: cuDNN conv nets) is outside
the transformer-block path and is not re-implemented here."""
from .halo_transformer import HaloTransformer
from .pvt import PyramidVisionTransformer
from .swin_transformer import SwinTransformer
from .twins import TwinsSVT  # noqa: F401  (not exported by the reference's __init__, registered as "twins_svt")
from .dino import DINOHead, dino
from .vit import FusedLinear, VisionTransformer

__all__ = ["HaloTransformer", "PyramidVisionTransformer", "SwinTransformer", "VisionTransformer",
           "DINOHead", "FusedLinear", "dino"]
