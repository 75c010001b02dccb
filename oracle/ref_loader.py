"""Import the UNMODIFIED reference model zoo (rosinality/vision-transformers-pytorch, models/*.py) as the
package `ref_models`, in the build container only (/root/reference does not exist on the GPU box).

The model files need exactly one thing from the absent `tensorfn` library: the `config_model` decorator
(vit.py:7, swin_transformer.py:7, twins.py:7, efficientnet.py:6, nfefficientnet.py:6).  A stub module whose
`config_model(...)` is the identity decorator is injected before import.

TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import importlib.util
import os
import sys
import types

REF_ROOTS = ("/root/reference", os.path.join(os.path.dirname(os.path.dirname(__file__)), "baseline", "_ref"))


def reference_root():
    for r in REF_ROOTS:
        if os.path.isfile(os.path.join(r, "models", "vit.py")):
            return r
    return None


def available():
    return reference_root() is not None


def _install_tensorfn_stub():
    if "tensorfn" in sys.modules:
        return
    tf = types.ModuleType("tensorfn")
    cfg = types.ModuleType("tensorfn.config")

    def config_model(*args, **kwargs):
        def deco(obj):
            return obj

        return deco

    cfg.config_model = config_model
    tf.config = cfg
    sys.modules["tensorfn"] = tf
    sys.modules["tensorfn.config"] = cfg


def load():
    """Returns the reference `models` package imported under the name `ref_models`."""
    if "ref_models" in sys.modules:
        return sys.modules["ref_models"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (expected in the build container only)")
    _install_tensorfn_stub()
    pkg_dir = os.path.join(root, "models")
    # The reference's models/__init__.py also pulls in the CNN zoo, whose nfnet.py does an absolute
    # `from models import layer` (nfnet.py:5) that would resolve to OUR drop-in package.  The transformer
    # files only use relative imports, so expose them through an empty namespace package instead.
    mod = types.ModuleType("ref_models")
    mod.__path__ = [pkg_dir]
    sys.modules["ref_models"] = mod
    for name in ("layer", "vit", "swin_transformer", "pvt", "halo_transformer", "twins"):
        setattr(mod, name, importlib.import_module(f"ref_models.{name}"))
    return mod


def load_reference_module(name):
    """Import a top-level file of the reference tree (e.g. `loss` -> loss.py) under the name `ref_<name>`; None when the
    reference tree is absent (GPU box)."""
    root = reference_root()
    if root is None:
        return None
    key = f"ref_{name}"
    if key in sys.modules:
        return sys.modules[key]
    _install_tensorfn_stub()
    spec = importlib.util.spec_from_file_location(key, os.path.join(root, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod
