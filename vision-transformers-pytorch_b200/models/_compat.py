"""`tensorfn.config.config_model` if tensorfn is installed (the reference registers its models with it:
vit.py:265, swin_transformer.py:236, twins.py:220), otherwise an identity decorator so that the zoo
imports without it (tensorfn is not a dependency of the hot path — SURVEY §2 row 18)."""
try:  # pragma: no cover - tensorfn is absent in this image
    from tensorfn.config import config_model  # type: ignore
except Exception:  # noqa: BLE001

    def config_model(*args, **kwargs):
        def deco(obj):
            return obj

        return deco


def reference_models_dir():
    """Directory of the REFERENCE's models/ package, for the parts of the zoo this drop-in does not replace (the CNN
    families).  Searched: $VTB_REFERENCE_ROOT, then every sys.path entry and the working directory (the documented
    install puts this package in front of the reference checkout on PYTHONPATH, so the checkout is still on the path)."""
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    roots = [os.environ.get("VTB_REFERENCE_ROOT", "")] + list(sys.path) + [os.getcwd(), "/root/reference"]
    for r in roots:
        d = os.path.join(r or ".", "models")
        if os.path.isfile(os.path.join(d, "nfnet.py")) and os.path.abspath(d) != here:
            return d
    return None


def load_reference_file(stem, as_name=None):
    """Import <reference>/models/<stem>.py as a submodule of THIS package (so its absolute `from models import layer`
    / `from models.layer import DropPath` statements resolve here).  Raises ImportError when the reference checkout is not
    importable — the CNN families are not re-implemented (they are outside the transformer-block hot path)."""
    import importlib.util
    import os
    import sys

    name = as_name or f"models.{stem}"
    if name in sys.modules:
        return sys.modules[name]
    d = reference_models_dir()
    if d is None:
        raise ImportError(
            f"models.{stem} belongs to the reference's CNN zoo, which vtb200 does not replace; put the reference checkout "
            "on sys.path (behind this package) or set VTB_REFERENCE_ROOT so that it can be loaded from there")
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, stem + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules[name]
        raise
    return mod
