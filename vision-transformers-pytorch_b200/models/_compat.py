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
