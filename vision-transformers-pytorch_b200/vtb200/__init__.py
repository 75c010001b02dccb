"""vtb200 — host side of the Blackwell-native transformer-block path."""
from . import lib  # noqa: F401

__all__ = ["lib", "ops", "blocks"]
