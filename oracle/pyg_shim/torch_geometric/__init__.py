"""Test-only minimal torch_geometric stand-in (see ../README.md)."""
from . import data, loader, nn, transforms, utils  # noqa: F401
