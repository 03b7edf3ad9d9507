"""Minimal fairseq surface for importing the reference's hot-path files.
Test infrastructure only; see ../README.md."""
from . import utils, metrics  # noqa: F401
def __getattr__(name):
    if name == "search":
        import importlib
        return importlib.import_module("fairseq.search")
    raise AttributeError(name)
