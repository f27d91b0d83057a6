"""ldpc_decoders_b200 — B200-native (sm_100a) batch decoders for the one hot path of
thadikari/ldpc_decoders: iterative SPA / MSA over H for BSC and BIAWGN, and BEC erasure decoding.

The package mirrors the reference's decoder protocol (``bpa``, ``bsc``, ``biawgn``, ``bec``,
``models``) on top of a C-ABI CUDA library (include/ldpc_b200.h).  No CPU fallback exists.
"""
from . import _lib, graph                                   # noqa: F401
from ._lib import LdpcError                                 # noqa: F401
from .graph import Tables                                   # noqa: F401

__all__ = ["Tables", "LdpcError", "bpa", "bsc", "biawgn", "bec", "models", "engine"]


def __getattr__(name):
    # decoder modules are imported lazily so that `import ldpc_decoders_b200` works on a box
    # without the built library (the CPU test tier checks tables and the ABI without a GPU)
    if name in ("bpa", "bsc", "biawgn", "bec", "models", "engine", "codes", "sim", "dist"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
