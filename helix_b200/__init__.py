"""helix_b200 — B200-native (sm_100a) inference runtime behind Helix's runner.Runtime surface.

The product is the C-ABI shared library ``libhelixb200.so`` (``include/helix_b200.h``); this package is
the thin Python host binding used by the tests, the benchmark and the OpenAI-compatible shim.
There is no CPU fallback: importing works anywhere, but every compute entry point needs a B200.
"""
from ._lib import lib, load_library, LibraryMissing  # noqa: F401
from .engine import Engine, EngineConfig, ModelDesc, Sampling, HBError  # noqa: F401
from . import configs  # noqa: F401

__all__ = ["lib", "load_library", "LibraryMissing", "Engine", "EngineConfig", "ModelDesc", "Sampling", "HBError", "configs"]
