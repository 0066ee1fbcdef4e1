"""chemlab_b200 -- B200-native engine for chemlab's reactive-MD hot path.

Layout: csrc/ (CUDA kernels + C-ABI, built to lib/libchemlab_b200.so), engine.py (ctypes mirror of
include/chemlab_b200.h), espressopp/ (the espressopp-compatible surface chemlab's driver imports).
There is no CPU fallback: importing works without a GPU, creating an Engine does not.
"""
from .engine import Engine, EngineError  # noqa: F401

__version__ = "0.1.0"
