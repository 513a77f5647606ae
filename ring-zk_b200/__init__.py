"""ring-zk_b200: B200-native batched engine for ring-zk's R_q hot path.

Import with importlib.import_module("ring-zk_b200") (the directory name carries a
hyphen).  Sub-modules:
  synth   seeded synthetic inputs (host-side randomness r, y, d)
  engine  ctypes binding of the C ABI in include/ringzk_b200.h (CUDA only, no fallback)
  api     host-side mirror of the reference's public Rust API on top of `engine`
"""
from . import synth  # noqa: F401

__all__ = ["synth", "engine", "api", "shard"]
