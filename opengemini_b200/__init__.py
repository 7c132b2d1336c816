"""opengemini_b200 — B200-native scan/aggregate path behind openGemini's cursor seam.

Product = libogpu.so (hand-written sm_100a CUDA behind the C ABI in include/ogpu.h).
This package only holds the host-side bindings; aggregation and decoding never run on the CPU
(the bindings only slice the records the library returns).
"""
from . import _lib  # noqa: F401
from .cursor import AggQuery, Comm, ScanCursor, Shard  # noqa: F401

__all__ = ["Shard", "AggQuery", "ScanCursor", "Comm", "_lib"]
