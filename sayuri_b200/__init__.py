"""sayuri_b200 — Blackwell-native (sm_100a) batched NN evaluation for Sayuri's self-play hot path.

The product is the C-ABI shared library built from sayuri_b200/csrc (see include/sayuri_b200.h);
this package is the thin Python host mirror used by tests/ and bench.py.
"""
__version__ = "0.1.0"
