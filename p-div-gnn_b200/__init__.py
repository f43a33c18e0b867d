"""pdivgnn_b200 -- B200-native (sm_100a) P-DivGNN message-passing hot path.

Importing the package never touches CUDA; the C-ABI library ``lib/libpdivgnn.so`` is
loaded on first use and the product path raises if it is missing (no CPU fallback).
"""
__version__ = "0.1.0"
