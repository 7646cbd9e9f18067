"""phylign_b200 -- B200-native match stage of Phylign (COBS classic query + top-N filter).

The package is a thin Python driver over the C-ABI CUDA library libphylign_cuda.so
(include/phylign_cuda.h).  No PyTorch, no CPU fallback.
"""
__version__ = "0.1.0"
