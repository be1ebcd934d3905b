"""maua_style_b200 -- B200-native (sm_100a) VGG-19 neural-style inner loop behind the maua-style module API.

`loss`, `models` and `optim` mirror the reference's modules of the same names (JCBrouwer/maua-style); the
arithmetic runs in libmaua_b200.so (hand-written CUDA: tcgen05/TMEM/TMA implicit-GEMM convolutions, a tensor-core
SYRK for the Gram matrix, fused memory-bound kernels for TV / content / Adam / L-BFGS).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "loss", "models", "optim", "shard", "parallel", "synthetic"]
__version__ = "0.1.0"
