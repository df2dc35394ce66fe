"""nhwcodec_b200 -- B200 (sm_100a) implementation of the NHW codec hot path.

The product is libnhw_cuda.so (hand-written CUDA behind the C-ABI in include/nhw_cuda.h);
this package is the thin Python host side used by bench.py and the tests.  PyTorch is used
only for device memory, streams and torch.distributed plumbing.

There is no CPU fallback: importing works anywhere (so CPU-only test collection works), but
creating a `Codec` without the compiled library or without a GPU raises.
"""
from .capi import Codec, NhwError, load_library, LIB_PATH  # noqa: F401

__all__ = ["Codec", "NhwError", "load_library", "LIB_PATH"]
