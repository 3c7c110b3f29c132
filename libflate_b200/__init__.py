"""libflate_b200: B200-native DEFLATE hot path (CUDA kernels behind the C ABI in include/b2f.h).

The product is libflate_b200/libb2f.so; this package is the thin Python harness over that ABI used by the
tests and bench.py, plus host-side mirrors of libflate's Encoder/Decoder surface (libflate_b200.{deflate,zlib,gzip})."""
from . import _native as native          # noqa: F401
from ._native import Context             # noqa: F401
