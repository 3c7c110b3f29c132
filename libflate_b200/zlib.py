"""Mirror of libflate::zlib::{Encoder, Decoder} (src/zlib.rs:284-681); zlib_flush_sync=True is FlushMode::Sync."""
from . import _native as nv
from .deflate import Decoder as _Dec, Encoder as _Enc, InvalidData, UnexpectedEof  # noqa: F401


class Encoder(_Enc):
    FMT = nv.FMT_ZLIB


class Decoder(_Dec):
    FMT = nv.FMT_ZLIB
