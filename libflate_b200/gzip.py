"""Mirror of libflate::gzip::{Encoder, Decoder, MultiDecoder} (src/gzip.rs:754-1167).  Header options map to
HeaderBuilder: mtime, os_, is_text, is_verified, extra, filename, comment (libflate defaults mtime to now(); pass it explicitly)."""
from . import _native as nv
from .deflate import Decoder as _Dec, Encoder as _Enc, InvalidData, UnexpectedEof  # noqa: F401


class Encoder(_Enc):
    FMT = nv.FMT_GZIP


class Decoder(_Dec):
    FMT = nv.FMT_GZIP


class MultiDecoder(_Dec):
    FMT = nv.FMT_GZIP_MULTI
