"""Host-side mirror of libflate::deflate::{Encoder, Decoder} (src/deflate/encode.rs:132-258, decode.rs:8-165) over the
streaming handles of the C ABI (b2f_encoder_*, b2f_decoder_*).  Same method names and argument meaning: write() always
consumes everything, flush() forces a non-final block, finish() returns the complete stream; Decoder.read() follows
io::Read (b'' at end of stream; InvalidData / UnexpectedEof raise IOError subclasses)."""
import ctypes as C

from . import _native as nv


class InvalidData(IOError):
    """io::ErrorKind::InvalidData"""


class UnexpectedEof(IOError):
    """io::ErrorKind::UnexpectedEof"""


def _raise(code, what):
    if code == nv.ERR_INVALID_DATA:
        raise InvalidData(what)
    if code == nv.ERR_UNEXPECTED_EOF:
        raise UnexpectedEof(what)
    raise nv.B2fError(code, what)


class Encoder:
    FMT = nv.FMT_DEFLATE

    def __init__(self, ctx, **options):
        self._ctx = ctx
        self._opts = nv.make_opts(**options)
        self._h = C.c_void_p()
        rc = nv.lib().b2f_encoder_new(ctx.handle, self.FMT, C.byref(self._opts), C.byref(self._h))
        if rc:
            raise nv.B2fError(rc, "b2f_encoder_new")

    def write(self, buf):
        buf = bytes(buf)
        rc = nv.lib().b2f_encoder_write(self._h, buf, len(buf))
        if rc:
            raise nv.B2fError(rc, "write")
        return len(buf)

    def write_all(self, buf):
        if len(buf):
            self.write(buf)

    def flush(self):
        rc = nv.lib().b2f_encoder_flush(self._h)
        if rc:
            raise nv.B2fError(rc, "flush")

    def finish(self):
        p, n = C.c_void_p(), C.c_size_t()
        rc = nv.lib().b2f_encoder_finish(self._h, C.byref(p), C.byref(n))
        if rc:
            raise nv.B2fError(rc, (nv.lib().b2f_last_error(self._ctx.handle) or b"").decode())
        return C.string_at(p, n.value)

    def __del__(self):
        if getattr(self, "_h", None):
            nv.lib().b2f_encoder_free(self._h)
            self._h = None


class Decoder:
    FMT = nv.FMT_DEFLATE

    def __init__(self, ctx, data):
        self._ctx = ctx
        self._data = bytes(data)
        self._h = C.c_void_p()
        rc = nv.lib().b2f_decoder_new(ctx.handle, self.FMT, self._data, len(self._data), C.byref(self._h))
        if rc:
            raise nv.B2fError(rc, "b2f_decoder_new")

    def read(self, n):
        buf = C.create_string_buffer(max(n, 1))
        k = nv.lib().b2f_decoder_read(self._h, buf, n)
        if k < 0:
            _raise(k, "read")
        return buf.raw[:k]

    def read_to_end(self):
        out = bytearray()
        while True:
            c = self.read(1 << 20)
            if not c:
                return bytes(out)
            out += c

    def unread_decoded_data(self):
        p = C.c_void_p()
        n = nv.lib().b2f_decoder_unread(self._h, C.byref(p))
        return C.string_at(p, n) if n else b""

    def consumed(self):
        return nv.lib().b2f_decoder_consumed(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            nv.lib().b2f_decoder_free(self._h)
            self._h = None
