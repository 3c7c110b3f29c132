"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/flate_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

OK, INVALID_DATA, UNEXPECTED_EOF, OUTPUT_TOO_SMALL = 0, -1, -2, -3
FMT_DEFLATE, FMT_ZLIB, FMT_GZIP, FMT_GZIP_MULTI = 0, 1, 2, 3
MODE_DYNAMIC, MODE_FIXED, MODE_STORED = 0, 1, 2
FLUSH = -1  # schedule entry: io::Write::flush()


class Opts(C.Structure):
    _fields_ = [
        ("block_size", C.c_uint64), ("window_size", C.c_uint32), ("max_length", C.c_uint32),
        ("mode", C.c_int32), ("zlib_flush_sync", C.c_int32), ("gzip_mtime", C.c_uint32),
        ("gzip_os", C.c_uint8), ("gzip_is_text", C.c_uint8), ("gzip_is_verified", C.c_uint8),
        ("gzip_has_extra", C.c_uint8), ("gzip_extra", C.c_char_p), ("gzip_extra_len", C.c_uint32),
        ("gzip_filename", C.c_char_p), ("gzip_comment", C.c_char_p),
    ]


def build(force=False):
    src = os.path.join(_HERE, "flate_oracle.c")
    if force or not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_encode.restype = C.c_int
        L.orc_encode.argtypes = [C.c_int, C.POINTER(Opts), C.c_char_p, C.c_size_t, C.POINTER(C.c_int64), C.c_size_t,
                                 C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_decode.restype = C.c_int
        L.orc_decode.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_size_t), C.c_char_p]
        L.orc_lz77_default.restype = C.c_int
        L.orc_lz77_default.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_size_t)]
        L.orc_huffman_lengths.restype = C.c_int
        L.orc_huffman_lengths.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_void_p]
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.orc_adler32.restype = C.c_uint32
        L.orc_adler32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.orc_dynamic_header_loads.restype = C.c_int
        L.orc_dynamic_header_loads.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_opts_size.restype = C.c_size_t
        assert L.orc_opts_size() == C.sizeof(Opts)
        _lib = L
    return _lib


def make_opts(block_size=1 << 20, window_size=32768, max_length=258, mode=MODE_DYNAMIC, zlib_flush_sync=False,
              mtime=0, os_=3, is_text=False, is_verified=False, extra=None, filename=None, comment=None):
    o = Opts()
    o.block_size, o.window_size, o.max_length, o.mode = block_size, window_size, max_length, mode
    o.zlib_flush_sync = 1 if zlib_flush_sync else 0
    o.gzip_mtime, o.gzip_os = mtime, os_
    o.gzip_is_text, o.gzip_is_verified = int(is_text), int(is_verified)
    o.gzip_has_extra = 1 if extra is not None else 0
    o.gzip_extra = extra if extra is not None else None
    o.gzip_extra_len = len(extra) if extra is not None else 0
    o.gzip_filename, o.gzip_comment = filename, comment
    return o


def encode(fmt, data, schedule=None, **kw):
    """schedule: None (single write_all) or list of ints (write sizes, FLUSH == -1 for flush())."""
    L = lib()
    o = make_opts(**kw)
    data = bytes(data)
    cap = len(data) + len(data) // 2 + 4096 + (64 * len(schedule) if schedule else 0)
    out = C.create_string_buffer(cap)
    n = C.c_size_t(0)
    if schedule is None:
        rc = L.orc_encode(fmt, C.byref(o), data, len(data), None, 0, out, cap, C.byref(n))
    else:
        arr = (C.c_int64 * len(schedule))(*schedule)
        rc = L.orc_encode(fmt, C.byref(o), data, len(data), arr, len(schedule), out, cap, C.byref(n))
    if rc != OK:
        raise RuntimeError(f"orc_encode rc={rc}")
    return out.raw[: n.value]


def decode(fmt, data, cap=None):
    """returns (rc, output bytes, in_consumed, message)"""
    L = lib()
    data = bytes(data)
    if cap is None:
        cap = max(1 << 16, len(data) * 1100 + 1024)
    out = C.create_string_buffer(cap)
    n, used = C.c_size_t(0), C.c_size_t(0)
    msg = C.create_string_buffer(128)
    rc = L.orc_decode(fmt, data, len(data), out, cap, C.byref(n), C.byref(used), msg)
    return rc, out.raw[: min(n.value, cap)], used.value, msg.value.decode("utf-8", "replace")


def lz77_default(data, window=32768, max_len=258):
    import numpy as np
    L = lib()
    data = bytes(data)
    codes = np.empty(max(len(data), 1), dtype=np.uint32)
    n = C.c_size_t(0)
    L.orc_lz77_default(data, len(data), window, max_len, codes.ctypes.data, C.byref(n))
    return codes[: n.value].copy()


def huffman_lengths(freqs, max_bitwidth):
    import numpy as np
    L = lib()
    f = (C.c_uint64 * len(freqs))(*[int(x) for x in freqs])
    w = np.zeros(len(freqs), dtype=np.uint8)
    rc = L.orc_huffman_lengths(f, len(freqs), max_bitwidth, w.ctypes.data)
    assert rc == 0
    return w


def crc32(data, init=0):
    return lib().orc_crc32(init, bytes(data), len(data))


def adler32(data, init=1):
    return lib().orc_adler32(init, bytes(data), len(data))


def dynamic_header_loads(data):
    return lib().orc_dynamic_header_loads(bytes(data), len(data))
