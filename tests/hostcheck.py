"""ctypes loader for tests/native/libhostcheck.so (test harness around the product's host/device-portable code)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "native"), "-s"])
        _lib = C.CDLL(os.path.join(_HERE, "native", "_build", "libhostcheck.so"))
        _lib.hc_block_codes.restype = C.c_uint32
        _lib.hc_hdr_words.restype = C.c_uint32
    return _lib


def code_lengths(freq, cap):
    f = np.ascontiguousarray(freq, dtype=np.uint32)
    w = np.zeros(len(f), dtype=np.uint8)
    lib().hc_code_lengths(f.ctypes.data_as(C.c_void_p), len(f), cap, w.ctypes.data_as(C.c_void_p))
    return w


def block_codes(hist320):
    L = lib()
    h = np.ascontiguousarray(hist320, dtype=np.uint32)
    lit = np.zeros(288, dtype=np.uint32)
    dist = np.zeros(32, dtype=np.uint32)
    hdr = np.zeros(L.hc_hdr_words(), dtype=np.uint32)
    nbits = L.hc_block_codes(h.ctypes.data_as(C.c_void_p), lit.ctypes.data_as(C.c_void_p),
                             dist.ctypes.data_as(C.c_void_p), hdr.ctypes.data_as(C.c_void_p))
    return lit, dist, hdr, nbits


def fixed_codes():
    lit = np.zeros(288, dtype=np.uint32)
    dist = np.zeros(32, dtype=np.uint32)
    lib().hc_fixed_codes(lit.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p))
    return lit, dist


def length_code(n):
    o = (C.c_uint32 * 3)()
    lib().hc_length_code(n, o)
    return tuple(o)


def dist_code(d):
    o = (C.c_uint32 * 3)()
    lib().hc_dist_code(d, o)
    return tuple(o)


def inflate(data, cap=None):
    """product inflate core on the host: returns (status, out bytes, consumed, end_bit)"""
    data = bytes(data)
    if cap is None:
        cap = max(1 << 16, len(data) * 1100 + 1024)
    out = C.create_string_buffer(cap + 64)
    res = (C.c_int64 * 4)()
    buf = C.create_string_buffer(data, len(data) + 64)      # padded like device buffers
    lib().hc_inflate(buf, C.c_uint64(len(data)), out, C.c_uint64(cap), res)
    n = min(res[1], cap)
    return res[0], out.raw[:n], res[2], res[3]


def block_starts(data, cap=None):
    data = bytes(data)
    if cap is None:
        cap = max(1 << 16, len(data) * 1100 + 1024)
    scratch = C.create_string_buffer(cap + 64)
    starts = (C.c_uint64 * 8192)()
    buf = C.create_string_buffer(data, len(data) + 64)
    lib().hc_block_starts.restype = C.c_uint32
    n = lib().hc_block_starts(buf, C.c_uint64(len(data)), scratch, C.c_uint64(cap), starts, 8192)
    return list(starts[: min(n, 8192)])


def find_candidates(data):
    data = bytes(data)
    cands = (C.c_uint64 * 65536)()
    buf = C.create_string_buffer(data, len(data) + 64)
    lib().hc_find_candidates.restype = C.c_uint32
    n = lib().hc_find_candidates(buf, C.c_uint64(len(data)), cands, 65536)
    return n, list(cands[: min(n, 65536)])
