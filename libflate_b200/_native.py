"""ctypes binding of libb2f.so (include/b2f.h).  Python is only a harness around the C ABI: tests and bench.py
call the same entry points a Rust/C++ host would.  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libb2f.so")
if os.environ.get("B2F_LIB"):          # developer switch: load another build of the same library (kernel experiments)
    SO_PATH = os.path.abspath(os.environ["B2F_LIB"])

OK, ERR_INVALID_DATA, ERR_UNEXPECTED_EOF, ERR_OUTPUT_TOO_SMALL, ERR_NOMEM, ERR_CUDA, ERR_INVALID_ARG = 0, -1, -2, -3, -4, -5, -6
FMT_DEFLATE, FMT_ZLIB, FMT_GZIP, FMT_GZIP_MULTI = 0, 1, 2, 3
MODE_DYNAMIC, MODE_FIXED, MODE_STORED = 0, 1, 2
FLUSH = -1


class EncodeOpts(C.Structure):
    _fields_ = [
        ("block_size", C.c_uint64), ("window_size", C.c_uint32), ("max_length", C.c_uint32),
        ("mode", C.c_int32), ("zlib_flush_sync", C.c_int32), ("gzip_mtime", C.c_uint32),
        ("gzip_os", C.c_uint8), ("gzip_is_text", C.c_uint8), ("gzip_is_verified", C.c_uint8),
        ("gzip_has_extra", C.c_uint8), ("gzip_extra", C.c_char_p), ("gzip_extra_len", C.c_uint32),
        ("gzip_filename", C.c_char_p), ("gzip_comment", C.c_char_p),
    ]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("last_kernel_ms", C.c_float * 16), ("last_n_stages", C.c_uint32),
                ("last_device_ms", C.c_float), ("decode_parallel_streams", C.c_uint64), ("decode_inorder_streams", C.c_uint64),
                ("staged_h2d_bytes", C.c_uint64), ("staged_d2h_bytes", C.c_uint64)]


EXPORTS = [
    "b2f_encode_opts_default", "b2f_ctx_create", "b2f_ctx_destroy", "b2f_last_error", "b2f_version",
    "b2f_plan_from_writes", "b2f_lz77_default", "b2f_encode_batch", "b2f_encode_bound", "b2f_decode_batch",
    "b2f_adler32_batch", "b2f_crc32_batch", "b2f_encode_device", "b2f_decode_device", "b2f_header_len",
    "b2f_encoder_new", "b2f_encoder_write", "b2f_encoder_flush", "b2f_encoder_finish", "b2f_encoder_free",
    "b2f_decoder_new", "b2f_decoder_read", "b2f_decoder_unread", "b2f_decoder_consumed", "b2f_decoder_free",
    "b2f_get_stats", "b2f_stage_name", "b2f_ctx_stream", "b2f_ctx_set_overlap",
    "b2f_host_alloc", "b2f_host_free", "b2f_host_register", "b2f_host_unregister",
    "b2f_encode_part_device", "b2f_bits_shift_device", "b2f_crc32_combine", "b2f_adler32_combine", "b2f_stream_header", "b2f_stream_trailer",
]

_lib = None


def lib():
    """Loads libb2f.so; raises if it has not been built (there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA library is the product; there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        vp, sz, u64 = C.c_void_p, C.c_size_t, C.c_uint64
        L.b2f_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.b2f_ctx_destroy.argtypes = [vp]
        L.b2f_last_error.restype = C.c_char_p
        L.b2f_last_error.argtypes = [vp]
        L.b2f_version.restype = C.c_char_p
        L.b2f_encode_opts_default.argtypes = [C.POINTER(EncodeOpts)]
        L.b2f_plan_from_writes.argtypes = [vp, sz, u64, u64, C.c_uint32, vp, C.POINTER(sz), vp, vp, vp, C.POINTER(sz)]
        L.b2f_lz77_default.argtypes = [vp, vp, sz, C.c_uint32, C.c_uint32, vp, C.POINTER(sz)]
        L.b2f_encode_batch.argtypes = [vp, C.c_int, C.POINTER(EncodeOpts), sz, vp, vp, vp, vp, vp, vp, vp, vp]
        L.b2f_encode_bound.restype = sz
        L.b2f_encode_bound.argtypes = [sz, sz, C.POINTER(EncodeOpts)]
        L.b2f_decode_batch.argtypes = [vp, C.c_int, sz, vp, vp, vp, vp, vp, vp, vp]
        L.b2f_adler32_batch.argtypes = [vp, sz, vp, vp, vp, vp]
        L.b2f_crc32_batch.argtypes = [vp, sz, vp, vp, vp, vp]
        L.b2f_encode_device.argtypes = [vp, C.c_int, C.POINTER(EncodeOpts), sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.b2f_decode_device.argtypes = [vp, C.c_int, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.b2f_header_len.restype = sz
        L.b2f_header_len.argtypes = [C.c_int, C.POINTER(EncodeOpts)]
        L.b2f_encoder_new.argtypes = [vp, C.c_int, C.POINTER(EncodeOpts), C.POINTER(vp)]
        L.b2f_encoder_write.argtypes = [vp, vp, sz]
        L.b2f_encoder_flush.argtypes = [vp]
        L.b2f_encoder_finish.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
        L.b2f_encoder_free.argtypes = [vp]
        L.b2f_decoder_new.argtypes = [vp, C.c_int, vp, sz, C.POINTER(vp)]
        L.b2f_decoder_read.restype = C.c_int64
        L.b2f_decoder_read.argtypes = [vp, vp, sz]
        L.b2f_decoder_unread.restype = sz
        L.b2f_decoder_unread.argtypes = [vp, C.POINTER(vp)]
        L.b2f_decoder_consumed.restype = sz
        L.b2f_decoder_consumed.argtypes = [vp]
        L.b2f_decoder_free.argtypes = [vp]
        L.b2f_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.b2f_stage_name.restype = C.c_char_p
        L.b2f_stage_name.argtypes = [vp, C.c_uint32]
        L.b2f_ctx_stream.restype = vp
        L.b2f_ctx_stream.argtypes = [vp]
        L.b2f_ctx_set_overlap.argtypes = [vp, C.c_int]
        L.b2f_host_alloc.argtypes = [sz, C.POINTER(vp)]
        L.b2f_host_free.argtypes = [vp]
        L.b2f_host_register.argtypes = [vp, sz]
        L.b2f_host_unregister.argtypes = [vp]
        L.b2f_encode_part_device.argtypes = [vp, C.POINTER(EncodeOpts), vp, sz, vp, sz, C.c_int, vp, sz, C.POINTER(u64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.b2f_bits_shift_device.argtypes = [vp, vp, u64, C.c_uint32, vp]
        L.b2f_crc32_combine.restype = C.c_uint32
        L.b2f_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, u64]
        L.b2f_adler32_combine.restype = C.c_uint32
        L.b2f_adler32_combine.argtypes = [C.c_uint32, C.c_uint32, u64]
        L.b2f_stream_header.restype = sz
        L.b2f_stream_header.argtypes = [C.c_int, C.POINTER(EncodeOpts), vp, sz]
        L.b2f_stream_trailer.restype = sz
        L.b2f_stream_trailer.argtypes = [C.c_int, C.c_uint32, C.c_uint32, u64, vp]
        _lib = L
    return _lib


class B2fError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"libb2f error {code}: {msg}")
        self.code = code


def make_opts(block_size=1 << 20, window_size=32768, max_length=258, mode=MODE_DYNAMIC, zlib_flush_sync=False,
              mtime=0, os_=3, is_text=False, is_verified=False, extra=None, filename=None, comment=None):
    o = EncodeOpts()
    lib().b2f_encode_opts_default(C.byref(o))
    o.block_size, o.window_size, o.max_length, o.mode = block_size, window_size, max_length, mode
    o.zlib_flush_sync = 1 if zlib_flush_sync else 0
    o.gzip_mtime, o.gzip_os = mtime, os_
    o.gzip_is_text, o.gzip_is_verified = int(is_text), int(is_verified)
    o.gzip_has_extra = 1 if extra is not None else 0
    o.gzip_extra = extra
    o.gzip_extra_len = len(extra) if extra is not None else 0
    o.gzip_filename, o.gzip_comment = filename, comment
    return o


def _ptr_array(bufs):
    """array of pointers to the numpy uint8 buffers (kept alive by the caller)"""
    arr = (C.c_void_p * len(bufs))()
    for i, b in enumerate(bufs):
        arr[i] = b.ctypes.data if b.size else None
    return arr


def _as_u8(x):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.uint8)
    return np.frombuffer(bytes(x), dtype=np.uint8) if len(x) else np.zeros(0, dtype=np.uint8)


class Context:
    """b2f_ctx: one CUDA device, its stream and scratch memory (not thread-safe)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().b2f_ctx_create(device, C.byref(self._h))
        if rc != OK:
            raise B2fError(rc, (lib().b2f_last_error(None) or b"").decode())

    def close(self):
        if self._h:
            lib().b2f_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise B2fError(rc, (lib().b2f_last_error(self._h) or b"").decode())

    @property
    def handle(self):
        return self._h

    # ---- E2
    def lz77_default(self, data, window=32768, max_len=258):
        d = _as_u8(data)
        codes = np.empty(max(d.size, 1), dtype=np.uint32)
        n = C.c_size_t(0)
        self._check(lib().b2f_lz77_default(self._h, d.ctypes.data if d.size else None, d.size, window, max_len, codes.ctypes.data, C.byref(n)))
        return codes[: n.value].copy()

    # ---- batch encode: returns list of bytes
    def encode_batch(self, fmt, datas, schedules=None, **kw):
        L = lib()
        o = make_opts(**kw)
        ins = [_as_u8(d) for d in datas]
        n = len(ins)
        in_len = (C.c_size_t * n)(*[a.size for a in ins])
        sched_arrs, sched_ptrs, n_sched = [], (C.c_void_p * n)(), (C.c_size_t * n)()
        for i in range(n):
            sc = None if schedules is None else schedules[i]
            if sc is None:
                sched_ptrs[i] = None
                n_sched[i] = 0
            else:
                a = np.asarray(list(sc) + [0], dtype=np.int64)      # +1 so that an empty schedule still has an address
                sched_arrs.append(a)
                sched_ptrs[i] = a.ctypes.data
                n_sched[i] = len(sc)
        caps = [L.b2f_encode_bound(a.size, int(n_sched[i]), C.byref(o)) for i, a in enumerate(ins)]
        outs = [np.empty(c, dtype=np.uint8) for c in caps]
        out_cap = (C.c_size_t * n)(*caps)
        out_len = (C.c_size_t * n)()
        status = (C.c_int * n)()
        in_ptrs, out_ptrs = _ptr_array(ins), _ptr_array(outs)
        self._check(L.b2f_encode_batch(self._h, fmt, C.byref(o), n, in_ptrs, in_len, sched_ptrs if schedules is not None else None,
                                       n_sched if schedules is not None else None, out_ptrs, out_cap, out_len, status))
        res = []
        for i in range(n):
            if status[i] != OK:
                raise B2fError(status[i], f"stream {i}")
            res.append(outs[i][: out_len[i]].tobytes())
        return res

    def encode(self, fmt, data, schedule=None, **kw):
        return self.encode_batch(fmt, [data], None if schedule is None else [schedule], **kw)[0]

    # ---- batch decode: returns list of (status, bytes, in_consumed)
    def decode_batch(self, fmt, datas, caps=None):
        L = lib()
        ins = [_as_u8(d) for d in datas]
        n = len(ins)
        if caps is None:
            caps = [max(1 << 16, a.size * 1100 + 1024) for a in ins]
        outs = [np.empty(c, dtype=np.uint8) for c in caps]
        in_len = (C.c_size_t * n)(*[a.size for a in ins])
        out_cap = (C.c_size_t * n)(*caps)
        out_len, used, status = (C.c_size_t * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        in_ptrs, out_ptrs = _ptr_array(ins), _ptr_array(outs)
        self._check(L.b2f_decode_batch(self._h, fmt, n, in_ptrs, in_len, out_ptrs, out_cap, out_len, used, status))
        return [(status[i], outs[i][: min(out_len[i], caps[i])].tobytes(), used[i], out_len[i]) for i in range(n)]

    def decode(self, fmt, data, cap=None):
        return self.decode_batch(fmt, [data], None if cap is None else [cap])[0]

    # ---- device-resident variants (pointers are raw device addresses, e.g. torch tensor.data_ptr())
    def encode_device(self, fmt, d_in, in_off, in_len, d_out, out_off, out_cap, schedules=None, **kw):
        L = lib()
        o = make_opts(**kw)
        n = len(in_len)
        a_off = (C.c_uint64 * n)(*in_off); a_len = (C.c_size_t * n)(*in_len)
        o_off = (C.c_uint64 * n)(*out_off); o_cap = (C.c_size_t * n)(*out_cap)
        out_len, status = (C.c_size_t * n)(), (C.c_int * n)()
        sched_arrs, sched_ptrs, n_sched = [], (C.c_void_p * n)(), (C.c_size_t * n)()
        if schedules is not None:
            for i, sc in enumerate(schedules):
                if sc is None:
                    sched_ptrs[i] = None; n_sched[i] = 0
                else:
                    a = sc if isinstance(sc, np.ndarray) else np.asarray(list(sc) + [0], dtype=np.int64)
                    sched_arrs.append(a); sched_ptrs[i] = a.ctypes.data; n_sched[i] = len(sc)
        self._check(L.b2f_encode_device(self._h, fmt, C.byref(o), n, C.c_void_p(d_in), a_off, a_len,
                                        sched_ptrs if schedules is not None else None, n_sched if schedules is not None else None,
                                        C.c_void_p(d_out), o_off, o_cap, out_len, status))
        return list(out_len), list(status)

    def decode_device(self, fmt, d_in, in_off, in_len, d_out, out_off, out_cap):
        L = lib()
        n = len(in_len)
        a_off = (C.c_uint64 * n)(*in_off); a_len = (C.c_size_t * n)(*in_len)
        o_off = (C.c_uint64 * n)(*out_off); o_cap = (C.c_size_t * n)(*out_cap)
        out_len, used, status = (C.c_size_t * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        self._check(L.b2f_decode_device(self._h, fmt, n, C.c_void_p(d_in), a_off, a_len, C.c_void_p(d_out), o_off, o_cap, out_len, used, status))
        return list(out_len), list(used), list(status)

    # ---- one part of a stream (whole blocks), device resident: returns (bits, crc32, adler32)
    def encode_part_device(self, d_in, in_len, d_out, out_cap, schedule=None, is_last=False, **kw):
        o = make_opts(**kw)
        bits, crc, adler = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
        if schedule is not None:
            a = schedule if isinstance(schedule, np.ndarray) else np.asarray(list(schedule) + [0], dtype=np.int64)
            sp, ns = C.c_void_p(a.ctypes.data), len(schedule)
        else:
            sp, ns = None, 0
        self._check(lib().b2f_encode_part_device(self._h, C.byref(o), C.c_void_p(d_in), in_len, sp, ns, 1 if is_last else 0,
                                                 C.c_void_p(d_out), out_cap, C.byref(bits), C.byref(crc), C.byref(adler)))
        return bits.value, crc.value, adler.value

    def bits_shift_device(self, d_src, n_bits, shift, d_dst):
        self._check(lib().b2f_bits_shift_device(self._h, C.c_void_p(d_src), n_bits, shift, C.c_void_p(d_dst)))

    # ---- host-buffer calls on caller-owned numpy buffers (no copies in Python; used by bench.py's e2e leg)
    def encode_into(self, fmt, src, dst, schedule=None, **kw):
        L = lib()
        o = make_opts(**kw)
        in_len = (C.c_size_t * 1)(src.size); out_cap = (C.c_size_t * 1)(dst.size)
        out_len, status = (C.c_size_t * 1)(), (C.c_int * 1)()
        in_ptrs = (C.c_void_p * 1)(src.ctypes.data); out_ptrs = (C.c_void_p * 1)(dst.ctypes.data)
        if schedule is not None:
            a = schedule if isinstance(schedule, np.ndarray) else np.asarray(list(schedule) + [0], dtype=np.int64)
            sp = (C.c_void_p * 1)(a.ctypes.data); ns = (C.c_size_t * 1)(len(schedule))
        else:
            sp, ns = None, None
        self._check(L.b2f_encode_batch(self._h, fmt, C.byref(o), 1, in_ptrs, in_len, sp, ns, out_ptrs, out_cap, out_len, status))
        if status[0] != OK:
            raise B2fError(status[0], "encode_into")
        return out_len[0]

    def decode_into(self, fmt, src, src_len, dst):
        L = lib()
        in_len = (C.c_size_t * 1)(src_len); out_cap = (C.c_size_t * 1)(dst.size)
        out_len, used, status = (C.c_size_t * 1)(), (C.c_size_t * 1)(), (C.c_int * 1)()
        in_ptrs = (C.c_void_p * 1)(src.ctypes.data); out_ptrs = (C.c_void_p * 1)(dst.ctypes.data)
        self._check(L.b2f_decode_batch(self._h, fmt, 1, in_ptrs, in_len, out_ptrs, out_cap, out_len, used, status))
        return out_len[0], used[0], status[0]

    # ---- checksums
    def _cksum(self, fn, datas, init):
        ins = [_as_u8(d) for d in datas]
        n = len(ins)
        lens = (C.c_size_t * n)(*[a.size for a in ins])
        out = (C.c_uint32 * n)()
        ini = None if init is None else (C.c_uint32 * n)(*init)
        self._check(fn(self._h, n, _ptr_array(ins), lens, ini, out))
        return list(out)

    def crc32(self, datas, init=None):
        return self._cksum(lib().b2f_crc32_batch, datas, init)

    def adler32(self, datas, init=None):
        return self._cksum(lib().b2f_adler32_batch, datas, init)

    def set_overlap(self, on):
        self._check(lib().b2f_ctx_set_overlap(self._h, 1 if on else 0))

    # ---- stats
    def stats(self):
        s = Stats()
        lib().b2f_get_stats(self._h, C.byref(s))
        stages = [((lib().b2f_stage_name(self._h, i) or b"").decode(), s.last_kernel_ms[i]) for i in range(s.last_n_stages)]
        return {"kernel_launches": s.kernel_launches, "stages": stages, "device_ms": s.last_device_ms,
                "decode_parallel_streams": s.decode_parallel_streams, "decode_inorder_streams": s.decode_inorder_streams,
                "staged_h2d_bytes": s.staged_h2d_bytes, "staged_d2h_bytes": s.staged_d2h_bytes}


def host_alloc(nbytes):
    """page-locked numpy uint8 buffer from b2f_host_alloc (the DMA engines use it in place); free with host_free(arr)"""
    p = C.c_void_p()
    rc = lib().b2f_host_alloc(nbytes, C.byref(p))
    if rc != OK:
        raise B2fError(rc, "b2f_host_alloc")
    arr = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(p.value))[:nbytes]
    return arr


def host_free(arr):
    lib().b2f_host_free(C.c_void_p(arr.ctypes.data))


def plan_from_writes(sched, in_len, block_size=1 << 20, window=32768):
    L = lib()
    a = np.asarray(list(sched) + [0], dtype=np.int64) if sched is not None else None
    nc, nb = C.c_size_t(0), C.c_size_t(0)
    L.b2f_plan_from_writes(a.ctypes.data if a is not None else None, len(sched) if sched is not None else 0, in_len, block_size, window,
                           None, C.byref(nc), None, None, None, C.byref(nb))
    ce = np.zeros(max(nc.value, 1), dtype=np.uint64)
    be = np.zeros(max(nb.value, 1), dtype=np.uint64)
    bc = np.zeros(max(nb.value, 1), dtype=np.uint32)
    bf = np.zeros(max(nb.value, 1), dtype=np.uint8)
    L.b2f_plan_from_writes(a.ctypes.data if a is not None else None, len(sched) if sched is not None else 0, in_len, block_size, window,
                           ce.ctypes.data, C.byref(nc), be.ctypes.data, bc.ctypes.data, bf.ctypes.data, C.byref(nb))
    return ce[: nc.value].tolist(), be[: nb.value].tolist(), bc[: nb.value].tolist(), bf[: nb.value].tolist()
