"""ONE stream encoded by several contexts / GPUs: the host logic around b2f_encode_part_device.

libflate's DEFLATE blocks are independent, so contiguous runs of whole blocks ("parts") are encoded separately; the only data the
parts exchange is 8 bytes each (their length in bits -> exclusive scan -> bit position) plus their checksums (combined
algebraically).  `plan_parts` cuts a write schedule on block boundaries, `assemble` places the shifted parts and appends the
trailer.  The result is byte-identical to the single-call encode (tests/test_gpu_split.py; host arithmetic: tests/test_split_host.py).
"""
import ctypes as C

import numpy as np

from . import _native as nv


def block_cuts(sched, in_len, block_size=1 << 20):
    """(byte offset, schedule index) after every write / flush that ends a DEFLATE block -- the bookkeeping of Block::write
    (src/deflate/encode.rs:277-286): a block ends when the bytes written since the last block end reach block_size, or at flush()"""
    cuts, pos, orig = [], 0, 0
    for k, w in enumerate(sched):
        if w < 0:
            cuts.append((pos, k + 1)); orig = 0
            continue
        w = min(int(w), in_len - pos)
        pos += w; orig += w
        if orig >= block_size:
            cuts.append((pos, k + 1)); orig = 0
    return cuts


def plan_parts(sched, in_len, nparts, block_size=1 << 20):
    """cuts the schedule into <= nparts runs of whole blocks, balanced by bytes: [(byte_lo, byte_hi, sched_lo, sched_hi)]"""
    cuts = block_cuts(sched, in_len, block_size)
    parts, lo_b, lo_s = [], 0, 0
    for p in range(1, nparts):
        want = in_len * p // nparts
        best = min((c for c in cuts if c[0] > lo_b and c[1] < len(sched) + 1), key=lambda c: abs(c[0] - want), default=None)
        if best is None or best[0] >= in_len and best[1] >= len(sched):
            break
        if best[0] <= lo_b:
            continue
        parts.append((lo_b, best[0], lo_s, best[1]))
        lo_b, lo_s = best
    parts.append((lo_b, in_len, lo_s, len(sched)))
    return parts


def or_bits(dst, bit_off, src, n_bits):
    """dst |= src << bit_off for a src that was already shifted by (bit_off & 7) on its GPU: whole bytes, OR at the seams"""
    nb = ((bit_off & 7) + n_bits + 7) // 8
    b0 = bit_off >> 3
    dst[b0:b0 + nb] |= src[:nb]


def assemble(fmt, opts, parts, total_len):
    """parts: list of (shifted bytes as numpy uint8 (shift = position & 7), n_bits, crc32, adler32, in_len) in stream order"""
    L = nv.lib()
    hdr = np.zeros(1 << 17, dtype=np.uint8)
    hl = L.b2f_stream_header(fmt, C.byref(opts), hdr.ctypes.data, hdr.size)
    total_bits = sum(p[1] for p in parts)
    out = np.zeros(hl + (total_bits + 7) // 8 + 8, dtype=np.uint8)
    out[:hl] = hdr[:hl]
    pos, crc, adler = 8 * hl, 0, 1
    for data, n_bits, c, a, n in parts:
        or_bits(out, pos, data, n_bits)
        pos += n_bits
        crc = L.b2f_crc32_combine(crc, c, n)
        adler = L.b2f_adler32_combine(adler, a, n)
    end = (pos + 7) // 8
    tl = L.b2f_stream_trailer(fmt, crc, adler, total_len, out[end:].ctypes.data)
    return out[:end + tl]


def encode_split(ctxs, fmt, data, sched, **kw):
    """encodes ONE stream with len(ctxs) contexts (one per GPU in a real deployment; any contexts in a test), part p on ctxs[p]"""
    import torch
    data = np.ascontiguousarray(data, dtype=np.uint8)
    sched = list(sched)
    opts = nv.make_opts(**kw)
    pl = plan_parts(sched, data.size, len(ctxs), kw.get("block_size", 1 << 20))
    enc = []
    for p, (b0, b1, s0, s1) in enumerate(pl):
        d_in = torch.from_numpy(data[b0:b1].copy()).cuda() if b1 > b0 else torch.zeros(16, dtype=torch.uint8, device="cuda")
        cap = nv.lib().b2f_encode_bound(b1 - b0, s1 - s0, C.byref(opts)) + 64
        d_out = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        bits, crc, adler = ctxs[p].encode_part_device(d_in.data_ptr(), b1 - b0, d_out.data_ptr(), cap, sched[s0:s1], is_last=p + 1 == len(pl), **kw)
        enc.append((d_out, bits, crc, adler, b1 - b0))
    # the one exchange: every part's bit length (8 bytes) -> exclusive scan
    hl = nv.lib().b2f_stream_header(fmt, C.byref(opts), None, 0)
    pos, parts = 8 * hl, []
    for p, (d_out, bits, crc, adler, n) in enumerate(enc):
        d_sh = torch.zeros((bits + 7) // 8 + 16, dtype=torch.uint8, device="cuda")
        ctxs[p].bits_shift_device(d_out.data_ptr(), bits, pos & 7, d_sh.data_ptr())
        parts.append((d_sh.cpu().numpy(), bits, crc, adler, n))
        pos += bits
    return assemble(fmt, opts, parts, data.size).tobytes()
