"""Command-line mirror of the reference's examples/flate.rs (:14-111): same subcommands and -i/-o/-v options, running on the
B200 path.  `python -m libflate_b200.flate -i in -o out gzip-encode`.

Like the example, the encoders are fed by an `io::copy` loop, i.e. 8 KiB writes, so the output bytes equal what
`cargo run --example flate -- gzip-encode` produces for the same input (gzip mtime: the reference stamps now(); pass --mtime to
reproduce a given file)."""
import argparse
import sys
import time

from . import _native as nv
from . import gzip, zlib

COPY_BUF = 8192        # std::io::copy's buffer → the write schedule of the example


def _open_in(name):
    return sys.stdin.buffer if name == "-" else open(name, "rb")


def _open_out(name):
    return sys.stdout.buffer if name == "-" else open(name, "wb")


def _copy_into(src, enc):
    while True:
        b = src.read(COPY_BUF)
        if not b:
            return
        enc.write(b)


def _describe_header(cmd, d):
    """Debug print of the container header, the analogue of `{:?}` on gzip::Header / zlib::Header (display only)."""
    if cmd.startswith("zlib"):
        if len(d) < 2:
            return "<truncated>"
        return f"zlib(window_bits={(d[0] >> 4) + 8}, level={d[1] >> 6}, dict={bool(d[1] & 0x20)})"
    if len(d) < 10:
        return "<truncated>"
    flg, pos, f = d[3], 10, {}
    f["mtime"], f["xfl"], f["os"] = int.from_bytes(d[4:8], "little"), d[8], d[9]
    f["is_text"] = bool(flg & 1)
    if flg & 4 and pos + 2 <= len(d):
        n = int.from_bytes(d[pos:pos + 2], "little")
        f["extra"] = d[pos + 2:pos + 2 + n].hex()
        pos += 2 + n
    for bit, key in ((8, "filename"), (16, "comment")):
        if flg & bit:
            e = d.find(b"\0", pos)
            e = len(d) if e < 0 else e
            f[key] = d[pos:e].decode("latin-1")
            pos = e + 1
    f["is_verified"] = bool(flg & 2)
    return "gzip(" + ", ".join(f"{k}={v!r}" for k, v in f.items()) + ")"


def main(argv=None):
    ap = argparse.ArgumentParser(prog="flate")
    ap.add_argument("-i", "--input", default="-")
    ap.add_argument("-o", "--output", default="-")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--mtime", type=int, default=None, help="gzip MTIME (default: now, as HeaderBuilder::new does)")
    sub = ap.add_subparsers(dest="command", required=True)
    sub.add_parser("copy")
    br = sub.add_parser("byte-read")
    br.add_argument("-u", "--unit", type=int, default=1)
    for c in ("gzip-decode", "gzip-decode-multi", "gzip-encode", "zlib-decode", "zlib-encode"):
        sub.add_parser(c)
    args = ap.parse_args(argv)
    src = _open_in(args.input)
    cmd = args.command
    if cmd == "byte-read":
        count = 0
        while True:
            b = src.read(args.unit)
            if not b:
                break
            count += len(b)
        print(f"COUNT: {count}")
        return 0
    dst = _open_out(args.output)
    if cmd == "copy":
        while True:
            b = src.read(COPY_BUF)
            if not b:
                break
            dst.write(b)
    else:
        ctx = nv.Context(args.device)
        if cmd in ("gzip-encode", "zlib-encode"):
            if cmd == "gzip-encode":
                enc = gzip.Encoder(ctx, mtime=int(time.time()) if args.mtime is None else args.mtime)
            else:
                enc = zlib.Encoder(ctx)
            _copy_into(src, enc)
            dst.write(enc.finish())
        else:
            cls = {"gzip-decode": gzip.Decoder, "gzip-decode-multi": gzip.MultiDecoder, "zlib-decode": zlib.Decoder}[cmd]
            data = src.read()
            if args.verbose and cmd != "gzip-decode-multi":
                print(f"HEADER: {_describe_header(cmd, data)}", file=sys.stderr)
            dec = cls(ctx, data)
            while True:
                b = dec.read(1 << 22)
                if not b:
                    break
                dst.write(b)
    if dst is not sys.stdout.buffer:
        dst.close()
    else:
        dst.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
