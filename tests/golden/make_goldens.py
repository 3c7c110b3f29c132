#!/usr/bin/env python3
"""Extracts the golden vectors that sile/libflate's own tests and doctests hold for the DEFLATE hot path
into small fixtures (goldens.json + a few binary files).  Run in the authoring container, where the reference
is mounted read-only at /root/reference; the fixtures are committed because /root/reference does not exist
on the GPU box.  Only test DATA (byte vectors) is extracted -- no reference source code is copied.

    python tests/golden/make_goldens.py
"""
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def src(path):
    with open(os.path.join(REF, path), "r", encoding="utf-8") as f:
        return f.read()


def ints(text):
    text = text.replace("///", "   ")             # doc-comment prefix inside doctest arrays
    text = re.sub(r"//[^\n]*", "", text)          # strip line comments
    return [int(x.replace("_", "")) for x in re.findall(r"\b\d[\d_]*\b", text)]


def array_after(text, anchor, nth=0):
    """integers of the first [...] literal after the nth occurrence of `anchor`"""
    pos = -1
    for _ in range(nth + 1):
        pos = text.index(anchor, pos + 1)
    lb = text.index("[", pos + len(anchor))
    depth, i = 0, lb
    while True:
        if text[i] == "[":
            depth += 1
        elif text[i] == "]":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return ints(text[lb + 1:i])


def rust_bytestr(text, anchor):
    pos = text.index(anchor)
    s = text.index('b"', pos) + 2
    out, i = [], s
    while text[i] != '"':
        if text[i] == "\\":
            if text[i + 1] == "x":
                out.append(int(text[i + 2:i + 4], 16)); i += 4
            else:
                out.append({"n": 10, "r": 13, "t": 9, "0": 0, "\\": 92, '"': 34}[text[i + 1]]); i += 2
        else:
            out.append(ord(text[i])); i += 1
    return out


G = {}
enc = src("src/deflate/encode.rs")
G["deflate_hello_dynamic"] = {"src": "src/deflate/encode.rs:152-154", "plain": "Hello World!",
                              "bytes": array_after(enc, "assert_eq!(encoder.finish().into_result().unwrap(),", 0)}
G["deflate_hello_stored"] = {"src": "src/deflate/encode.rs:178-180", "plain": "Hello World!",
                             "bytes": array_after(enc, "assert_eq!(encoder.finish().into_result().unwrap(),", 1)}
zl = src("src/zlib.rs")
G["zlib_decode_works"] = {"src": "src/zlib.rs:708-710", "plain": "Hello World!",
                          "bytes": array_after(zl, "const DECODE_WORKS_TESTDATA: [u8; 20] =")}
G["zlib_raw_encode"] = {"src": "src/zlib.rs:750-753", "plain": "Hello World!",
                        "bytes": array_after(zl, "const RAW_ENCODE_WORKS_EXPECTED: [u8; 23] =")}
G["zlib_hello_default"] = {"src": "src/zlib.rs:547-549 (doctest)", "plain": "Hello World!",
                           "bytes": array_after(zl, "assert_eq!(encoder.finish().into_result().unwrap(),", 0)}
t27 = zl[zl.index("fn test_issues_27"):]
G["zlib_issue27_none"] = {"src": "src/zlib.rs:840-866", "writes": ["fooooooooooooooooo", "bar", "baz"],
                          "bytes": array_after(t27, "let expected = vec!", 0)}
G["zlib_issue27_sync"] = {"src": "src/zlib.rs:883-890", "writes": ["fooooooooooooooooo", "bar", "baz"],
                          "bytes": array_after(t27, "let expected = vec!", 1)}
t2 = zl[zl.index("fn test_issue_2"):zl.index("fn test_issues_16")]
G["zlib_issue2_inputs"] = {"src": "src/zlib.rs:778-796",
                           "inputs": [array_after(t2, "assert_encode_decode!(", k) for k in range(4)]}
t71 = zl[zl.index("fn issue71"):]
G["zlib_issue71"] = {"src": "src/zlib.rs:917-934", "encoded": array_after(t71, "let encoded_data ="),
                     "partial": array_after(t71, "let decoded_data =")}
G["zlib_issue_method0"] = {"src": "src/zlib.rs:938-943", "encoded": [0, 0]}
gz = src("src/gzip.rs")
G["gzip_stored_mtime123"] = {"src": "src/gzip.rs:800-802 (doctest)", "plain": "Hello World!", "mtime": 123,
                             "bytes": array_after(gz, "assert_eq!(encoder.finish().into_result().unwrap(),", 0)}
for k in (1, 2, 3):
    G[f"gzip_issue15_{k}"] = {"src": "src/gzip.rs:1230-1247", "encoded": rust_bytestr(gz, f"fn issue_15_{k}")}
dec = src("src/deflate/decode.rs")
G["deflate_fixed_hello"] = {"src": "src/deflate/decode.rs:28-33 (doctest)", "plain": "Hello World!",
                            "bytes": array_after(dec, "let encoded_data =")}
G["deflate_issue3_header"] = {"src": "src/deflate/decode.rs:176-190", "encoded": array_after(dec[dec.index("fn test_issues_3"):], "let input =")}
G["deflate_it_works_too_long"] = {"src": "src/deflate/decode.rs:194-212", "encoded": array_after(dec[dec.index("fn it_works"):], "let input =")}
G["deflate_issue64"] = {"src": "src/deflate/decode.rs:216-220", "encoded": rust_bytestr(dec, "fn test_issue_64")}
G["checksum_kat"] = {"src": "src/checksum.rs:45-56", "input": "abcde", "crc32": 0x8587D865, "adler32": 0x05C801F0}
G["lz77_issue21"] = {"src": "src/lz77.rs:16-31", "input": "aaaaa", "codes": [["L", 97], ["P", 4, 1]]}

with open(os.path.join(HERE, "goldens.json"), "w") as f:
    json.dump(G, f, indent=0, separators=(",", ":"))

td = src("src/deflate/test_data.rs")
issue52 = bytes(array_after(td, "pub const ISSUE_52_INPUT: [u8; 16_052] ="))
assert len(issue52) == 16052
open(os.path.join(HERE, "issue52_input.bin"), "wb").write(issue52)
for rel in ("data/issues_16/crash-1bb6d408475a5bd57247ee40f290830adfe2086e",
            "data/issues_16/crash-369e8509a0e76356f4549c292ceedee429cfe125",
            "data/issues_16/crash-e75959d935650306881140df7f6d1d73e33425cb",
            "data/noncompressed_block_offset_sync/offset", "data/noncompressed_block_offset_sync/offset.gz"):
    dst = os.path.join(HERE, os.path.basename(rel).replace("crash-", "issue16_crash-") + ("" if "." in os.path.basename(rel) or "crash" in rel else ".bin"))
    shutil.copyfile(os.path.join(REF, rel), dst)
    os.chmod(dst, 0o644)
print("wrote", len(G), "goldens")
