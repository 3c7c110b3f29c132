import os, sys, random, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B2F_DEBUG"] = "1"
from libflate_b200 import native, titles
from oracle import oracle as orc
ctx = native.Context(0)
def run(name, enc, plain):
    b = ctx.stats()
    st, out, used, _ = ctx.decode(native.FMT_DEFLATE, enc, cap=len(plain) + 64)
    a = ctx.stats()
    print(f"== {name}: status {st} ok {out == plain} parallel +{a['decode_parallel_streams']-b['decode_parallel_streams']} inorder +{a['decode_inorder_streams']-b['decode_inorder_streams']}", flush=True)
for mib in (16, 64):
    d = titles.generate(mib << 20, seed=42, workers=8)
    enc = ctx.encode(native.FMT_DEFLATE, d, [8192] * (d.size // 8192 + 1))
    run(f"titles{mib}", enc, d.tobytes())
rng = random.Random(41)
def _text(rng, n, nwords=300):
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz_") for _ in range(rng.randint(2, 9))) for _ in range(nwords)]
    b = bytearray()
    while len(b) < n:
        b += rng.choice(words) + b"\n"
    return bytes(b[:n])
text = _text(rng, 6 << 20, nwords=5000)
rnd = bytes(rng.getrandbits(8) for _ in range(1 << 20))
mixed = text[: 1 << 20] + rnd[: 300000] + b"\x00" * 700000 + (b"abc" * 100000) + text[1 << 20: 2 << 20] + bytes([7]) * 500000
run("text_A", orc.encode(0, text, [8192] * (len(text) // 8192 + 1)), text)
run("text_single_write", orc.encode(0, text), text)
run("text_small_blocks", orc.encode(0, text, block_size=100000), text)
run("mixed", orc.encode(0, mixed, [8192] * (len(mixed) // 8192 + 1)), mixed)
run("random", orc.encode(0, rnd * 3, [8192] * 400), rnd * 3)
