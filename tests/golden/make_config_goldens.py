"""Generates tests/golden/config_goldens.json: length and CRC-32 of the ORACLE's encoding of the headline workload
(BASELINE config 3: 277 303 937 B titles-shaped text, seed 42, gzip, mtime 0, 8 KiB writes).  bench.py compares the GPU
encoder's output with these two numbers in its warm-up; tests/test_gpu_fullsize.py compares every byte with the oracle AND
checks that the oracle still reproduces this file.  Run: python tests/golden/make_config_goldens.py   (about 20 s, CPU only)."""
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libflate_b200 import titles
from oracle import oracle as orc

SIZE, SEED, WRITE = 277_303_937, 42, 8192
t = time.time()
d = titles.generate(SIZE, seed=SEED)
sched = [WRITE] * (SIZE // WRITE + 1)
enc = orc.encode(orc.FMT_GZIP, d.tobytes(), sched, mtime=0)
out = {"config3": {"size": SIZE, "seed": SEED, "write": WRITE, "plain_crc32": zlib.crc32(d), "enc_len": len(enc), "enc_crc32": zlib.crc32(enc),
                   "oracle_encode_seconds": round(time.time() - t, 1)}}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config_goldens.json"), "w"), indent=1)
print(out)
