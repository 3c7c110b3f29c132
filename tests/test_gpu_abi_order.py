"""The Rust shim (rust/libflate_b200) cannot be compiled here (no toolchain), so its exact call order against the C ABI --
Lz77Encode::encode/flush buffering with one b2f_lz77_default per chunk, Encoder::write/flush/finish with one b2f_encode_batch,
Decoder with the OUTPUT_TOO_SMALL retry -- is driven from C++ (tests/native/abi_order.cpp) and checked against the oracle."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_abi_order_program_links_against_the_header():
    """CPU: the program compiles and links against include/b2f.h + libb2f.so (every symbol the shim binds exists)"""
    from libflate_b200 import build
    from oracle import oracle as orc
    build.build()
    orc.build()
    subprocess.check_call(["make", "-C", os.path.join(HERE, "native"), "-s", "abi_order"])
    assert os.path.exists(os.path.join(HERE, "native", "_build", "abi_order"))


@pytest.mark.gpu
def test_abi_order_matches_oracle():
    test_abi_order_program_links_against_the_header()
    r = subprocess.run([os.path.join(HERE, "native", "_build", "abi_order")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi order ok" in r.stdout and r.stdout.count("0 mismatching chunks") == 3
