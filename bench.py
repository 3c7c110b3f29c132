#!/usr/bin/env python3
"""bench.py -- benchmarks of the B200 DEFLATE hot path, one JSON line per run.

Workloads (BASELINE.json `configs`; `--workload`, default = the one the metric is quoted on):
  config3         gzip encode + decode round trip of a 277 303 937-byte enwiki-titles-shaped text (README's exact size; synthetic,
                  seeded), mtime=0, written in 8 KiB writes (=> 1058 LZ77 chunks of 256 KiB, 264 one-MiB blocks + 1 final block),
                  one stream per GPU (weak scaling: streams are independent, NCCL only gathers the counters)        [headline]
  config2         raw DEFLATE encode of 64 x 4 MiB independent text streams, one write_all each, per GPU
  config4         zlib encode of 1024 x 1 MiB streams sharded across the N GPUs (strong scaling; Adler-32 path)
  config5         decode only: ONE gzip member of 8192 one-MiB blocks (8 GiB), block-parallel inflate, per GPU
  decode-foreign  decode only: zlib level-6 output (cross-block references) of the 277 MB text -- what the reference's own decode
                  benchmark inflates (flate_bench/src/main.rs:49-55)

    python bench.py --gpus 1 --steps 5 --warmup 3 [--workload W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # libflate's own CPU path (the C restatement in oracle/), host cores, same workload

value = whole-job GiB/s of uncompressed bytes with inputs resident in HBM; e2e = the same through the C ABI with page-locked HOST
buffers (H2D/D2H inside the timed region); e2e_pageable = the same with ordinary pageable buffers (config3).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

SIZE = 277_303_937          # README.md:60 (enwiki-latest-all-titles-in-ns0)
WRITE = 8192                # io::copy-style 8 KiB writes (schedule "A")
SEED = 42
METRIC = "GiB/s encode+decode, enwiki-titles-shaped 265 MiB, 1/2/4/8 B200 vs CPU ref"
GIB = float(1 << 30)
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
CACHE = os.path.join("/tmp", "b2f_bench_cache")

WORKLOADS = {
    "config3": "gzip encode+decode round trip, 277303937 B titles-shaped text per GPU, 8 KiB writes, mtime=0 (BASELINE config 3)",
    "config2": "raw deflate encode, 64 x 4 MiB independent titles-shaped streams per GPU, one write_all each (BASELINE config 2)",
    "config4": "zlib encode of 1024 x 1 MiB titles-shaped streams sharded across the GPUs, one write_all each, Adler-32 (BASELINE config 4)",
    "config5": "decode only: one gzip member of 8192 one-MiB dynamic blocks (8 GiB), block-parallel inflate, per GPU (BASELINE config 5)",
    "decode-foreign": "decode only: zlib level-6 stream (cross-block references) of the 277303937 B titles-shaped text, per GPU",
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """samples nvidia-smi while the timed region runs (B200_PROFILING.md 'clocks line')"""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload data (host side)
def sched_a(n):
    return np.full(n // WRITE + 1, WRITE, dtype=np.int64)


def gen_config3(rank):
    from libflate_b200 import titles
    return titles.generate(SIZE, seed=SEED + 1000 * rank, cache_dir=CACHE)


def gen_config2(rank):
    from libflate_b200 import titles
    big = titles.generate(256 << 20, seed=1000 + 1000 * rank, cache_dir=CACHE)
    return [big[i << 22:(i + 1) << 22] for i in range(64)]


def gen_config4(rank, world):
    """the same 1024 streams whatever the world size: stream i = MiB i of the 1 GiB seed-2000 text; rank r takes a contiguous range"""
    from libflate_b200 import titles
    per = 1024 // world
    lo = rank * per
    n = per << 20
    segs = [titles.segment(2000 + (lo >> 1) + k, titles.SEGMENT) for k in range((n + titles.SEGMENT - 1) // titles.SEGMENT)] if world > 1 else None
    if world == 1:
        big = titles.generate(1 << 30, seed=2000, cache_dir=CACHE)
    else:
        big = np.frombuffer(b"".join(segs), dtype=np.uint8)[:n]
    return [big[i << 20:(i + 1) << 20] for i in range(per)]


def gen_config5(rank, gib):
    from libflate_b200 import titles
    base = titles.generate(256 << 20, seed=5000 + 1000 * rank, cache_dir=CACHE)
    return np.tile(base, 4 * gib)


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def _cpu_job(args):
    """one worker: the CPU oracle on one piece of the workload (ctypes releases the GIL)"""
    from oracle import oracle as orc
    kind, fmt, data, sched = args
    if kind == "rt":
        enc = orc.encode(fmt, data, sched, mtime=0)
        rc, out, used, _ = orc.decode(fmt, enc, cap=len(data) + 64)
        assert rc == 0 and len(out) == len(data)
        return len(data)
    if kind == "enc":
        return len(orc.encode(fmt, data, sched, mtime=0)) and len(data)
    rc, out, used, _ = orc.decode(fmt, data, cap=sched)             # kind == "dec": data = compressed, sched = output size
    assert rc == 0 and len(out) == sched
    return sched


def cpu_jobs(workload, cores, sample_mib=None):
    """(jobs, uncompressed bytes per pass, description).  Same bytes as the GPU arm's rank 0; a single stream is cut into `cores`
    contiguous MiB-aligned ranges for the round trip (libflate is single-threaded per stream: this is the most it can use)."""
    from oracle import oracle as orc
    if workload == "config3":
        d = gen_config3(0)
        n = SIZE if sample_mib is None else min(SIZE, sample_mib << 20)
        per = ((n + cores - 1) // cores + (1 << 20) - 1) >> 20 << 20
        jobs = [("rt", orc.FMT_GZIP, d[o:min(n, o + per)].tobytes(), sched_a(min(n, o + per) - o).tolist()) for o in range(0, n, per)]
        return jobs, n, f"the same {n} B text (seed 42) cut into {len(jobs)} contiguous ranges, one gzip stream (8 KiB writes) per range and thread"
    if workload == "config2":
        st = gen_config2(0)
        jobs = [("enc", orc.FMT_DEFLATE, a.tobytes(), None) for a in st]
        return jobs, sum(a.size for a in st), "the same 64 x 4 MiB streams, one per thread at a time"
    if workload == "config4":
        st = gen_config4(0, 1)
        if sample_mib:
            st = st[:sample_mib]
        jobs = [("enc", orc.FMT_ZLIB, a.tobytes(), None) for a in st]
        return jobs, sum(a.size for a in st), f"{len(st)} of the same 1024 x 1 MiB streams, one per thread at a time"
    if workload == "config5":
        from libflate_b200 import titles
        base = titles.generate(256 << 20, seed=5000, cache_dir=CACHE)[: (sample_mib or 64) << 20]
        enc = orc.encode(orc.FMT_GZIP, base.tobytes(), sched_a(base.size).tolist(), mtime=0)
        return [("dec", orc.FMT_GZIP, enc, base.size)] * cores, base.size * cores, \
            f"{cores} copies (one per thread) of a {base.size >> 20} MiB prefix of the 8 GiB member: one stream is one thread in libflate"
    if workload == "decode-foreign":
        d = gen_config3(0)[: (sample_mib or 64) << 20]
        enc = zlib.compress(d.tobytes(), 6)
        return [("dec", orc.FMT_ZLIB, enc, d.size)] * cores, d.size * cores, \
            f"{cores} copies (one per thread) of the zlib-6 stream of the first {d.size >> 20} MiB: one stream is one thread in libflate"
    raise SystemExit(f"unknown workload {workload}")


def cpu_time(jobs, cores, steps, warmup):
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(cores) as ex:
        for _ in range(warmup):
            list(ex.map(_cpu_job, jobs))
        t0 = time.perf_counter()
        for _ in range(steps):
            list(ex.map(_cpu_job, jobs))
        return (time.perf_counter() - t0) / steps


def cpu_baseline_leg(workload, cores):
    """bounded sample for the GPU arm's `cpu_baseline` key: all cores and one core"""
    sample = {"config3": 16 * cores, "config2": None, "config4": 8 * cores, "config5": 32, "decode-foreign": 32}[workload]
    jobs, nbytes, desc = cpu_jobs(workload, cores, sample)
    dt = cpu_time(jobs, cores, 1, 1)
    j1 = jobs[:1]
    n1 = len(j1[0][2]) if j1[0][0] != "dec" else j1[0][3]
    dt1 = cpu_time(j1, 1, 1, 0)
    return {"value": round(nbytes / dt / GIB, 4), "unit": "GiB/s", "cores": cores, "kind": "port", "value_1core": round(n1 / dt1 / GIB, 4),
            "sample": desc + "; 1 timed pass after 1 warm-up; C restatement of libflate (oracle/); value_1core = the first piece alone on one core",
            "published_reference": "README.md:60-67: 32.5 MiB/s encode, 195 MiB/s decode, 27.9 MiB/s round trip (1 thread, unspecified hardware)"}


def run_reference(args, rank, emit):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    jobs, nbytes, desc = cpu_jobs(args.workload, cores, {"config5": 64, "decode-foreign": 64}.get(args.workload))
    dt = cpu_time(jobs, cores, args.steps, args.warmup)
    val = nbytes / dt / GIB
    j1 = jobs[:1]
    n1 = len(j1[0][2]) if j1[0][0] != "dec" else j1[0][3]
    v1 = n1 / cpu_time(j1, 1, 1, 0) / GIB
    emit({
        "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "GiB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "sample": desc,
                   "note": "libflate is single-threaded per stream; this arm runs one piece of the workload per host core"},
        "cpu_baseline": {"value": round(val, 4), "unit": "GiB/s", "cores": cores, "kind": "port", "value_1core": round(v1, 4),
                         "sample": desc + "; C restatement of libflate (oracle/); value_1core = the first piece alone on one core"},
        "e2e": {"value": round(val, 4), "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--gib", type=int, default=8, help="config5: GiB of uncompressed data in the member (default: the BASELINE 8)")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (debugging)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # stdout carries exactly ONE JSON line: libraries (NCCL's version banner, ...) that print to fd 1 are diverted to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, emit)
        return

    import torch
    import torch.distributed as dist
    from libflate_b200 import native, shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = native.Context(local)
    L = native.lib()
    W = args.workload

    def pinned(n):
        return torch.empty(n, dtype=torch.uint8).pin_memory()

    # ---------------- per-workload setup: step_device() -> list of stats dicts, step_e2e(), bytes per rank and step, checks
    pageable_step = None
    golden_note = None
    if W == "config3":
        data = gen_config3(rank)
        size = data.size
        sched = sched_a(size)
        bound = L.b2f_encode_bound(size, len(sched), None)
        d_in = torch.from_numpy(data).cuda()
        d_enc = torch.empty(bound + 256, dtype=torch.uint8, device="cuda")
        d_dec = torch.empty(size + 256, dtype=torch.uint8, device="cuda")
        box = {"enc_len": 0}

        def step_device():
            ol, st = ctx.encode_device(native.FMT_GZIP, d_in.data_ptr(), [0], [size], d_enc.data_ptr(), [0], [bound], [sched], mtime=0)
            assert st[0] == 0
            se = ctx.stats()
            box["enc_len"] = ol[0]
            dl, used, st = ctx.decode_device(native.FMT_GZIP, d_enc.data_ptr(), [0], [ol[0]], d_dec.data_ptr(), [0], [size + 64])
            assert st[0] == 0 and dl[0] == size and used[0] == ol[0]
            return [se, ctx.stats()]

        h_in, h_enc, h_dec = torch.from_numpy(data).pin_memory(), pinned(bound + 256), pinned(size + 256)
        n_in, n_enc, n_dec = h_in.numpy(), h_enc.numpy(), h_dec.numpy()

        def step_e2e():
            ol = ctx.encode_into(native.FMT_GZIP, n_in, n_enc, sched, mtime=0)
            dl, used, st = ctx.decode_into(native.FMT_GZIP, n_enc, ol, n_dec)
            assert st == 0 and dl == size

        p_enc, p_dec = np.empty(bound + 256, dtype=np.uint8), np.empty(size + 256, dtype=np.uint8)     # ordinary pageable memory

        def pageable_step():
            ol = ctx.encode_into(native.FMT_GZIP, data, p_enc, sched, mtime=0)
            dl, used, st = ctx.decode_into(native.FMT_GZIP, p_enc, ol, p_dec)
            assert st == 0 and dl == size

        def check():
            assert torch.equal(d_dec[:size], d_in), "round trip mismatch (device leg)"
            assert np.array_equal(n_dec[:size], n_in), "round trip mismatch (host leg)"
            assert np.array_equal(p_dec[:size], data), "round trip mismatch (pageable host leg)"
            if rank == 0:
                # every compressed byte equals the oracle's: tests/test_gpu_fullsize.py compares them one by one and pins this length + CRC
                gold = json.load(open(os.path.join(ROOT, "tests", "golden", "config_goldens.json")))["config3"]
                m = box["enc_len"]
                got = (m, zlib.crc32(d_enc[:m].cpu().numpy()))
                assert got == (gold["enc_len"], gold["enc_crc32"]), f"encoded stream differs from the oracle's golden: {got} vs {gold}"
                assert zlib.crc32(n_enc[:m]) == gold["enc_crc32"] and zlib.crc32(p_enc[:m]) == gold["enc_crc32"]
                return "encoded bytes: length and CRC-32 equal the oracle's (tests/golden/config_goldens.json)"
            return None

        units = size
        h2d, d2h = (lambda: size + box["enc_len"]), (lambda: size + box["enc_len"])
        extra_cfg = lambda: {"bytes_per_stream": size, "streams_per_gpu": 1, "compressed_bytes": int(box["enc_len"]), "ratio": round(box["enc_len"] / size, 4)}
        scaling = "weak"
        total_units = units * world
    elif W in ("config2", "config4"):
        fmt = native.FMT_DEFLATE if W == "config2" else native.FMT_ZLIB
        streams = gen_config2(rank) if W == "config2" else gen_config4(rank, world)
        ns = len(streams)
        lens = [a.size for a in streams]
        units = sum(lens)
        bounds = [L.b2f_encode_bound(n, 0, None) for n in lens]
        in_off = [int(x) for x in np.cumsum([0] + [(n + 255) & ~255 for n in lens])[:-1]]
        out_off = [int(x) for x in np.cumsum([0] + [(b + 255) & ~255 for b in bounds])[:-1]]
        d_in = torch.zeros(in_off[-1] + lens[-1] + 256, dtype=torch.uint8, device="cuda")
        for o, a in zip(in_off, streams):
            d_in[o:o + a.size] = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        d_enc = torch.empty(out_off[-1] + bounds[-1] + 256, dtype=torch.uint8, device="cuda")
        box = {"ol": None}

        def step_device():
            ol, st = ctx.encode_device(fmt, d_in.data_ptr(), in_off, lens, d_enc.data_ptr(), out_off, bounds)
            assert not any(st)
            box["ol"] = ol
            return [ctx.stats()]

        h_in = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in streams]
        h_out = [pinned(b) for b in bounds]
        import ctypes as C
        a_in = (C.c_void_p * ns)(*[t.data_ptr() for t in h_in]); a_len = (C.c_size_t * ns)(*lens)
        a_out = (C.c_void_p * ns)(*[t.data_ptr() for t in h_out]); a_cap = (C.c_size_t * ns)(*bounds)
        a_ol, a_st = (C.c_size_t * ns)(), (C.c_int * ns)()
        opts = native.make_opts()

        def step_e2e():
            rc = L.b2f_encode_batch(ctx.handle, fmt, C.byref(opts), ns, a_in, a_len, None, None, a_out, a_cap, a_ol, a_st)
            assert rc == 0 and not any(a_st)

        def check():
            # independent inflate + trailer of a few streams (every byte against the oracle: tests/test_gpu_fullsize.py)
            for i in (0, ns // 2, ns - 1):
                e = h_out[i].numpy()[:a_ol[i]].tobytes()
                assert e == d_enc[out_off[i]:out_off[i] + box["ol"][i]].cpu().numpy().tobytes()
                assert zlib.decompress(e, -15 if W == "config2" else 15) == streams[i].tobytes(), i
            return "3 streams re-inflated with zlib; device and host legs produce identical bytes"

        h2d, d2h = (lambda: units), (lambda: int(sum(a_ol)))
        extra_cfg = lambda: {"streams_per_gpu": ns, "bytes_per_stream": lens[0], "compressed_bytes": int(sum(box["ol"])), "ratio": round(sum(box["ol"]) / units, 4)}
        scaling = "weak" if W == "config2" else "strong"
        total_units = units * world
    else:                                                           # decode-only workloads
        if W == "config5":
            data = gen_config5(rank, args.gib)
            size = data.size
            sched = np.full(size // WRITE, WRITE, dtype=np.int64)
            enc_ctx = native.Context(local)                        # the single-call encode of 8 GiB needs ~12 B of scratch per byte: own context, freed afterwards
            enc = np.empty(L.b2f_encode_bound(size, len(sched), None), dtype=np.uint8)
            m = enc_ctx.encode_into(native.FMT_GZIP, data, enc, sched, mtime=0)
            enc_ctx.close()
            fmt = native.FMT_GZIP
        else:
            data = gen_config3(rank)
            size = data.size
            enc = np.frombuffer(zlib.compress(data.tobytes(), 6), dtype=np.uint8)
            m = enc.size
            fmt = native.FMT_ZLIB
        plain_crc = zlib.crc32(data)
        d_enc = torch.from_numpy(enc[:m].copy()).cuda()
        d_enc = torch.cat([d_enc, torch.zeros(64, dtype=torch.uint8, device="cuda")])
        d_dec = torch.empty(size + 256, dtype=torch.uint8, device="cuda")
        par0 = ctx.stats()["decode_parallel_streams"]

        def step_device():
            dl, used, st = ctx.decode_device(fmt, d_enc.data_ptr(), [0], [m], d_dec.data_ptr(), [0], [size + 64])
            assert st[0] == 0 and dl[0] == size and used[0] == m
            return [ctx.stats()]

        h_enc, h_dec = torch.from_numpy(enc[:m].copy()).pin_memory(), pinned(size + 256)
        n_enc, n_dec = h_enc.numpy(), h_dec.numpy()

        def step_e2e():
            dl, used, st = ctx.decode_into(fmt, n_enc, m, n_dec)
            assert st == 0 and dl == size

        def check():
            assert zlib.crc32(n_dec[:size]) == plain_crc, "decoded bytes differ from the plain text (host leg)"
            assert torch.equal(d_dec[:size].cpu(), torch.from_numpy(data)), "decoded bytes differ from the plain text (device leg)"
            assert ctx.stats()["decode_parallel_streams"] > par0, "the stream was decoded by the in-order kernel, not the parallel path"
            return "decoded bytes equal the plain text (CRC-32 and full compare); parallel path taken"

        units = size
        h2d, d2h = (lambda: m), (lambda: size)
        extra_cfg = lambda: {"bytes_per_stream": size, "streams_per_gpu": 1, "compressed_bytes": int(m), "ratio": round(m / size, 4)}
        scaling = "weak"
        total_units = units * world

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return shard.max_over_ranks(x, world, device="cuda")

    # warm-up + parity check of the step.  The clock sampler starts here: the timed regions last a few hundred ms, less than
    # nvidia-smi needs for its first line, so it runs from the warm-up (same kernels, same load) through the timed regions.
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    if pageable_step:
        pageable_step()
    golden_note = check()

    local = {}

    def timed(fn, key):
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        local[key] = (time.perf_counter() - t0) / args.steps          # this rank alone (gathered below)
        barrier()
        return max_over_ranks(time.perf_counter() - t0) / args.steps  # the job: slowest rank, barrier to barrier

    launches0 = ctx.stats()["kernel_launches"]
    t_dev = timed(step_device, "device")
    launches = (ctx.stats()["kernel_launches"] - launches0) // args.steps
    t_e2e = timed(step_e2e, "e2e")
    t_page = timed(pageable_step, "pageable") if pageable_step else None
    staged = ctx.stats()
    extra = 0
    while len(sampler.rows) < 3 and extra < 40:          # keep the same load on the GPU (untimed) until nvidia-smi has reported
        step_device(); extra += 1
    clocks = sampler.stop()

    # per-kernel durations for the roofline: the same step with the LZ77 slices serialised on the library's stream, so that
    # every kernel runs alone between two CUDA events (the timed regions above run with the slices overlapped)
    ctx.set_overlap(False)
    stage_acc, stage_cnt = {}, {}
    step_device()
    n_roof = 3
    for _ in range(n_roof):
        for st in step_device():                          # one stats record per library call (a step = an encode call + a decode call)
            for name, ms in st["stages"]:
                stage_acc[name] = stage_acc.get(name, 0.0) + ms
                stage_cnt[name] = stage_cnt.get(name, 0) + 1
    ctx.set_overlap(True)

    # the one collective of the path: every rank's byte / second counters (NCCL all-gather)
    counters = shard.gather_counters({"bytes": float(units), "seconds_device": local["device"], "seconds_e2e": local["e2e"]}, world, device="cuda")

    # ---------------- roofline of the dominant kernel (device time from CUDA events on the library's stream)
    # per library call that runs the stage: "checksum" runs in the encode call (over the input) and in the decode call (over the output),
    # each time over the same number of bytes -- its time is per call, not the sum of both
    stage_ms = {k: v / stage_cnt[k] for k, v in stage_acc.items() if k not in ("sync", "results", "clear", "h2d", "spec_retry", "lz_pipeline")}
    dom = max(stage_ms, key=stage_ms.get)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"
    ratio = extra_cfg()["ratio"]
    # algorithmic bytes per uncompressed byte of every kernel (DESIGN.md "Roofline accounting")
    alg_per_byte = {"lz_find": 1.0, "lz_fixup": 0.0, "lz_chain": 1.0, "lz_match": 1.0, "checksum": 1.0, "parse_emit": 1.0, "parse_exits": 1.0, "bitpack": ratio,
                    "find_blocks": ratio, "spec_parse": ratio, "spec_tokens": ratio, "lz_resolve": 1.0, "lz_subst": 1.0, "inflate_inorder": 1.0 + ratio,
                    "huff_build": 0.0, "tile_bits": 0.0, "scan": 0.0, "write_headers": 0.0, "framing": 0.0, "parse_stitch": 0.0, "spec_retry": 0.0,
                    "checksum_final": 0.0}

    def roof(name):
        alg = alg_per_byte.get(name, 0.0) * units
        ach = alg / (stage_ms[name] * 1e-3) / 1e9 if stage_ms[name] > 0 else 0.0
        return {"kernel": name, "ms": round(stage_ms[name], 4), "algorithmic_bytes": int(alg), "achieved": round(ach, 2), "frac": round(ach / peak, 5)}

    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")          # filled from an `ncu --set full` capture (see profiles/README.md)
    if os.path.exists(prof):
        per_byte = json.load(open(prof)).get("per_input_byte", {}).get(dom)       # measured DRAM bytes per uncompressed byte of that kernel
        traffic = int(per_byte * units) if per_byte is not None else None
    r = roof(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": r["achieved"], "peak": peak, "unit": "GB/s", "frac": r["frac"], "traffic": traffic,
                "peak_source": peak_src, "kernel_ms": r["ms"], "algorithmic_bytes_per_launch": r["algorithmic_bytes"],
                "timing": "CUDA events on the library stream, serialised pass (b2f_ctx_set_overlap(0)), mean of 3 steps, per library call that runs the stage",
                "all_kernels": [roof(k) for k in sorted(stage_ms, key=stage_ms.get, reverse=True)]}

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = {"value": None, "unit": "GiB/s", "cores": 0, "kind": "port", "sample": "skipped"} if args.skip_cpu else cpu_baseline_leg(W, cores)
        cfg = {"workload": WORKLOADS[W]}
        cfg.update(extra_cfg())
        cfg.update({"l2": "inputs and outputs are larger than the 126 MB L2; no explicit flush",
                    "parallelism": f"{world} GPU(s), independent streams, no data-path collective; NCCL all-gather of the per-rank counters + the timing all-reduce",
                    "check": golden_note})
        line = {
            "metric": METRIC, "value": round(total_units / t_dev / GIB, 4), "unit": "GiB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(t_dev * 1e3, 3), "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": round(total_units / t_e2e / GIB, 4), "unit": "GiB/s", "ms_per_step": round(t_e2e * 1e3, 3),
                    "h2d_bytes_per_step": int(h2d()), "d2h_bytes_per_step": int(d2h()),
                    "api": "C ABI batch calls on page-locked host buffers (b2f_host_alloc-style memory)"},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "per_rank": [{"bytes": int(c["bytes"]), "ms_device": round(c["seconds_device"] * 1e3, 3), "ms_e2e": round(c["seconds_e2e"] * 1e3, 3)} for c in counters],
        }
        if t_page is not None:
            line["e2e_pageable"] = {"value": round(total_units / t_page / GIB, 4), "unit": "GiB/s", "ms_per_step": round(t_page * 1e3, 3),
                                    "api": "the same calls on ordinary pageable numpy buffers: staged through the library's pinned buffers",
                                    "staged_h2d_bytes_total": int(staged["staged_h2d_bytes"]), "staged_d2h_bytes_total": int(staged["staged_d2h_bytes"])}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
