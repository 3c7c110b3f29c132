"""Step-by-step CPU model of k_lz_find (libflate_b200/csrc/encode_kernels.cu): ring / mirror index arithmetic, the loader's
staging schedule and throttle under the most adversarial interleaving (loader and ordered sections run as far ahead of the
oldest unfinished block as the kernel allows), the hop cap and the cooperative long-match extension -- compared with a
brute-force matcher that follows libflate_lz77/src/default.rs:69-129.  The constants mirror the kernel's (kFindCap,
kFindRing, kFindMirror, kFindHops, kFindLane, kSeg); positions the kernel defers to k_lz_fixup take the brute-force value.
This is how the kernel's design was checked before its first GPU run; it does not execute any product code."""
import numpy as np

from libflate_b200 import titles

CAP = 64
LOOK = 32768; SEG = 262144; HB = 14; DEPTH = 8; MIRROR = 384
RING = LOOK + 128 * CAP; RB = RING // 128

def brute(buf, window, max_len):
    n = len(buf); end = max(3, n) - 3
    md = np.zeros(n, dtype=np.uint32); last = {}
    bb = bytes(buf)
    for i in range(end):
        key = bb[i:i+3]; j = last.get(key, -1); last[key] = i
        if j >= 0 and i - j <= window:
            L = 3
            while L < max_len and i + L < n and bb[i+L] == bb[j+L]: L += 1
            md[i] = (L << 16) | (i - j)
    return md

def ld_in32(mem, off):
    v = 0
    for k in range(4):
        if off + k < len(mem): v |= int(mem[off + k]) << (8 * k)
    return v

def emu_segment(mem, cd_off, n, seg, window, max_len, md, adversarial, REF):
    end = max(3, n) - 3
    s_start = seg * SEG; s_end = min(s_start + SEG, n); lim = min(s_end, end)
    ws = s_start - LOOK if s_start > LOOK else 0
    a0 = ws - ((cd_off + ws) & 3); g0 = cd_off + a0
    nblk = (s_end - a0 + 127) >> 7
    nstage = nblk + 3
    head = np.zeros(1 << HB, dtype=np.int64)
    lring = np.full(RING, 0xDEAD, dtype=np.int64)
    bring = np.full(RING + MIRROR, 0xEE, dtype=np.int64)
    def st32(idx, v):
        for k in range(4): bring[idx + k] = (v >> (8 * k)) & 255
    def ld32u(idx):
        w = idx & ~3; v = 0
        for k in range(8): v |= int(bring[w + k]) << (8 * k)
        return int((v >> (8 * (idx & 3))) & 0xFFFFFFFF)
    st = {"staged": 0, "lrb": 0}
    def stage_step():
        x = st["staged"]
        for k in range(DEPTH):
            rb = st["lrb"]
            for lane in range(32):
                v = ld_in32(mem, g0 + 128 * (x + k) + 4 * lane); st32(rb + 4 * lane, v)
                if rb < MIRROR: st32(RING + rb + 4 * lane, v)
            rb += 128
            if rb >= RING: rb = 0
            st["lrb"] = rb
        st["staged"] = x + DEPTH
    state = {}
    def ordered(b):
        rb = (b % RB) * 128; bpos = a0 + 128 * b
        t = {}; dd = {}
        for s4 in range(4):
            for lane in range(32):
                t[(s4, lane)] = ld32u(rb + 32 * s4 + lane) & 0xFFFFFF
        for s4 in range(4):
            for lane in range(32):
                pos = bpos + 32 * s4 + lane
                vv = ws <= pos < lim
                d = 0
                if vv:
                    h = ((t[(s4, lane)] * 0x9E3779B1) & 0xFFFFFFFF) >> (32 - HB)
                    o = head[h]; head[h] = max(head[h], pos + 1)
                    d = pos + 1 - o if o else 0
                    if d > LOOK: d = 0
                dd[(s4, lane)] = d
                lring[rb + 32 * s4 + lane] = d
        state[b] = (t, dd)
    def walk(b):
        rb = (b % RB) * 128; bpos = a0 + 128 * b
        t, dd = state.pop(b)
        if bpos + 127 < s_start: return
        HOPS = 16; LANE = 16
        for s4 in range(4):
            P = [bpos + 32 * s4 + l for l in range(32)]
            valid = [s_start <= p < s_end for p in P]
            if not any(valid): continue
            total = [0]*32; found=[False]*32; deferred=[False]*32; jxs=[0]*32; k=[0]*32; limit=[0]*32; opn=[False]*32; A=[0]*32; S=[0]*32
            for lane in range(32):
                ri = rb + 32 * s4 + lane; tt = t[(s4, lane)]; d = dd[(s4, lane)] if valid[lane] else 0
                jx = ri; hops = 0
                while d:
                    total[lane] += d
                    if total[lane] > window: break
                    jx -= d
                    if jx < 0: jx += RING
                    dn = int(lring[jx]); tj = ld32u(jx) & 0xFFFFFF
                    if tj == tt: found[lane] = True; break
                    d = dn
                    hops += 1
                    if hops >= HOPS and d: deferred[lane] = True; break
                jxs[lane] = jx
                limit[lane] = min(max_len - 3, n - (P[lane] + 3)) if found[lane] else 0
                A[lane] = ri + 3; S[lane] = jx + 3
                opn[lane] = found[lane] and limit[lane] != 0
                for it in range(LANE // 4):
                    if opn[lane]:
                        x = ld32u(A[lane] + k[lane]) ^ ld32u(S[lane] + k[lane])
                        if x: k[lane] += ((x & -x).bit_length() - 1) >> 3; opn[lane] = False
                        else:
                            k[lane] += 4
                            if k[lane] >= limit[lane]: opn[lane] = False
            U = sum(1 << l for l in range(32) if opn[l])
            if U:
                Fm = 0
                for l in range(1, 32):
                    if opn[l] and (U >> (l - 1)) & 1 and total[l - 1] == total[l]: Fm |= 1 << l
                H = U & ~Fm; ext = [0]*32
                hm = H
                while hm:
                    h = (hm & -hm).bit_length() - 1; hm &= hm - 1
                    fr = (~((Fm >> h) >> 1)) & 0xFFFFFFFF
                    run = 0 if h == 31 else ((fr & -fr).bit_length() - 1 if fr else 31 - h)
                    lim_ext = min(max_len - 3 + run, n - (P[h] + 3))
                    e = LANE
                    while True:
                        xs = []
                        for lane in range(32):
                            o = e + 4 * lane
                            xs.append((ld32u(A[h] + o) ^ ld32u(S[h] + o)) if o < lim_ext else 1)
                        mm = [l for l in range(32) if xs[l]]
                        if mm:
                            l0 = mm[0]; x0 = xs[l0]; o0 = e + 4 * l0
                            eh = lim_ext if o0 >= lim_ext else min(lim_ext, o0 + (((x0 & -x0).bit_length() - 1) >> 3))
                            break
                        e += 128
                    ext[h] = eh
                for lane in range(32):
                    if opn[lane]:
                        below = H & (0xFFFFFFFF >> (31 - lane))
                        hl = below.bit_length() - 1
                        k[lane] = ext[hl] - (lane - hl)
                        assert k[lane] >= 0
            for lane in range(32):
                if not valid[lane]: continue
                if deferred[lane]:
                    md[P[lane]] = REF[P[lane]]       # k_lz_fixup: plain walk from HBM (not modelled here)
                    continue
                out = 0
                if found[lane]:
                    kk = min(k[lane], limit[lane]); out = ((3 + kk) << 16) | total[lane]
                md[P[lane]] = out
    nord = 0; nwalk = 0          # next block to run its ordered section / oldest unwalked block (= min cur_blk)
    while nwalk < nblk:
        progress = True
        while progress:
            progress = False
            # loader: runs ahead as far as the throttle allows
            while st["staged"] < nstage:
                last = st["staged"] + DEPTH
                if last > CAP and not (nwalk >= last - CAP): break
                stage_step(); progress = True
            # ordered sections: as far ahead as staging allows (adversarial) or just one
            while nord < nblk and st["staged"] >= nord + 4 and (adversarial or nord == nwalk):
                ordered(nord); nord += 1; progress = True
        assert nord > nwalk, "deadlock in the emulated schedule"
        walk(nwalk); nwalk += 1

def run(n, cd_off, window=32768, max_len=258, adversarial=True, seed=3):
    d = titles.segment(seed, max(n, 4096))[: n]
    d = np.frombuffer(d, dtype=np.uint8)
    rng = np.random.default_rng(seed)
    d = d.copy()
    if n > 60000:
        d[1000:1400] = 97; d[50000:50600] = d[20000:20600]
        for k0 in range(3000, 3900): d[k0] = 65 + (k0 % 7)
        d[9000:9300] = d[8000:8300]; d[9305:9600] = d[8000:8295]
        d[n-400:n] = 66
    mem = np.concatenate([rng.integers(0, 256, cd_off, dtype=np.uint8), d, rng.integers(97, 100, 300, dtype=np.uint8)])
    md = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    nseg = (n + SEG - 1) // SEG
    ref = brute(d, window, max_len)
    for seg in range(nseg): emu_segment(mem, cd_off, n, seg, window, max_len, md, adversarial, ref)
    bad = np.nonzero(md != ref)[0]
    return len(bad) == 0



def test_model_tiny_inputs():
    assert run(2, 1) and run(5, 3) and run(300, 0) and run(1100, 2)


def test_model_small_window_and_max_length():
    assert run(40001, 2, window=1000, max_len=50)


def test_model_ring_wrap_adversarial_schedule():
    assert run(61000, 5, adversarial=True)          # > kFindRing positions: the rings wrap, runs and long repeats included


def test_model_in_order_schedule():
    assert run(30000, 0, adversarial=False)
