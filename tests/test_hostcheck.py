"""CPU tier: the host/device core headers the CUDA kernels are built from (lq_*_core.h), compiled for the
host (liblqcov_hostcheck.so, test-only) and checked against the oracle.  These are the exactness
arguments of DESIGN.md exercised without a GPU."""
import ctypes as C

import numpy as np
import pytest

import liblq


@pytest.mark.parametrize("w,k", [(5, 12), (10, 15), (5, 15), (3, 4), (1, 6), (7, 11), (16, 20), (5, 28), (32, 12)])
def test_position_parallel_sketch(w, k):
    rng = np.random.default_rng(5 + w * 100 + k)
    for s in liblq.adversarial_seqs(rng, 105, 1500):
        want = liblq.oracle_sketch(s, w, k, 3)
        assert np.array_equal(liblq.hc_sketch("lqhc_sketch_parallel", s, w, k, 3), want)
        assert np.array_equal(liblq.hc_sketch("lqhc_sketch_replay", s, w, k, 3, 0), want)
        if k <= 16:   # the kernel's windowed closed form (palindromes / N inside the look-back handled without replay)
            assert np.array_equal(liblq.hc_sketch("lqhc_sketch_parallel_win", s, w, k, 3), want)


@pytest.mark.parametrize("w,k", [(5, 12), (5, 15), (5, 11), (10, 15), (10, 12), (3, 8), (2, 4), (7, 13)])
def test_packed_key_sketch(w, k):
    """64 bases per thread, candidates as single integers (lq_sketch_pk_core.h): what lq_sketch_pk_k runs.  Every segment the
    form accepts comes from it (unrolled blocks, the general blocks at read starts / ends / after equal k-mers), the others from
    the reference state machine; the concatenation must be the oracle's sketch."""
    rng = np.random.default_rng(900 + w * 100 + k)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = liblq.adversarial_seqs(rng, 210, 2500)
    for _ in range(12):    # long clean reads: the common case
        seqs.append(bytes(rng.choice(acgt, size=int(rng.integers(300, 6000))).tobytes()))
    for _ in range(20):    # palindrome- and repeat-rich blocks inside random sequence: twins, look-backs that decline
        unit = bytes(rng.choice(np.frombuffer(b"AT", dtype=np.uint8), size=int(rng.integers(1, 9))).tobytes())
        body = bytearray(rng.choice(acgt, size=1500).tobytes())
        for _ in range(6):
            p0 = int(rng.integers(0, 1400)); rep = unit * int(rng.integers(3, 40)); body[p0:p0 + len(rep)] = rep[:max(0, 1500 - p0)]
        seqs.append(bytes(body))
    hc = liblq.hostcheck()
    lean_total = seg_total = 0
    for i, s in enumerate(seqs):
        want = liblq.oracle_sketch(s, w, k, 3)
        cap = 2 * len(s) + 64
        out = np.zeros(cap, dtype=liblq.mm128_dtype)
        n_lean = C.c_int(0)
        n = hc.lqhc_sketch_pk(s, len(s), w, k, 3, C.byref(n_lean), out.ctypes.data, cap)
        assert 0 <= n <= cap
        assert np.array_equal(out[:n], want)
        lean_total += n_lean.value; seg_total += (len(s) + 63) // 64
        if 210 <= i < 222 and k >= 11:   # plain random reads: the general state machine is for the odd segment (on the GPU it costs a warp ~4x)
            assert n_lean.value >= (len(s) + 63) // 64 - 1
    assert lean_total > 0.7 * seg_total   # the form itself is what is being tested


@pytest.mark.parametrize("w,k", [(5, 12), (10, 15), (3, 4)])
def test_bounded_replay_everywhere(w, k):
    rng = np.random.default_rng(6)
    for s in liblq.adversarial_seqs(rng, 70, 700):
        assert np.array_equal(liblq.hc_sketch("lqhc_sketch_slow_everywhere", s, w, k, 1), liblq.oracle_sketch(s, w, k, 1))


def test_hpc_replay():
    rng = np.random.default_rng(7)
    for s in liblq.adversarial_seqs(rng, 140, 1500):
        assert np.array_equal(liblq.hc_sketch("lqhc_sketch_replay", s, 10, 15, 3, 1), liblq.oracle_sketch(s, 10, 15, 3, 1))


def test_hash32_equals_hash64():
    hc, o = liblq.hostcheck(), liblq.oracle()
    rng = np.random.default_rng(8)
    for k in (4, 8, 12, 15, 16):
        mask = (1 << (2 * k)) - 1
        for key in rng.integers(0, mask + 1, 500):
            assert hc.lqhc_hash32(int(key), mask) == o.lqo_hash64(int(key), mask) == hc.lqhc_hash64(int(key), mask)


def _keys(rng, n, mode):
    u = np.uint64
    if mode == 0:
        return rng.integers(0, 1 << 62, n, dtype=np.uint64)
    if mode == 1:
        return (rng.integers(0, 2, n).astype(u) << u(63)) | (rng.integers(0, 300, n).astype(u) << u(32)) | rng.integers(0, 2000, n).astype(u)
    if mode == 2:
        return (rng.integers(0, 2, n).astype(u) << u(63)) | (rng.integers(0, 1 << 25, n).astype(u) << u(32)) | rng.integers(0, 50, n).astype(u)
    if mode == 3:
        return rng.integers(0, 5, n).astype(u) * u(0x0101010101010101)
    if mode == 5:   # few regions with small digit values on several levels (rid >> 16 in 0..5, rid & 255 in 0..11, rpos >> 8 in 0..9): many ties
        return ((rng.integers(0, 2, n).astype(u) << u(63)) | (rng.integers(0, 6, n).astype(u) << u(48)) | (rng.integers(0, 12, n).astype(u) << u(32))
                | rng.integers(0, 2560, n).astype(u))
    return ((rng.integers(0, 2, n).astype(u) << u(63)) | (rng.integers(0, 2, n).astype(u) << u(48)) | (rng.integers(0, 2, n).astype(u) << u(33))
            | rng.integers(0, 70000, n).astype(u))


# bit 0: two-region closed form, bit 1: cached-digit walk, bit 2: packed one-load-per-step walk, bit 3: digit-stream walk + expansion (device, long
# buckets), bit 4: few-region digit-stream walk (device, all digits < 16)
@pytest.mark.parametrize("use_two", [0, 1, 2, 3, 4, 5, 8, 9, 24, 25])
def test_seed_sort_walk_and_closed_form(use_two):
    """the queue-walk formulation (and the two-region closed form) reproduce ksort.h's permutation, ties included"""
    hc, o = liblq.hostcheck(), liblq.oracle()
    hc.lqhc_afsort.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
    rng = np.random.default_rng(30 + use_two)
    for trial in range(250):
        n = int(rng.integers(1, 5000)) if trial % 10 else int(rng.integers(1, 130))
        x = _keys(rng, n, trial % 6)
        a = np.zeros(n, dtype=liblq.mm128_dtype)
        a["x"] = x
        a["y"] = np.arange(n, dtype=np.uint64)
        o.lqo_radix_sort_128x(a.ctypes.data, a.ctypes.data + 16 * n)
        idx = np.zeros(n, dtype=np.uint32)
        hc.lqhc_afsort(x.ctypes.data, n, idx.ctypes.data, use_two)
        assert np.array_equal(idx.astype(np.uint64), a["y"])


def test_chain_chunked_equals_sequential():
    """32 predecessors per step with speculative stamping == chain.c's sequential inner loop (max_skip, t[] stamps)"""
    hc = liblq.hostcheck()
    args = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_float] + [C.c_void_p] * 3
    hc.lqhc_chain_seq.argtypes = args
    hc.lqhc_chain_chunked.argtypes = args
    rng = np.random.default_rng(9)
    for trial in range(200):
        n = int(rng.integers(1, 1500))
        mode = trial % 3
        if mode == 0:
            r = np.sort(rng.integers(0, 12000, n)).astype(np.uint32)
            q = (r.astype(np.int64) + rng.integers(-30, 30, n) + 500).astype(np.int32)
        elif mode == 1:
            r = np.sort(rng.integers(0, 300, n)).astype(np.uint32)
            q = rng.integers(0, 400, n).astype(np.int32)
        else:
            r = np.sort(rng.integers(0, 30000, n)).astype(np.uint32)
            q = rng.integers(0, 30000, n).astype(np.int32)
        sp = np.full(n, 12, dtype=np.uint8)
        outs = []
        for fn in (hc.lqhc_chain_seq, hc.lqhc_chain_chunked):
            f, p, v = (np.zeros(n, np.int32) for _ in range(3))
            fn(r.ctypes.data, q.ctypes.data, sp.ctypes.data, n, 10000, 500, 25, np.float32(12.0), f.ctypes.data, p.ctypes.data, v.ctypes.data)
            outs.append((f, p, v))
        assert all(np.array_equal(a, b) for a, b in zip(*outs))


def test_sdust_core():
    hc = liblq.hostcheck()
    hc.lqhc_sdust_masked.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    hc.lqhc_sdust_masked.restype = C.c_long
    rng = np.random.default_rng(21)
    seqs = liblq.adversarial_seqs(rng, 210, 3000) + liblq.sdust_stale_seqs(rng)   # the latter fill the interval list to ~(W-2)^2
    rs = liblq.reads_from_seqs(seqs)
    want = [int(ln.split(b"\t")[1]) for ln in liblq.oracle_sdust_table(rs).strip().split(b"\n")]
    got = [hc.lqhc_sdust_masked(s, len(s), 20, 64) for s in seqs]
    assert got == want
    # the segment form of the device kernel: every 96 / 256 / 1000 bases scanned on their own after 3 W bases of warm-up
    hc.lqhc_sdust_segments.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    hc.lqhc_sdust_segments.restype = C.c_long
    for S in (64, 96, 256, 1000):
        assert [hc.lqhc_sdust_segments(s, len(s), 20, 64, S, 4096, 512) for s in seqs] == want, S
