"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs.  Bit-exact (integer / byte work); the only floating-point columns are formatted with
%.3f from host libm on both sides."""
import numpy as np
import pytest

import liblq

pytestmark = pytest.mark.gpu


def _L():
    import longqc_b200 as L
    return L


@pytest.mark.parametrize("w,k", [(5, 12), (10, 15), (5, 15), (3, 4), (7, 11), (32, 9), (10, 19), (5, 16), (11, 28)])
def test_sketch_adversarial(w, k):
    L = _L()
    rng = np.random.default_rng(100 + w * 31 + k)
    seqs = liblq.adversarial_seqs(rng, 210, 2500)
    rs = liblq.reads_from_seqs(seqs)
    x, y = L.sketch(rs, L.Opt(w=w, k=k), rid_base=7)
    want = liblq.oracle_sketch_set(rs, w, k, 0, rid_base=7)
    assert len(x) == len(want)
    assert np.array_equal(x, want["x"]) and np.array_equal(y, want["y"])


@pytest.mark.parametrize("w,k", [(5, 12), (10, 15)])
def test_sketch_tiled_kernel_where_rolling_is_default(w, k):
    """both exact formulations exist for LongQC's (w,k): the tiled closed form must agree too"""
    L = _L()
    rng = np.random.default_rng(77)
    rs = liblq.reads_from_seqs(liblq.adversarial_seqs(rng, 140, 2500))
    L.load().lqcov_debug_sketch_tiled(1)
    try:
        x, y = L.sketch(rs, L.Opt(w=w, k=k))
    finally:
        L.load().lqcov_debug_sketch_tiled(0)
    want = liblq.oracle_sketch_set(rs, w, k, 0)
    assert np.array_equal(x, want["x"]) and np.array_equal(y, want["y"])


@pytest.mark.parametrize("w,k,mode", [(5, 12, 0), (5, 15, 0), (5, 12, 2), (5, 15, 2), (5, 12, 3), (5, 15, 3)])
def test_sketch_packed_key_kernel(w, k, mode):
    """the 64-bases-per-thread packed-key kernel (lq_sketch_pk_core.h), fed by bulk copies (0) or plain loads (2), and the rolling
    kernel it replaced (3): long clean reads so that nearly every segment takes the unrolled blocks, repeat-rich reads for the
    twin records, the adversarial set for the segments the form declines, N-rich reads for tiles that run the general machine"""
    L = _L()
    rng = np.random.default_rng(300 + w + k)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = liblq.adversarial_seqs(rng, 120, 2500)
    for _ in range(40):
        seqs.append(bytes(rng.choice(acgt, size=int(rng.integers(500, 20000))).tobytes()))
    for _ in range(20):
        unit = bytes(rng.choice(acgt, size=int(rng.integers(1, 9))).tobytes())
        body = bytearray(rng.choice(acgt, size=6000).tobytes())
        for _ in range(12):
            p0 = int(rng.integers(0, 5800)); rep = unit * int(rng.integers(3, 40)); body[p0:p0 + len(rep)] = rep[:max(0, 6000 - p0)]
        seqs.append(bytes(body))
    rs = liblq.reads_from_seqs(seqs)
    L.load().lqcov_debug_sketch_tiled(mode)
    try:
        x, y = L.sketch(rs, L.Opt(w=w, k=k), rid_base=3)
    finally:
        L.load().lqcov_debug_sketch_tiled(0)
    want = liblq.oracle_sketch_set(rs, w, k, 0, rid_base=3)
    assert len(x) == len(want)
    assert np.array_equal(x, want["x"]) and np.array_equal(y, want["y"])


def test_sketch_hpc():
    L = _L()
    rng = np.random.default_rng(5)
    rs = liblq.reads_from_seqs(liblq.adversarial_seqs(rng, 140, 2500))
    x, y = L.sketch(rs, L.Opt(w=10, k=15, is_hpc=1))
    want = liblq.oracle_sketch_set(rs, 10, 15, 1)
    assert np.array_equal(x, want["x"]) and np.array_equal(y, want["y"])


def test_sketch_reads_10k():
    L = _L()
    from longqc_b200 import synth
    T, _ = synth.standard_set(300, 10000, 0.13, seed=1, n_query=1)
    x, y = L.sketch(T, L.Opt())
    want = liblq.oracle_sketch_set(T, 5, 12)
    assert np.array_equal(x, want["x"]) and np.array_equal(y, want["y"])


def _tandem_set(seed, n=300, L_=6000, nq=40):
    from longqc_b200 import synth
    rng = np.random.default_rng(seed)
    g = synth.make_genome(60000, rng)
    g = synth.add_tandem_repeats(g, rng, 40, unit_len=(2, 40), copies=(10, 150))
    T = synth.simulate_reads(g, n, L_, 0.10, rng)
    qi = np.sort(rng.choice(n, nq, replace=False))
    return T, T.subset(qi)


@pytest.mark.parametrize("kind", ["plain", "tandem"])
def test_seeds_and_sort(kind):
    """collect_seed_hits order and the exact radix_sort_128x permutation, query by query."""
    L = _L()
    from longqc_b200 import synth
    if kind == "plain":
        T, Q = synth.standard_set(400, 8000, 0.13, seed=7, n_query=12)
    else:
        T, Q = _tandem_set(11, nq=12)
    opt, oopt = liblq.opt_pair(min_score_med=80, min_score_good=160)
    TANDEM = np.uint64(1 << 42)
    with L.Coverage(opt) as c:
        c.set_queries(Q)
        c.index_part(T)
        mid = c.stats()["mid_occ"]
        for q in range(Q.n):
            tr = liblq.oracle_trace(T, Q, q, oopt)
            assert tr["mid_occ"] == mid
            ux, uy, sx, sy = c.debug_seeds(q)
            assert len(ux) == tr["n_seeds"]
            assert np.array_equal(ux, tr["unsorted"]["x"]) and np.array_equal(uy, tr["unsorted"]["y"] & ~TANDEM)
            assert np.array_equal(sx, tr["sorted"]["x"]), "sorted keys differ"
            assert np.array_equal(sy, tr["sorted"]["y"] & ~TANDEM), "tie order differs from radix_sort_128x"


CASES = {
    "plain": dict(gen=("std", 400, 8000, 0.13, 7, 60), opt=dict(min_score_med=80, min_score_good=160)),
    "ont": dict(gen=("std", 300, 8000, 0.15, 8, 50), opt=dict(min_score_med=160, min_score_good=160)),
    "tandem": dict(gen=("tandem", 11), opt=dict(min_score_med=80, min_score_good=160)),
    "multipart": dict(gen=("std", 400, 8000, 0.13, 7, 60), opt=dict(min_score_med=80, min_score_good=160, batch_size=500000)),
    "tandem_multipart": dict(gen=("tandem", 12), opt=dict(min_score_med=80, min_score_good=160, batch_size=300000)),
    "k15": dict(gen=("std", 300, 8000, 0.13, 9, 40), opt=dict(k=15, min_score_med=160, min_score_good=160)),
    "ava": dict(gen=("std", 300, 6000, 0.13, 10, 40), opt=dict(ava=1, min_score_med=80, min_score_good=160)),
    "hpc_filter": dict(gen=("std", 300, 6000, 0.13, 10, 40), opt=dict(is_hpc=1, k=15, w=10, min_coverage=1, filter=1)),
}


def _gen(spec):
    from longqc_b200 import synth
    if spec[0] == "std":
        _, n, L_, err, seed, nq = spec
        return synth.standard_set(n, L_, err, seed=seed, n_query=nq)
    return _tandem_set(spec[1])


@pytest.mark.parametrize("name", sorted(CASES))
def test_table_vs_oracle(name):
    L = _L()
    T, Q = _gen(CASES[name]["gen"])
    opt, oopt = liblq.opt_pair(**CASES[name]["opt"])
    want, mid, parts = liblq.oracle_table(T, Q, oopt)
    with L.Coverage(opt) as c:
        c.set_queries(Q)
        c.add_targets(T)
        got = c.table()
        st = c.stats()
    assert st["mid_occ"] == mid and st["n_parts"] == parts
    if got != want:
        g, w_ = got.split(b"\n"), want.split(b"\n")
        bad = [i for i in range(min(len(g), len(w_))) if g[i] != w_[i]]
        raise AssertionError("%d/%d rows differ, first: %r vs %r" % (len(bad), len(w_), g[bad[0]] if bad else None, w_[bad[0]] if bad else None))


def test_n_and_fasta_queries():
    L = _L()
    from longqc_b200 import synth
    rng = np.random.default_rng(13)
    T, Q = synth.standard_set(300, 5000, 0.10, seed=13, n_query=30)
    T = synth.sprinkle_n(T, 0.01, rng)
    Q = synth.sprinkle_n(Q, 0.005, rng)
    Q = synth.ReadSet(Q.seq, Q.seq_off, None, Q.names)
    opt, oopt = liblq.opt_pair(min_score_med=80, min_score_good=160)
    want, _, _ = liblq.oracle_table(T, Q, oopt)
    assert L.coverage_table(T, Q, opt) == want


def test_sdust_table():
    L = _L()
    rng = np.random.default_rng(21)
    seqs = [s for s in liblq.adversarial_seqs(rng, 280, 4000)]
    rs = liblq.reads_from_seqs(seqs, qual=True, rng=rng)
    got = L.sdust_table(rs)
    want = liblq.oracle_sdust_table(rs)
    assert got == want
    rs2 = liblq.reads_from_seqs(seqs[:50])
    assert L.sdust_table(rs2) == liblq.oracle_sdust_table(rs2)
    rs3 = liblq.reads_from_seqs(liblq.sdust_stale_seqs(rng), qual=True, rng=rng)   # repeats interrupted by N: the interval list at its bound
    assert L.sdust_table(rs3) == liblq.oracle_sdust_table(rs3)
    # long reads cut into segments by the kernel: random stretches, long repeats (the whole-read fallback) and stale-N repeats joined
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    stale = liblq.sdust_stale_seqs(rng)
    long_reads = []
    for i in range(60):
        parts = []
        for _ in range(int(rng.integers(2, 9))):
            kind = int(rng.integers(0, 4))
            if kind == 0: parts.append(acgt[rng.integers(0, 4, int(rng.integers(100, 3000)))].tobytes())
            elif kind == 1: parts.append(stale[int(rng.integers(0, len(stale)))])
            elif kind == 2: parts.append(np.tile(acgt[rng.integers(0, 4, int(rng.integers(1, 5)))], 2000)[: int(rng.integers(300, 4000))].tobytes())
            else: parts.append(b"N" * int(rng.integers(1, 70)))
        long_reads.append(b"".join(parts))
    long_reads += [b"A" * 511, b"A" * 512, b"A" * 513, b"ACGT" * 128, b"ACGT" * 256 + b"N", b"A" * 1024 + b"N" + b"A" * 1023]
    rs4 = liblq.reads_from_seqs(long_reads, qual=True, rng=rng)
    assert L.sdust_table(rs4) == liblq.oracle_sdust_table(rs4)
