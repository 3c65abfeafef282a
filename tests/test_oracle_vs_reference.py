"""CPU tier, build container only: function-level pinning of the oracle against the compiled, unmodified
reference (oracle/_ref/libmm2ref.so).  Skipped where oracle/_ref was not built."""
import ctypes as C

import numpy as np
import pytest

import liblq

pytestmark = pytest.mark.skipif(liblq.ref() is None, reason="oracle/_ref/libmm2ref.so not built (needs /root/reference)")


@pytest.mark.parametrize("w,k,hpc", [(5, 12, 0), (10, 15, 0), (5, 15, 0), (3, 4, 0), (16, 20, 0), (5, 28, 0), (10, 15, 1), (5, 12, 1)])
def test_sketch(w, k, hpc):
    rng = np.random.default_rng(1000 + w + 31 * k + hpc)
    for s in liblq.adversarial_seqs(rng, 140, 2000):
        assert np.array_equal(liblq.ref_sketch(s, w, k, 3, hpc), liblq.oracle_sketch(s, w, k, 3, hpc))


def test_radix_sort_128x_tie_order():
    rng = np.random.default_rng(3)
    o, r = liblq.oracle(), liblq.ref()
    for trial in range(200):
        n = int(rng.integers(1, 4000)) if trial % 10 else int(rng.integers(1, 130))
        x = ((rng.integers(0, 2, n).astype(np.uint64) << np.uint64(63)) | (rng.integers(0, 300, n).astype(np.uint64) << np.uint64(32))
             | rng.integers(0, 500, n).astype(np.uint64))
        a = np.zeros(n, dtype=liblq.mm128_dtype)
        a["x"] = x
        a["y"] = np.arange(n, dtype=np.uint64)
        b = a.copy()
        o.lqo_radix_sort_128x(a.ctypes.data, a.ctypes.data + 16 * n)
        r.radix_sort_128x(b.ctypes.data, b.ctypes.data + 16 * n)
        assert np.array_equal(a["y"], b["y"])


def test_q2p_table_and_meanq():
    r = liblq.ref()
    tab = (C.c_double * 127).in_dll(r, "q2p")
    o = liblq.oracle()
    o.lqo_q2p.restype = C.c_double
    hc = liblq.hostcheck()
    hc.lqhc_q2p.restype = C.c_double
    for q in range(127):   # oracle and product both rebuild the reference's literal table
        assert o.lqo_q2p(q) == tab[q] and hc.lqhc_q2p(q) == tab[q]
    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(1, 5000))
        qs = (rng.integers(0, 60, n).astype(np.uint8) + 33).tobytes()
        assert r.meanQ(qs, n) == liblq.oracle().lqo_meanQ(qs, n)


def _ref_table(T, Q, flags):
    import os
    import subprocess
    import tempfile
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "minimap2-coverage")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/minimap2-coverage not built")
    with tempfile.TemporaryDirectory() as d:
        tf, qf = os.path.join(d, "t.fq"), os.path.join(d, "q.fq")
        T.write_fastx(tf)
        Q.write_fastx(qf)
        return subprocess.run([exe] + flags.split() + [tf, qf], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout


def test_table_ultralong_reads():
    """BASELINE configs[3] flavour (ONT ultra-long 50 kb reads, -x ont-rapid => -p 160): long chains, many anchors per target"""
    from longqc_b200 import synth
    T, Q = synth.standard_set(60, 50000, 0.15, seed=31, n_query=12)
    want = _ref_table(T, Q, "-Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 160 -t 4")
    got, _, parts = liblq.oracle_table(T, Q, liblq.oracle_opt(min_score_med=160, min_score_good=160))
    assert parts == 1 and got == want and got.count(b"\n") == 12


def test_table_many_parts_with_covt_gate():
    """a deep, tiny genome cut into many index parts: the COVT gate (esterr.c:87-91) closes for later parts"""
    from longqc_b200 import synth
    T, Q = synth.standard_set(600, 3000, 0.10, seed=32, coverage=300.0, n_query=25)   # ~300x of a 6 kb genome
    want = _ref_table(T, Q, "-Y -l 0 -q 160 -k 12 -w 5 -I 150K -p 80 -t 2")
    got, _, parts = liblq.oracle_table(T, Q, liblq.oracle_opt(min_score_med=80, min_score_good=160, batch_size=150000))
    assert parts >= 6 and got == want


def test_sdust_stale_window_against_reference_binary(tmp_path):
    """repeats interrupted by N (the interval list near (W-2)^2 entries): oracle table == `sdust` of the unmodified reference"""
    import os
    import subprocess
    if not os.path.exists(liblq.REF_SDUST):
        pytest.skip("oracle/_ref/sdust not built")
    rng = np.random.default_rng(77)
    rs = liblq.reads_from_seqs(liblq.sdust_stale_seqs(rng), qual=True, rng=rng)
    fq = str(tmp_path / "s.fq")
    rs.write_fastx(fq)
    want = subprocess.run([liblq.REF_SDUST, fq], capture_output=True, check=True).stdout
    assert liblq.oracle_sdust_table(rs) == want
