"""CPU tier: the C-ABI library loads and exports every symbol include/lqcov.h declares; host-side logic
(reader, part boundaries, option defaults, CLI argument checks, loud failure without a GPU)."""
import ctypes as C
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

import liblq

ROOT = liblq.ROOT


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_exports_every_declared_symbol():
    import longqc_b200 as L
    lib = L.load()
    hdr = open(os.path.join(ROOT, "include", "lqcov.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(lqcov_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(lib, n), "liblqcov.so does not export %s" % n
    assert lib.lqcov_abi_version() == 1


def test_option_defaults_match_reference_main():
    import longqc_b200 as L
    o = L.Opt()
    assert (o.k, o.w, o.is_hpc, o.batch_size, o.mini_batch_size) == (12, 5, 0, 4000000000, 50000000)
    assert (o.max_gap, o.min_cnt, o.min_chain_score, o.max_chain_skip, o.bw) == (10000, 3, 40, 25, 500)
    assert (o.max_overhang, o.min_ovlp, o.min_coverage, o.min_ratio) == (2000, 1000, 3, 0.4)
    assert abs(o.mid_occ_frac - 2e-4) < 1e-9 and o.no_self == 1 and o.ava == 0


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_fails_loudly_without_gpu():
    import longqc_b200 as L
    with pytest.raises(L.LqcovError):
        L.Coverage(L.Opt())
    from longqc_b200 import synth
    T, Q = synth.standard_set(20, 500, 0.1, seed=1, n_query=5)
    with pytest.raises(L.LqcovError):
        L.sketch(T)
    with pytest.raises(L.LqcovError):
        L.sdust_table(T)


def test_cli_argument_checks(tmp_path):
    import longqc_b200 as L
    exe = L.bin_path("minimap2-coverage")
    fq = tmp_path / "a.fq"
    fq.write_bytes(b"@r0\nACGTACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIIIIIII\n")
    run = lambda *a: subprocess.run([exe] + list(a), capture_output=True)
    p = run("-X", "-Y", str(fq), str(fq))
    assert p.returncode == 1 and b"mutually exclusive" in p.stderr and p.stdout == b""
    p = run(str(fq), str(fq))
    assert p.returncode == 1 and b"Choose either -X" in p.stderr
    p = run("-Y", "-m", "50", "-p", "40", str(fq), str(fq))
    assert p.returncode == 1 and b"-p must be larger" in p.stderr
    p = run("-Y", "-p", "80", "-q", "60", str(fq), str(fq))
    assert p.returncode == 1 and b"-q must be larger" in p.stderr
    p = run("-Y", str(fq))            # missing query: argp usage error
    assert p.returncode != 0 and p.stdout == b""
    p = run("--version")
    assert p.returncode == 0 and b"minimap2-coverage" in p.stdout
    if not _has_gpu():
        p = run("-Y", str(fq), str(fq))
        assert p.returncode == 1 and b"no usable CUDA device" in p.stderr and p.stdout == b""
    p = subprocess.run([L.bin_path("sdust")], capture_output=True)
    assert p.returncode == 1 and b"Usage: sdust" in p.stderr


def _read_all(path, chunk=0):
    import longqc_b200 as L
    from longqc_b200 import _lib
    lib = L.load()
    r = lib.lqcov_reader_open(path.encode())
    assert r
    out = []
    st = _lib.ReadsStruct()
    while lib.lqcov_reader_next(r, chunk, C.byref(st)) > 0:
        n = st.n
        so = np.ctypeslib.as_array(C.cast(st.seq_off, C.POINTER(C.c_uint64)), (n + 1,)).copy()
        no = np.ctypeslib.as_array(C.cast(st.name_off, C.POINTER(C.c_uint64)), (n + 1,)).copy()
        seq = C.string_at(st.seq, int(so[n]))
        names = C.string_at(st.names, int(no[n]))
        qual = C.string_at(st.qual, int(so[n])) if st.qual else None
        for i in range(n):
            out.append((names[int(no[i]):int(no[i + 1])], seq[int(so[i]):int(so[i + 1])], None if qual is None else qual[int(so[i]):int(so[i + 1])]))
        if chunk <= 0:
            break
    lib.lqcov_reader_close(r)
    return out


def test_reader_kseq_semantics(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_bytes(b"garbage before\n>s1 some comment\nACGT\nacgu\r\n\nNN\n>s2\n\n>s3\tx\nA\n")
    assert _read_all(str(fa)) == [(b"s1", b"ACGTacguNN", None), (b"s2", b"", None), (b"s3", b"A", None)]
    fq = tmp_path / "y.fq.gz"
    with gzip.open(fq, "wb") as f:
        f.write(b"@q1 c\nACGT\nAC\n+q1\nIIII\nII\n@q2\nGG\n+\n@@\n@q3\nACGT\n+\nII\n")   # q3: truncated quality ends the stream
    assert _read_all(str(fq)) == [(b"q1", b"ACGTAC", b"IIIIII"), (b"q2", b"GG", b"@@")]
    # chunked reading returns the same records in order
    from longqc_b200 import synth
    T, _ = synth.standard_set(50, 700, 0.05, seed=4, n_query=2)
    p = tmp_path / "t.fq"
    T.write_fastx(str(p))
    whole = _read_all(str(p))
    parts = _read_all(str(p), chunk=5000)
    assert whole == parts and len(whole) == 50
    assert [w[0] for w in whole] == T.names


def test_part_boundaries_follow_the_minibatch_rule():
    from longqc_b200.dist import part_boundaries, split_even
    lens = np.full(100, 1000)
    assert part_boundaries(lens, 25000, 5000) == [(0, 30), (30, 60), (60, 90), (90, 100)]
    assert part_boundaries(lens, 4000000000) == [(0, 100)]
    assert part_boundaries(np.array([10, 10, 10]), 5, 50000000) == [(0, 1), (1, 2), (2, 3)]
    assert split_even(10, 3) == [(0, 4), (4, 7), (7, 10)]
    # agrees with the oracle's own part count on a real case
    import cases
    T, Q = cases.make_case("parts")
    _, _, parts = liblq.oracle_table(T, Q, liblq.oracle_opt(**cases.opts("parts")))
    assert len(part_boundaries(T.lengths(), 400000)) == parts


def test_host_table_code_against_oracle_rows():
    """lq_table.c (filter_redundant / reliable_region / row formatting) through the test-only host build"""
    hc = liblq.hostcheck()
    if not hasattr(hc, "lqhc_rows_selftest"):
        pytest.skip("host table self-test not built")
    assert hc.lqhc_rows_selftest() == 0
