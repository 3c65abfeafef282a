"""`-d FILE` / index files as targets (index.c:390-479, longQC.py --db).
CPU tier: the writer of lq_mmi.cpp (khash slot order restated) against the file the UNMODIFIED reference writes -- byte for byte,
several parts, k = 12 / 15 and -H; the reader against the same files.  Skipped where oracle/_ref is absent, except for the committed
golden dump (tests/golden/dump_tiny.mmi.gz, made by oracle/make_golden.py) which also pins the format on the GPU box.
GPU tier: the drop-in executable writes the reference's bytes with -d, maps against the reference's file, and prints the reference's
table when the index was built with another k than the command line's (the rows' `n` is the command line's)."""
import ctypes as C
import gzip
import hashlib
import os
import subprocess

import numpy as np
import pytest

import liblq

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "minimap2-coverage")


def tiny_set(seed=5, n=40, L=1500, nq=12):
    from longqc_b200 import synth
    return synth.standard_set(n, L, 0.10, seed=seed, coverage=15.0, n_query=nq)


def _our_dump(path, T, w, k, hpc, batch):
    """the writer on records sketched by the oracle, cut into parts like the reference cuts them"""
    from longqc_b200 import _lib
    from longqc_b200.dist import part_boundaries
    hc = liblq.hostcheck()
    hc.lqhc_mmi_dump.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    for pi, (s, e) in enumerate(part_boundaries(T.lengths(), batch, 50_000_000)):
        part = T.subset(range(s, e))
        rec = liblq.oracle_sketch_set(part, w, k, hpc)
        key = np.ascontiguousarray((rec["x"] >> np.uint64(8)).astype(np.uint32))
        y = np.ascontiguousarray(rec["y"])
        keep = _lib.reads_struct(part)
        assert hc.lqhc_mmi_dump(path.encode(), 1 if pi else 0, w, k, hpc, C.byref(keep.st), key.ctypes.data, y.ctypes.data, len(key)) == 0


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("k,w,hpc,batch,flag", [(12, 5, 0, 4_000_000_000, "-I 4G"), (12, 5, 0, 20000, "-I 20K"), (13, 5, 0, 4_000_000_000, ""), (13, 10, 1, 4_000_000_000, "-H"),
                                             (9, 3, 0, 30000, "-I 30K"), (6, 5, 0, 4_000_000_000, "")])   # (k = 15, with and without -H, checked too: 2 minutes of host loops)
def test_dump_bytes_equal_reference(k, w, hpc, batch, flag, tmp_path):
    T, Q = tiny_set(seed=5 + k)
    tf, qf = str(tmp_path / "t.fq"), str(tmp_path / "q.fq")
    T.write_fastx(tf); Q.write_fastx(qf)
    ref_dump, our_dump = str(tmp_path / "ref.mmi"), str(tmp_path / "our.mmi")
    # (the reference needs a query file to get as far as writing the index)
    subprocess.run([REF, "-Y", "-k", str(k), "-w", str(w)] + flag.split() + ["-d", ref_dump, tf, qf], check=True, capture_output=True)
    _our_dump(our_dump, T, w, k, hpc, batch)
    a, b = open(ref_dump, "rb").read(), open(our_dump, "rb").read()
    assert len(a) == len(b) and a == b
    hc = liblq.hostcheck()
    hc.lqhc_mmi_load_count.restype = C.c_long
    hc.lqhc_mmi_load_count.argtypes = [C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    nr, ns = C.c_long(), C.c_long()
    parts = hc.lqhc_mmi_load_count(ref_dump.encode(), C.byref(nr), C.byref(ns))
    from longqc_b200.dist import part_boundaries
    assert parts == len(part_boundaries(T.lengths(), batch, 50_000_000)) and ns.value == T.n
    assert nr.value == len(liblq.oracle_sketch_set(T, w, k, hpc)) if parts == 1 else nr.value > 0


def test_committed_golden_dump():
    """the reference's own dump of the tiny set (committed): our writer reproduces it without /root/reference"""
    T, _ = tiny_set()
    want = gzip.open(os.path.join(GOLD, "dump_tiny.mmi.gz"), "rb").read()
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "our.mmi")
        _our_dump(p, T, 5, 12, 0, 20000)
        assert open(p, "rb").read() == want


def _cli(args, **kw):
    import longqc_b200 as L
    return subprocess.run([L.bin_path("minimap2-coverage")] + args, capture_output=True, **kw)


@pytest.mark.gpu
def test_executable_dump_and_map_like_longqc_db(tmp_path):
    """longQC.py --db: `-k 12 -w 5 -I 20K -d db targets`, then `-Y -l 0 -q 160 -p 80 -t 4 db queries`"""
    T, Q = tiny_set()
    tf, qf, db = str(tmp_path / "t.fq"), str(tmp_path / "q.fq"), str(tmp_path / "db.mmi")
    T.write_fastx(tf); Q.write_fastx(qf)
    p = _cli(["-k", "12", "-w", "5", "-I", "20K", "-d", db, tf])
    assert p.returncode == 0 and p.stdout == b"", p.stderr.decode()[-1000:]
    assert open(db, "rb").read() == gzip.open(os.path.join(GOLD, "dump_tiny.mmi.gz"), "rb").read()
    ref_db = str(tmp_path / "ref.mmi")
    open(ref_db, "wb").write(gzip.open(os.path.join(GOLD, "dump_tiny.mmi.gz"), "rb").read())
    p = _cli("-Y -l 0 -q 160 -p 80 -t 4".split() + [ref_db, qf])
    assert p.returncode == 0, p.stderr.decode()[-1000:]
    assert p.stdout == open(os.path.join(GOLD, "dump_tiny.map.tsv"), "rb").read()


@pytest.mark.gpu
def test_executable_maps_against_index_with_other_parameters(tmp_path):
    """index built with -w 10, mapped with the command line's defaults (k = 12, w = 5): the index's parameters win for the mapping, the
    command line's for the rows' minimizer count n (minimap2-coverage.c:418-427, 552-563)"""
    T, Q = tiny_set()
    tf, qf, db = str(tmp_path / "t.fq"), str(tmp_path / "q.fq"), str(tmp_path / "db10.mmi")
    T.write_fastx(tf); Q.write_fastx(qf)
    assert _cli(["-k", "12", "-w", "10", "-d", db, tf]).returncode == 0
    assert hashlib.md5(open(db, "rb").read()).hexdigest() == open(os.path.join(GOLD, "dump_tiny_w10.md5")).read().strip()
    p = _cli("-Y -l 0 -q 160 -p 80 -t 4".split() + [db, qf])
    assert p.returncode == 0, p.stderr.decode()[-1000:]
    assert b"overridden by parameters used in the prebuilt index" in p.stderr
    assert p.stdout == open(os.path.join(GOLD, "dump_tiny_w10.map.tsv"), "rb").read()
