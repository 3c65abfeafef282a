"""CPU tier: the multi-threaded reader (lq_ingest.c) against the reference's OWN kseq/bseq reader (oracle/_ref/libmm2ref.so,
mm_bseq_open / mm_bseq_read, bseq.c:68-102) on hostile text: FASTA and FASTQ mixed, multi-line records, CR line ends, empty lines,
'@' and '>' inside qualities, truncated records, files whose size is a multiple of kseq's 16 KB buffer.  Every file is read three
ways -- tiny scan blocks (so that almost every block start is a wrong guess somewhere), default blocks, the sequential reader -- and
all three must deliver exactly the reference's records.  Where oracle/_ref is absent (GPU box) the three are compared with each other."""
import ctypes as C
import os

import numpy as np
import pytest

import liblq
from test_abi_and_host import _read_all


class Bseq1(C.Structure):
    _fields_ = [("l_seq", C.c_int), ("rid", C.c_int), ("name", C.c_char_p), ("seq", C.c_void_p), ("qual", C.c_void_p)]


def _ref_read(path, chunk=0x7fffffff, first_only=False):
    r = liblq.ref()
    r.mm_bseq_open.restype = C.c_void_p
    r.mm_bseq_open.argtypes = [C.c_char_p]
    r.mm_bseq_read.restype = C.POINTER(Bseq1)
    r.mm_bseq_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    r.mm_bseq_close.argtypes = [C.c_void_p]
    fp = r.mm_bseq_open(path.encode())
    assert fp
    out, n = [], C.c_int(0)
    while True:
        a = r.mm_bseq_read(fp, chunk, 1, C.byref(n))
        if n.value == 0:
            break
        for i in range(n.value):
            s = a[i]
            out.append((s.name, C.string_at(s.seq, s.l_seq), C.string_at(s.qual, s.l_seq) if s.qual else None))
        if first_only:
            break
    r.mm_bseq_close(fp)
    return out


def _ours_parts(path, mode):
    """records through the index-part rule with a huge part: the reference's mm_bseq_read loop (_ref_read) ends at its first empty
    batch, ours at the first call that does not return 1"""
    import longqc_b200 as L
    from longqc_b200 import _lib
    lib = L.load()
    env = {"tiny": {"LQCOV_READER_BLOCK": "97"}, "sequential": {"LQCOV_SEQUENTIAL_READER": "1"}}[mode]
    os.environ.update(env)
    try:
        r = lib.lqcov_reader_open(path.encode())
        st = _lib.ReadsStruct()
        out = []
        while lib.lqcov_reader_next(r, 0x7fffffff, C.byref(st)) == 1:
            n = st.n
            so = np.ctypeslib.as_array(C.cast(st.seq_off, C.POINTER(C.c_uint64)), (n + 1,)).copy()
            no = np.ctypeslib.as_array(C.cast(st.name_off, C.POINTER(C.c_uint64)), (n + 1,)).copy()
            seq = C.string_at(st.seq, int(so[n])); names = C.string_at(st.names, int(no[n]))
            qual = C.string_at(st.qual, int(so[n])) if st.qual else None
            for i in range(n):
                out.append((names[int(no[i]):int(no[i + 1])], seq[int(so[i]):int(so[i + 1])].replace(b"U", b"T").replace(b"u", b"t"),
                            None if qual is None else qual[int(so[i]):int(so[i + 1])]))
        lib.lqcov_reader_close(r)
    finally:
        for k in env:
            os.environ.pop(k, None)
    return out


def _ours(path, mode):
    env = {"tiny": {"LQCOV_READER_BLOCK": "97"}, "small": {"LQCOV_READER_BLOCK": "4096"}, "default": {}, "sequential": {"LQCOV_SEQUENTIAL_READER": "1"}}[mode]
    old = {k: os.environ.get(k) for k in ("LQCOV_READER_BLOCK", "LQCOV_SEQUENTIAL_READER")}
    for k in old:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        recs = _read_all(path)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    # bseq.c:61-63 turns U/u into T/t after kseq; our packer maps both to 3, the reader keeps the bytes
    return [(n, s.replace(b"U", b"T").replace(b"u", b"t"), None if (q is None or q == b"\0" * len(q) and len(q) > 0 and False) else q) for n, s, q in recs]


def _norm(recs, any_qual):
    """our reader hands out ONE quality blob per record set: records without qualities are zero-filled when any record has them"""
    out = []
    for n, s, q in recs:
        if any_qual and q is None:
            q = b"\0" * len(s)
        if not any_qual:
            q = None
        out.append((n, s, q))
    return out


def _hostile(rng, fastq_bias):
    acgt = b"ACGTNacgtnUu"
    parts = []
    for _ in range(int(rng.integers(1, 40))):
        L = int(rng.integers(0, 300))
        seq = bytes(acgt[i] for i in rng.integers(0, len(acgt), L))
        name = b"r%d" % int(rng.integers(0, 1000))
        sep = [b" ", b"\t", b"", b" cmt x", b"\r"][int(rng.integers(0, 5))]
        eol = b"\r\n" if rng.random() < 0.15 else b"\n"
        width = int(rng.integers(1, 120)) if rng.random() < 0.4 else 0
        lines = [seq[i:i + width] for i in range(0, L, width)] if width and L else [seq]
        if rng.random() < 0.1:
            lines.insert(int(rng.integers(0, len(lines) + 1)), b"")
        body = eol.join(lines) + eol
        if rng.random() < fastq_bias:
            qa = b"!I5@>+~#"
            q = bytes(qa[i] for i in rng.integers(0, len(qa), L))
            if rng.random() < 0.04:
                q = q[: max(0, L - int(rng.integers(1, 4)))]              # truncated: ends the stream
            if rng.random() < 0.03:
                q = q + b"II"                                              # too long
            ql = [q[i:i + width] for i in range(0, len(q), width)] if width and len(q) else [q]
            parts.append(b"@" + name + sep + eol + body + b"+" + (name if rng.random() < 0.3 else b"") + eol + eol.join(ql) + (eol if rng.random() < 0.95 else b""))
        else:
            parts.append(b">" + name + sep + eol + body)
        if rng.random() < 0.05:
            parts.append(b"junk line\n")
    txt = b"".join(parts)
    if rng.random() < 0.1:
        txt = txt[: int(rng.integers(0, len(txt) + 1))]                    # cut anywhere
    return txt


@pytest.mark.parametrize("seed", range(6))
def test_reader_equals_reference_kseq_on_hostile_text(seed, tmp_path):
    rng = np.random.default_rng(900 + seed)
    have_ref = liblq.ref() is not None
    for t in range(120):
        txt = _hostile(rng, fastq_bias=[0.0, 1.0, 0.5][t % 3])
        if t % 7 == 0 and len(txt) < 16384:                                # kseq's end-of-file knowledge: size a multiple of its buffer
            pad = b"\n" * 0 + b"\r" if t % 14 == 0 else b""
            txt = txt + pad
            txt = txt + b"A" * ((16384 - len(txt) % 16384) % 16384) if t % 14 else txt
        p = str(tmp_path / ("f%d.txt" % t))
        open(p, "wb").write(txt)
        got = {m: _ours(p, m) for m in ("tiny", "small", "default", "sequential")}
        any_qual = any(q is not None for _, _, q in got["sequential"])
        base = _norm(got["sequential"], any_qual)
        for m in ("tiny", "small", "default"):
            assert _norm(got[m], any_qual) == base, (seed, t, m)
        batches = {m: _ours_parts(p, m) for m in ("tiny", "sequential")}
        any_q2 = any(q is not None for _, _, q in batches["sequential"])
        assert _norm(batches["tiny"], any_q2) == _norm(batches["sequential"], any_q2), (seed, t, "batches")
        if have_ref:
            first = _ref_read(p, first_only=True)      # one kseq_read loop: ends at the first record kseq rejects
            assert _norm(first, any_qual) == base, (seed, t)
            want = _ref_read(p)                         # mm_bseq_read batches: go on after a rejected record
            assert _norm(want, any_q2) == _norm(batches["sequential"], any_q2), (seed, t, "batches vs reference")


def test_reader_big_file_parts_and_chunks(tmp_path):
    """a real multi-block file: part boundaries (index.c:244,316) and record order with many scan blocks and threads"""
    from longqc_b200 import synth
    import longqc_b200 as L
    T, _ = synth.standard_set(3000, 3000, 0.1, seed=12, n_query=2)
    p = str(tmp_path / "t.fq")
    T.write_fastx(p)
    os.environ["LQCOV_READER_BLOCK"] = "65536"
    try:
        whole = _read_all(p)
        lib = L.load()
        from longqc_b200 import _lib
        from longqc_b200.dist import part_boundaries
        r = lib.lqcov_reader_open(p.encode())
        st = _lib.ReadsStruct()
        sizes = []
        while lib.lqcov_reader_next_part(r, 2_000_000, 500_000, C.byref(st)) > 0:
            sizes.append(st.n)
        lib.lqcov_reader_close(r)
    finally:
        os.environ.pop("LQCOV_READER_BLOCK")
    assert [w[0] for w in whole] == T.names and all(w[1] == T.seq[T.seq_off[i]:T.seq_off[i + 1]].tobytes() for i, w in enumerate(whole))
    assert sizes == [e - s for s, e in part_boundaries(T.lengths(), 2_000_000, 500_000)]
