"""ctypes access to the oracle (test infrastructure), the optional compiled reference
(oracle/_ref/libmm2ref.so) and the product's test-only host build (liblqcov_hostcheck.so)."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liblqoracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmm2ref.so")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "minimap2-coverage")
REF_SDUST = os.path.join(ROOT, "oracle", "_ref", "sdust")
ORACLE_CLI = os.path.join(ROOT, "oracle", "lq_oracle_cli")
HOSTCHECK_SO = os.path.join(ROOT, "longqc_b200", "csrc", "liblqcov_hostcheck.so")


class MM128(C.Structure):
    _fields_ = [("x", C.c_uint64), ("y", C.c_uint64)]


class MM128V(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(MM128))]


mm128_dtype = np.dtype([("x", np.uint64), ("y", np.uint64)])

_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def _vec_to_np(v):
    n = v.n
    if n == 0:
        out = np.zeros(0, dtype=mm128_dtype)
    else:
        out = np.frombuffer(C.string_at(v.a, n * 16), dtype=mm128_dtype).copy()
    if v.a:
        _libc.free(C.cast(v.a, C.c_void_p))
    return out


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(ORACLE_SO)
        lib.lqo_sketch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(MM128V)]
        lib.lqo_radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
        lib.lqo_hash64.argtypes = [C.c_uint64, C.c_uint64]
        lib.lqo_hash64.restype = C.c_uint64
        lib.lqo_meanQ.argtypes = [C.c_char_p, C.c_int]
        lib.lqo_meanQ.restype = C.c_double
        _oracle = lib
    return _oracle


_ref = None


def ref():
    """The compiled, unmodified reference functions; None when oracle/_ref was not built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        lib = C.CDLL(REF_SO)
        lib.mm_sketch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(MM128V)]
        lib.radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
        lib.meanQ.argtypes = [C.c_char_p, C.c_int]
        lib.meanQ.restype = C.c_double
        _ref = lib
    return _ref


def oracle_sketch(seq: bytes, w, k, rid=0, hpc=0):
    v = MM128V(0, 0, None)
    oracle().lqo_sketch(seq, len(seq), w, k, rid, hpc, C.byref(v))
    return _vec_to_np(v)


def ref_sketch(seq: bytes, w, k, rid=0, hpc=0):
    v = MM128V(0, 0, None)
    ref().mm_sketch(None, seq, len(seq), w, k, rid, hpc, C.byref(v))
    return _vec_to_np(v)


_hc = None


def hostcheck():
    global _hc
    if _hc is None:
        lib = C.CDLL(HOSTCHECK_SO)
        for name in ("lqhc_sketch_parallel", "lqhc_sketch_parallel_win", "lqhc_sketch_slow_everywhere"):
            f = getattr(lib, name)
            f.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_int]
            f.restype = C.c_int
        lib.lqhc_sketch_replay.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_int]
        lib.lqhc_sketch_replay.restype = C.c_int
        lib.lqhc_sketch_pk.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        lib.lqhc_sketch_pk.restype = C.c_int
        lib.lqhc_hash32.argtypes = [C.c_uint32, C.c_uint32]
        lib.lqhc_hash32.restype = C.c_uint32
        lib.lqhc_hash64.argtypes = [C.c_uint64, C.c_uint64]
        lib.lqhc_hash64.restype = C.c_uint64
        _hc = lib
    return _hc


def hc_sketch(fn, seq: bytes, w, k, rid=0, *extra):
    cap = 2 * len(seq) + 64
    out = np.zeros(cap, dtype=mm128_dtype)
    n = getattr(hostcheck(), fn)(seq, len(seq), w, k, rid, *extra, out.ctypes.data, cap)
    assert n <= cap
    return out[:n]


def adversarial_seqs(rng, n, max_len=3000):
    """AT-only, tandem repeats, homopolymers, N-sprinkled and plain random sequences."""
    out = []
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for i in range(n):
        L = int(rng.integers(1, max_len + 1))
        kind = i % 7
        if kind == 0:
            s = acgt[rng.integers(0, 4, L)]
        elif kind == 1:  # AT only
            s = np.frombuffer(b"AT", dtype=np.uint8)[rng.integers(0, 2, L)]
        elif kind == 2:  # tandem repeat, short unit
            u = acgt[rng.integers(0, 4, int(rng.integers(1, 13)))]
            s = np.tile(u, L // len(u) + 1)[:L]
        elif kind == 3:  # homopolymer runs
            runs = rng.integers(1, 40, L)
            s = np.repeat(acgt[rng.integers(0, 4, L)], runs)[:L]
        elif kind == 4:  # 5 % N
            s = acgt[rng.integers(0, 4, L)].copy()
            s[rng.random(L) < 0.05] = ord("N")
        elif kind == 5:  # strict (AT)n / (ACGT)n palindromic runs with random flanks
            core = np.tile(np.frombuffer(b"AT" if rng.random() < 0.5 else b"ACGT", dtype=np.uint8), L)[:L]
            s = core.copy()
            m = rng.random(L) < 0.02
            s[m] = acgt[rng.integers(0, 4, int(m.sum()))]
        else:  # mixed: random with a repeat block and a few N and lowercase/U
            s = acgt[rng.integers(0, 4, L)].copy()
            if L > 50:
                a = int(rng.integers(0, L - 40)); s[a:a + 40] = np.tile(acgt[rng.integers(0, 4, 2)], 20)
            s[rng.random(L) < 0.003] = ord("N")
            lo = rng.random(L) < 0.1
            s[lo] = s[lo] | 0x20
            s[(s == ord("T")) & (rng.random(L) < 0.05)] = ord("U")
        out.append(s.tobytes())
    return out


# ---------------------------------------------------------------- whole-pipeline oracle (lqo_run) on ReadSets
class LqoOpt(C.Structure):
    _fields_ = [("k", C.c_int), ("w", C.c_int), ("is_hpc", C.c_int), ("batch_size", C.c_uint64), ("mini_batch_size", C.c_int),
                ("no_self", C.c_int), ("ava", C.c_int), ("max_gap", C.c_int), ("min_cnt", C.c_int), ("min_chain_score", C.c_int),
                ("min_score_med", C.c_int), ("min_score_good", C.c_int), ("max_chain_skip", C.c_int), ("bw", C.c_int),
                ("mid_occ_frac", C.c_float), ("seed", C.c_int), ("max_overhang", C.c_int), ("min_ovlp", C.c_int),
                ("min_coverage", C.c_int), ("min_ratio", C.c_double), ("filter", C.c_int)]


class LqoReads(C.Structure):
    _fields_ = [("n", C.c_int), ("name", C.POINTER(C.c_char_p)), ("seq", C.POINTER(C.c_char_p)), ("qual", C.POINTER(C.c_char_p)),
                ("len", C.POINTER(C.c_int))]


class LqoTrace(C.Structure):
    _fields_ = [("mid_occ", C.c_int), ("n_mini", C.c_int), ("n_kept", C.c_int), ("n_seeds", C.c_int64),
                ("seeds_unsorted", C.POINTER(MM128)), ("seeds_sorted", C.POINTER(MM128)), ("mini_pos", C.POINTER(C.c_uint64)),
                ("n_chains", C.c_int), ("u", C.POINTER(C.c_uint64)), ("anchors", C.POINTER(MM128)), ("n_anchors", C.c_int64)]


def lqo_reads(rs):
    """synth.ReadSet -> lqo_reads (keeps the byte strings alive on the returned object)."""
    n = rs.n
    seqb = rs.seq.tobytes()
    qualb = None if rs.qual is None else rs.qual.tobytes()
    off = rs.seq_off
    seqs = [seqb[int(off[i]):int(off[i + 1])] for i in range(n)]
    quals = [None if qualb is None else qualb[int(off[i]):int(off[i + 1])] for i in range(n)]
    r = LqoReads()
    r.n = n
    r._seqs, r._quals, r._names = seqs, quals, list(rs.names)
    r._a = (C.c_char_p * n)(*r._names)
    r._b = (C.c_char_p * n)(*seqs)
    r._c = (C.c_char_p * n)(*quals)
    r._d = (C.c_int * n)(*[len(s) for s in seqs])
    r.name, r.seq, r.qual, r.len = r._a, r._b, r._c, r._d
    return r


def oracle_opt(**kw):
    o = LqoOpt()
    oracle().lqo_opt_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def opt_pair(**kw):
    """The same options for the product (longqc_b200.Opt) and the oracle (LqoOpt)."""
    import longqc_b200 as L
    return L.Opt(**kw), oracle_opt(**kw)


def oracle_table(targets, queries, oopt):
    """stdout table of the CPU restatement; returns (bytes, mid_occ, n_parts)."""
    import tempfile
    lib = oracle()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    lib.lqo_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    t, q = lqo_reads(targets), lqo_reads(queries)
    mid, parts = C.c_int(), C.c_int()
    with tempfile.NamedTemporaryFile(suffix=".tsv") as tf:
        fp = libc.fopen(tf.name.encode(), b"w")
        rc = lib.lqo_run(C.byref(oopt), C.byref(t), C.byref(q), fp, C.byref(mid), C.byref(parts))
        libc.fclose(fp)
        assert rc == 0
        data = open(tf.name, "rb").read()
    return data, mid.value, parts.value


def oracle_trace(targets, queries, qi, oopt, mid_occ=0):
    lib = oracle()
    lib.lqo_trace_query.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.lqo_trace_free.argtypes = [C.c_void_p]
    t, q = lqo_reads(targets), lqo_reads(queries)
    tr = LqoTrace()
    assert lib.lqo_trace_query(C.byref(oopt), C.byref(t), C.byref(q), qi, mid_occ, C.byref(tr)) == 0
    n = tr.n_seeds

    def grab(p, cnt):
        return np.frombuffer(C.string_at(p, cnt * 16), dtype=mm128_dtype).copy() if cnt else np.zeros(0, dtype=mm128_dtype)
    out = dict(mid_occ=tr.mid_occ, n_mini=tr.n_mini, n_kept=tr.n_kept, n_seeds=n, unsorted=grab(tr.seeds_unsorted, n),
               sorted=grab(tr.seeds_sorted, n), n_chains=tr.n_chains,
               u=np.frombuffer(C.string_at(tr.u, tr.n_chains * 8), dtype=np.uint64).copy() if tr.n_chains else np.zeros(0, np.uint64),
               anchors=grab(tr.anchors, tr.n_anchors))
    lib.lqo_trace_free(C.byref(tr))
    return out


def oracle_sketch_set(rs, w, k, hpc=0, rid_base=0):
    seqb = rs.seq.tobytes()
    parts = []
    for i in range(rs.n):
        s = seqb[int(rs.seq_off[i]):int(rs.seq_off[i + 1])]
        if len(s):
            parts.append(oracle_sketch(s, w, k, rid_base + i, hpc))
    return np.concatenate(parts) if parts else np.zeros(0, dtype=mm128_dtype)


def reads_from_seqs(seqs, prefix=b"s", qual=False, rng=None):
    from longqc_b200.synth import ReadSet
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    seq = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    q = None
    if qual:
        q = (rng.integers(0, 41, size=len(seq), dtype=np.uint8) + 33).astype(np.uint8)
    return ReadSet(seq, off, q, [prefix + str(i).encode() for i in range(len(seqs))])


def sdust_stale_seqs(rng=None):
    """low-complexity repeats interrupted by a non-ACGT base: sdust keeps its deque and counts across the N (sdust.c:158-162), so
    find_perfect inserts up to W-2 stale intervals per base for ~W bases -- the worst case of the perfect-interval list"""
    out = [b"A" * 200 + b"N" + b"A" * 200, b"AC" * 50 + b"N" + b"AC" * 50, b"ACG" * 40 + b"NN" + b"ACG" * 40 + b"N" + b"T" * 90,
           b"T" * 70 + b"N" + b"T" * 3 + b"N" + b"T" * 64 + b"X" + b"AT" * 100, b"G" * 63 + b"N" + b"G" * 63 + b"N" + b"G" * 63,
           b"A" * 64 + b"N" + b"C" * 64 + b"N" + b"AAC" * 30 + b"n" + b"AAC" * 30]
    if rng is not None:
        acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
        for _ in range(40):
            parts = []
            for _ in range(int(rng.integers(2, 6))):
                u = acgt[rng.integers(0, 4, int(rng.integers(1, 4)))]
                parts.append(np.tile(u, int(rng.integers(20, 120)))[: int(rng.integers(30, 260))].tobytes())
            out.append(b"N".join(parts))
    return out


def oracle_sdust_table(rs, W=64, T=20):
    import tempfile
    lib = oracle()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    lib.lqo_sdust_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    r = lqo_reads(rs)
    with tempfile.NamedTemporaryFile(suffix=".tsv") as tf:
        fp = libc.fopen(tf.name.encode(), b"w")
        assert lib.lqo_sdust_run(C.byref(r), W, T, fp) == 0
        libc.fclose(fp)
        return open(tf.name, "rb").read()
