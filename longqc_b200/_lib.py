"""ctypes binding of include/lqcov.h (liblqcov.so).  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class LqcovError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(_HERE, "liblqcov.so")


def bin_path(name: str) -> str:
    return os.path.join(_HERE, "bin", name)


class Opt(C.Structure):
    """lqcov_opt_t"""
    _fields_ = [
        ("k", C.c_int), ("w", C.c_int), ("is_hpc", C.c_int), ("batch_size", C.c_uint64), ("mini_batch_size", C.c_int),
        ("no_self", C.c_int), ("ava", C.c_int), ("max_gap", C.c_int), ("min_cnt", C.c_int), ("min_chain_score", C.c_int),
        ("min_score_med", C.c_int), ("min_score_good", C.c_int), ("max_chain_skip", C.c_int), ("bw", C.c_int),
        ("mid_occ_frac", C.c_float), ("max_overhang", C.c_int), ("min_ovlp", C.c_int), ("min_coverage", C.c_int),
        ("min_ratio", C.c_double), ("filter", C.c_int), ("n_threads", C.c_int), ("device", C.c_int),
        ("seed_budget", C.c_uint64), ("verbose", C.c_int),
    ]

    def __init__(self, **kw):
        super().__init__()
        load().lqcov_opt_init(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)


class ReadsStruct(C.Structure):
    """lqcov_reads_t"""
    _fields_ = [("n", C.c_uint32), ("seq", C.c_void_p), ("seq_off", C.c_void_p), ("qual", C.c_void_p),
                ("names", C.c_void_p), ("name_off", C.c_void_p), ("seq_on_device", C.c_int)]


class Stats(C.Structure):
    """lqcov_stats_t"""
    _fields_ = [("target_bases", C.c_uint64), ("query_bases", C.c_uint64), ("target_minimizers", C.c_uint64),
                ("query_minimizers", C.c_uint64), ("seeds", C.c_uint64), ("groups", C.c_uint64), ("chains", C.c_uint64),
                ("overlaps", C.c_uint64), ("batches", C.c_uint64), ("walk_buckets", C.c_uint64), ("mid_occ", C.c_int32),
                ("n_parts", C.c_int32), ("t_upload_ms", C.c_double), ("t_sketch_ms", C.c_double), ("t_index_ms", C.c_double),
                ("t_map_ms", C.c_double), ("t_post_ms", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_lib = None

SYMBOLS = ["lqcov_abi_version", "lqcov_device_count", "lqcov_opt_init", "lqcov_create", "lqcov_destroy", "lqcov_reset", "lqcov_part_sketch", "lqcov_part_device_views", "lqcov_part_gather_buffers", "lqcov_part_finish", "lqcov_map_part", "lqcov_profile_enable", "lqcov_profile_reset", "lqcov_profile_json", "lqcov_set_queries", "lqcov_add_part",
           "lqcov_add_targets", "lqcov_table", "lqcov_get_stats", "lqcov_free", "lqcov_sdust_table", "lqcov_sketch",
           "lqcov_debug_seeds", "lqcov_index_part", "lqcov_reader_open", "lqcov_reader_next", "lqcov_reader_next_part",
           "lqcov_reader_close", "lqcov_main", "lqcov_sdust_main", "lqcov_part_begin", "lqcov_stage", "lqcov_part_chunk", "lqcov_stage_wait", "lqcov_part_end",
           "lqcov_comm_unique_id", "lqcov_comm_init_rank", "lqcov_comm_init_all", "lqcov_comm_size", "lqcov_comm_rank", "lqcov_part_exchange", "lqcov_comm_gather_rows", "lqcov_sdust_begin", "lqcov_sdust_chunk", "lqcov_sdust_end", "lqcov_index_dump", "lqcov_index_peek", "lqcov_load_part", "lqcov_set_prepass_counts"]


def load() -> C.CDLL:
    """Load liblqcov.so (built in-tree by __graft_entry__.build()).  Fails loudly when missing."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise LqcovError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(p)
    lib.lqcov_abi_version.restype = C.c_int
    lib.lqcov_opt_init.argtypes = [C.c_void_p]
    lib.lqcov_create.argtypes = [C.c_void_p]
    lib.lqcov_create.restype = C.c_void_p
    lib.lqcov_destroy.argtypes = [C.c_void_p]
    lib.lqcov_reset.argtypes = [C.c_void_p]
    for f in ("lqcov_set_queries", "lqcov_add_part", "lqcov_add_targets", "lqcov_index_part"):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_void_p]
        getattr(lib, f).restype = C.c_int
    lib.lqcov_table.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.lqcov_get_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.lqcov_free.argtypes = [C.c_void_p]
    lib.lqcov_sdust_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    lib.lqcov_sketch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.lqcov_debug_seeds.argtypes = [C.c_void_p, C.c_uint32] + [C.POINTER(C.c_void_p)] * 4 + [C.POINTER(C.c_uint64)]
    lib.lqcov_reader_open.argtypes = [C.c_char_p]
    lib.lqcov_reader_open.restype = C.c_void_p
    lib.lqcov_reader_next.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.lqcov_reader_next_part.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    lib.lqcov_reader_close.argtypes = [C.c_void_p]
    _lib = lib
    return lib


class _Keep:
    """A lqcov_reads_t plus the numpy arrays that own its memory."""

    def __init__(self, st, owners):
        self.st = st
        self.owners = owners


def reads_struct(rs, device_seq_ptr: Optional[int] = None) -> _Keep:
    """synth.ReadSet -> lqcov_reads_t.  ``device_seq_ptr``: CUDA pointer to a copy of rs.seq already in HBM."""
    seq = np.ascontiguousarray(rs.seq, dtype=np.uint8)
    off = np.ascontiguousarray(rs.seq_off, dtype=np.uint64)
    names = b"".join(rs.names)
    nlen = np.fromiter((len(n) for n in rs.names), dtype=np.uint64, count=len(rs.names))
    noff = np.zeros(len(rs.names) + 1, dtype=np.uint64)
    np.cumsum(nlen, out=noff[1:])
    nbuf = np.frombuffer(names if names else b"\0", dtype=np.uint8)
    qual = None if rs.qual is None else np.ascontiguousarray(rs.qual, dtype=np.uint8)
    st = ReadsStruct()
    st.n = rs.n
    st.seq = device_seq_ptr if device_seq_ptr is not None else seq.ctypes.data
    st.seq_off = off.ctypes.data
    st.qual = None if qual is None else qual.ctypes.data
    st.names = nbuf.ctypes.data
    st.name_off = noff.ctypes.data
    st.seq_on_device = 1 if device_seq_ptr is not None else 0
    return _Keep(st, (seq, off, nbuf, noff, qual))


def _take(ptr, nbytes) -> bytes:
    out = C.string_at(ptr, nbytes) if nbytes else b""
    load().lqcov_free(ptr)
    return out


class Coverage:
    """lqcov_ctx: queries -> parts -> table, the library form of `minimap2-coverage`."""

    def __init__(self, opt: Optional[Opt] = None):
        self.opt = opt or Opt()
        self._h = load().lqcov_create(C.byref(self.opt))
        if not self._h:
            raise LqcovError("lqcov_create failed (no usable B200 / bad options); there is no CPU fallback")

    def close(self):
        if self._h:
            load().lqcov_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _rc(self, rc, what):
        if rc != 0:
            raise LqcovError(f"{what} failed (rc={rc})")

    def set_queries(self, rs, device_seq_ptr=None):
        k = reads_struct(rs, device_seq_ptr)
        self._rc(load().lqcov_set_queries(self._h, C.byref(k.st)), "lqcov_set_queries")

    def add_part(self, rs, device_seq_ptr=None):
        k = reads_struct(rs, device_seq_ptr)
        self._rc(load().lqcov_add_part(self._h, C.byref(k.st)), "lqcov_add_part")

    def index_part(self, rs, device_seq_ptr=None):
        k = reads_struct(rs, device_seq_ptr)
        self._rc(load().lqcov_index_part(self._h, C.byref(k.st)), "lqcov_index_part")

    def add_targets(self, rs, device_seq_ptr=None):
        k = reads_struct(rs, device_seq_ptr)
        self._rc(load().lqcov_add_targets(self._h, C.byref(k.st)), "lqcov_add_targets")

    def add_targets_struct(self, keep):
        self._rc(load().lqcov_add_targets(self._h, C.byref(keep.st)), "lqcov_add_targets")

    def set_queries_struct(self, keep):
        self._rc(load().lqcov_set_queries(self._h, C.byref(keep.st)), "lqcov_set_queries")

    def table(self) -> bytes:
        p = C.c_void_p()
        n = C.c_size_t()
        self._rc(load().lqcov_table(self._h, C.byref(p), C.byref(n)), "lqcov_table")
        return _take(p, n.value)

    def stats(self) -> dict:
        s = Stats()
        load().lqcov_get_stats(self._h, C.byref(s))
        return s.as_dict()

    def debug_seeds(self, q: int):
        ps = [C.c_void_p() for _ in range(4)]
        n = C.c_uint64()
        self._rc(load().lqcov_debug_seeds(self._h, q, *[C.byref(p) for p in ps], C.byref(n)), "lqcov_debug_seeds")
        arrs = [np.frombuffer(_take(p, n.value * 8), dtype=np.uint64).copy() for p in ps]
        return arrs  # unsorted x, y, sorted x, y


def coverage_table(targets, queries, opt: Optional[Opt] = None) -> bytes:
    """All-vs-subsample coverage table == stdout of `minimap2-coverage <flags> targets queries`."""
    with Coverage(opt) as c:
        c.set_queries(queries)
        c.add_targets(targets)
        return c.table()


def sketch(rs, opt: Optional[Opt] = None, rid_base: int = 0):
    """(x, y) minimizer records of every read, as mm_sketch lays them out."""
    opt = opt or Opt()
    k = reads_struct(rs)
    px, py, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    rc = load().lqcov_sketch(C.byref(opt), C.byref(k.st), rid_base, C.byref(px), C.byref(py), C.byref(n))
    if rc != 0:
        raise LqcovError("lqcov_sketch failed")
    x = np.frombuffer(_take(px, n.value * 8), dtype=np.uint64).copy()
    y = np.frombuffer(_take(py, n.value * 8), dtype=np.uint64).copy()
    return x, y


def sdust_table(rs, W: int = 64, T: int = 20, opt: Optional[Opt] = None) -> bytes:
    opt = opt or Opt()
    k = reads_struct(rs)
    p, n = C.c_void_p(), C.c_size_t()
    rc = load().lqcov_sdust_table(C.byref(opt), C.byref(k.st), W, T, C.byref(p), C.byref(n))
    if rc != 0:
        raise LqcovError("lqcov_sdust_table failed")
    return _take(p, n.value)
