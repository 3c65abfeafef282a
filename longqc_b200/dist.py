"""Multi-GPU plumbing: one process per GPU; torch.distributed only hands the NCCL id around, the exchange step runs in liblqcov.so.

How one index part is shared by N GPUs:
  * the part's reads are owned by the ranks in contiguous, rank-ordered ranges, so rid == position in
    the part and the concatenation of the ranks' minimizer records is already ordered by y;
  * every rank packs + sketches its own reads and counts its minimizers (lqcov_part_sketch);
  * lqcov_part_exchange (lq_comm.cu, NCCL inside the C library): ALL-REDUCE (sum) of the per-minimizer
    count table -- the counts decide mid_occ (index.c:123-144) and the high-frequency filter
    (lqmap.c:159,166) --, every rank sorts only its own records by key, the sorted shards are exchanged
    and placed into the replicated index (slot = global offset + occurrences on lower ranks + own rank);
  * the queries are split across the ranks, each rank maps its share against every part
    (lqcov_map_part) and rank 0 concatenates the rows in query order.
Parts follow the reference's mini-batch rule on the GLOBAL read list (index.c:238-246).

Runner is also what bench.py drives at N=1 (no collective is issued then).
"""
from __future__ import annotations

import ctypes as C
import hashlib
from typing import List, Tuple

import numpy as np

from . import _lib
from . import synth


def part_boundaries(lengths: np.ndarray, batch_size: int, mini_batch_size: int = 50_000_000) -> List[Tuple[int, int]]:
    """[start, end) read ranges of the index parts: a part keeps taking mini-batches (each = reads until
    its own size >= min(mini_batch_size, batch_size)) while the part's sum of lengths is <= batch_size
    (reference index.c:244,316; bseq.c:82-87)."""
    mini = min(int(mini_batch_size), int(batch_size))
    out, n, t0 = [], len(lengths), 0
    while t0 < n:
        sum_len, t1 = 0, t0
        while t1 < n and not (sum_len > batch_size):
            sz = 0
            while t1 < n:
                sz += int(lengths[t1]); sum_len += int(lengths[t1]); t1 += 1
                if sz >= mini:
                    break
        out.append((t0, t1))
        t0 = t1
    return out


def split_even(n: int, world: int) -> List[Tuple[int, int]]:
    """contiguous, rank-ordered shares of n items"""
    base, rem = divmod(n, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


class _DevArr:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def _view(ptr, n, typestr):
    import torch
    return torch.as_tensor(_DevArr(ptr, n, typestr), device="cuda")


def _meta_struct(names: List[bytes], lengths: np.ndarray):
    """lqcov_reads_t describing a whole part (lengths + names, no bases)"""
    off = np.zeros(len(lengths) + 1, dtype=np.uint64)
    np.cumsum(lengths.astype(np.uint64), out=off[1:])
    blob = b"".join(names)
    nlen = np.fromiter((len(x) for x in names), dtype=np.uint64, count=len(names))
    noff = np.zeros(len(names) + 1, dtype=np.uint64)
    np.cumsum(nlen, out=noff[1:])
    nbuf = np.frombuffer(blob if blob else b"\0", dtype=np.uint8)
    st = _lib.ReadsStruct()
    st.n = len(lengths); st.seq = None; st.seq_off = off.ctypes.data; st.qual = None
    st.names = nbuf.ctypes.data; st.name_off = noff.ctypes.data; st.seq_on_device = 0
    return _lib._Keep(st, (off, nbuf, noff))


def _sub_struct(keep, a: int, b: int, seq_ptr, on_device: int):
    """reads [a, b) of a reads_struct(), with the bases taken from `seq_ptr` (same offsets)"""
    st = _lib.ReadsStruct()
    st.n = b - a
    st.seq = seq_ptr
    st.seq_off = keep.st.seq_off + 8 * a
    st.qual = None
    st.names = keep.st.names
    st.name_off = keep.st.name_off + 8 * a
    st.seq_on_device = on_device
    return st


def comm_init(cov, rank: int, world: int):
    """NCCL communicator of the C library (lq_comm.cu) for this rank's context: rank 0 makes the 128-byte id, the launcher's
    process group (torch.distributed: plumbing only) hands it to the other ranks.  All data-path collectives then run inside
    liblqcov.so (lqcov_part_exchange, lqcov_comm_gather_rows)."""
    if world == 1:
        return
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    lib.lqcov_comm_unique_id.argtypes = [C.c_void_p]
    lib.lqcov_comm_init_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    buf = C.create_string_buffer(128)
    if rank == 0 and lib.lqcov_comm_unique_id(buf) != 0:
        raise _lib.LqcovError("lqcov_comm_unique_id failed")
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    ident = bytes(t.cpu().tolist())
    if lib.lqcov_comm_init_rank(cov._h, ident, world, rank) != 0:
        raise _lib.LqcovError("lqcov_comm_init_rank failed")


def run_job(cov, t_keep, n_my_targets, my_lo, parts, part_meta, q_keep, n_my_queries, tptr, qptr, on_device, rank, world, exchange=None):
    """One coverage job of this rank: its queries against every index part; targets [my_lo, my_lo+n_my_targets) of the global
    read list are this rank's to sketch.  Returns the table (all ranks' rows on rank 0 when world > 1).  The context must have a
    communicator (comm_init) when world > 1."""
    lib = _lib.load()
    lib.lqcov_part_sketch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    lib.lqcov_part_finish.argtypes = [C.c_void_p, C.c_void_p]
    lib.lqcov_map_part.argtypes = [C.c_void_p]
    lib.lqcov_part_exchange.argtypes = [C.c_void_p]
    lib.lqcov_comm_gather_rows.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    h = cov._h
    lib.lqcov_reset(h)
    qs = _sub_struct(q_keep, 0, n_my_queries, qptr, on_device)
    qs.qual = q_keep.st.qual
    if lib.lqcov_set_queries(h, C.byref(qs)) != 0:
        raise _lib.LqcovError("lqcov_set_queries failed")
    my_hi = my_lo + n_my_targets
    for (s, e), meta in zip(parts, part_meta):
        a, b = max(s, my_lo), min(e, my_hi)                # my reads inside this part
        if b < a:
            a = b = max(min(s, my_hi), my_lo)
        sub = _sub_struct(t_keep, a - my_lo, b - my_lo, tptr, on_device)
        if lib.lqcov_part_sketch(h, C.byref(sub), a - s if b > a else 0) != 0:
            raise _lib.LqcovError("lqcov_part_sketch failed")
        if world > 1 and lib.lqcov_part_exchange(h) != 0:   # all-reduce of the counts, own shard sorted, shards exchanged + placed
            raise _lib.LqcovError("lqcov_part_exchange failed")
        if lib.lqcov_part_finish(h, C.byref(meta.st)) != 0:
            raise _lib.LqcovError("lqcov_part_finish failed")
        if lib.lqcov_map_part(h) != 0:
            raise _lib.LqcovError("lqcov_map_part failed")
    table = cov.table()
    if world > 1:
        out, n = C.c_void_p(), C.c_size_t()
        if lib.lqcov_comm_gather_rows(h, table, len(table), C.byref(out), C.byref(n)) != 0:
            raise _lib.LqcovError("lqcov_comm_gather_rows failed")
        if rank == 0:
            table = C.string_at(out.value, n.value)
        if out.value:
            lib.lqcov_free(out)
    return table


def rank_inputs(a, rank: int, world: int):
    """(targets, queries) of one rank of bench.py's workload: every rank owns `a.reads` target reads of ONE genome shared by all
    ranks (30x overall) and maps `a.queries / world` of its own reads.  Deterministic in (a.seed, rank, world), so any process can
    regenerate any rank's share (bench.py --impl reference, the N > 1 table check)."""
    rng_g = np.random.default_rng(a.seed)
    G = max(int(a.reads * world * a.read_len / 30.0), 2 * a.read_len)
    genome = synth.make_genome(G, rng_g)
    rng = np.random.default_rng(a.seed + 7919 * (rank + 1))
    targets = synth.simulate_reads(genome, a.reads, a.read_len, a.err, rng, name_start=rank * a.reads)
    nq_lo, nq_hi = split_even(a.queries, world)[rank]
    qidx = np.sort(rng.choice(a.reads, size=min(nq_hi - nq_lo, a.reads), replace=False))
    return targets, targets.subset(qidx)


class Runner:
    """One rank of the (possibly multi-GPU) coverage job on the synthetic workload of bench.py."""

    def __init__(self, a, opt, rank: int, world: int, local: int):
        self.a, self.opt, self.rank, self.world, self.local = a, opt, rank, world, local
        self.cov = None
        self.last_stats = {}
        self.tables = []

    # -- inputs: every rank owns `reads` target reads of ONE genome shared by all ranks (30x overall) and
    #    maps `queries/world` of its own reads
    def make_inputs(self):
        import torch
        a, world, rank = self.a, self.world, self.rank
        self.targets, self.queries = rank_inputs(a, rank, world)
        # global read list = rank-major: lengths and names of every rank's reads
        lens = self.targets.lengths().astype(np.int64)
        if world > 1:
            import torch.distributed as dist
            t = torch.from_numpy(lens).cuda()
            allt = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(allt, t)
            self.all_lens = np.concatenate([x.cpu().numpy() for x in allt])
            qb = torch.tensor([self.queries.n_bases], dtype=torch.int64, device="cuda")
            dist.all_reduce(qb)
            self.query_bases_all = int(qb.item())
        else:
            self.all_lens = lens
            self.query_bases_all = self.queries.n_bases
        self.target_bases_all = int(self.all_lens.sum())
        self.all_names = [b"r" + str(i).encode() for i in range(a.reads * world)]
        self.parts = part_boundaries(self.all_lens, int(self.opt.batch_size), int(self.opt.mini_batch_size))
        self.part_meta = [_meta_struct(self.all_names[s:e], self.all_lens[s:e]) for s, e in self.parts]
        # host (pinned) and device copies of the bases
        self.t_keep = _lib.reads_struct(self.targets)
        self.q_keep = _lib.reads_struct(self.queries)
        self.t_pin = torch.from_numpy(self.targets.seq).pin_memory()
        self.q_pin = torch.from_numpy(self.queries.seq).pin_memory()
        self.t_dev = self.t_pin.cuda()
        self.q_dev = self.q_pin.cuda()
        torch.cuda.synchronize()
        self.cov = _lib.Coverage(self.opt)
        comm_init(self.cov, rank, world)
        return self.targets, self.queries

    def job_bases(self):
        return self.target_bases_all + self.query_bases_all

    def parallelism(self):
        if self.world == 1:
            return "1 GPU"
        return ("%d GPUs: targets sharded for sketch + count + sort, NCCL all-reduce of the 4^k count table, key-sorted shards exchanged and placed into "
                "the replicated index (lqcov_part_exchange, NCCL inside liblqcov.so), queries sharded" % self.world)

    def step(self, resident: bool):
        tptr = self.t_dev.data_ptr() if resident else self.t_pin.data_ptr()
        qptr = self.q_dev.data_ptr() if resident else self.q_pin.data_ptr()
        table = run_job(self.cov, self.t_keep, self.targets.n, self.rank * self.a.reads, self.parts, self.part_meta, self.q_keep, self.queries.n,
                        tptr, qptr, 1 if resident else 0, self.rank, self.world)
        self.last_stats = self.cov.stats()
        self.last_table = table
        return table

    def verify_against_one_gpu(self):
        """N > 1: rank 0 runs the WHOLE job (every rank's targets in one index, all queries) on its own GPU alone and compares the
        table with the merged table of the N-GPU job.  Untimed.  The other ranks only ship their reads (a padded gather)."""
        import torch
        import torch.distributed as dist
        world, rank = self.world, self.rank
        if world == 1:
            return None

        def gather_bytes(arr):
            n = torch.tensor([arr.size], dtype=torch.int64, device="cuda")
            sizes = [torch.zeros_like(n) for _ in range(world)]
            dist.all_gather(sizes, n)
            sizes = [int(x.item()) for x in sizes]
            pad = torch.zeros(max(sizes), dtype=torch.uint8, device="cuda")
            pad[:arr.size] = torch.from_numpy(arr).cuda()
            out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, out, dst=0)
            return [o[:sz].cpu().numpy() for o, sz in zip(out, sizes)] if rank == 0 else None

        def gather_set(rs):
            seqs = gather_bytes(np.ascontiguousarray(rs.seq))
            quals = gather_bytes(np.ascontiguousarray(rs.qual)) if rs.qual is not None else None
            lens = gather_bytes(np.ascontiguousarray(rs.lengths().astype(np.int64)).view(np.uint8))
            names = [None] * world if rank == 0 else None
            dist.gather_object(rs.names, names, dst=0)
            if rank != 0:
                return None
            sets = []
            for r in range(world):
                ln = lens[r].view(np.int64)
                off = np.zeros(len(ln) + 1, dtype=np.int64)
                np.cumsum(ln, out=off[1:])
                sets.append(synth.ReadSet(seqs[r], off, None if quals is None else quals[r], names[r]))
            return synth.ReadSet.concat(sets)

        T, Q = gather_set(self.targets), gather_set(self.queries)
        res = None
        if rank == 0:
            with _lib.Coverage(self.opt) as cov:            # a fresh context without a communicator: the plain 1-GPU path
                cov.set_queries(Q)
                cov.add_targets(T)
                one = cov.table()
            res = "identical (%d rows, byte for byte)" % one.count(b"\n") if one == self.last_table else "DIFFERS"
        dist.barrier()
        return res

    def parity_note(self):
        """md5 / row count of the benchmarked table (rank 0) -- the exact check lives in tests/ (oracle at small sizes)."""
        if self.rank != 0:
            return None
        t = self.last_table
        return {"rows": t.count(b"\n"), "md5": hashlib.md5(t).hexdigest(),
                "nonsense_frac": sum(1 for ln in t.split(b"\n") if ln and ln.split(b"\t")[4] == b"0") / max(1, t.count(b"\n"))}
