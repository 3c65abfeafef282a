"""CPU tier: the N>1 plan (rank-ordered target ownership, global part boundaries, query sharding, row merge)
exercised with world_size 2 over gloo.  The compute inside each rank is the oracle -- the point here is the
host-side sharding logic of longqc_b200/dist.py, which must reproduce the single-process table."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import liblq


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from longqc_b200.dist import part_boundaries, split_even
    T, Q = cases.make_case(name)
    opts = cases.opts(name)
    # ownership: rank-ordered contiguous target ranges; lengths are all-gathered like dist.Runner does
    lo, hi = split_even(T.n, world)[rank]
    mine = torch.from_numpy(T.lengths()[lo:hi].astype(np.int64))
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(mine)]))
    bufs = [torch.zeros(int(s.item()), dtype=torch.int64) for s in sizes]
    pad = torch.zeros(max(int(s.item()) for s in sizes), dtype=torch.int64)
    pad[:len(mine)] = mine
    gathered = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    all_lens = np.concatenate([g[:int(s.item())].numpy() for g, s in zip(gathered, sizes)])
    assert np.array_equal(all_lens, T.lengths())
    parts = part_boundaries(all_lens, opts.get("batch_size", 4000000000))
    # per-part minimizer counts: local shard counts summed over ranks == whole-part counts (the all-reduce invariant)
    for (s, e) in parts:
        a, b = max(s, lo), min(e, hi)
        local = liblq.oracle_sketch_set(T.subset(range(a, b)), 5, opts.get("k", 12), rid_base=a - s) if b > a else None
        nloc = torch.tensor([0 if local is None else len(local)], dtype=torch.int64)
        dist.all_reduce(nloc)
        whole = liblq.oracle_sketch_set(T.subset(range(s, e)), 5, opts.get("k", 12))
        assert int(nloc.item()) == len(whole)
        if local is not None and len(local):   # rid_base keeps global y order: my records are a contiguous slice of the whole
            pos = np.searchsorted(whole["y"], local["y"][0])
            assert np.array_equal(whole[pos:pos + len(local)], local)
    # query sharding + merge
    qlo, qhi = split_even(Q.n, world)[rank]
    table, _, _ = liblq.oracle_table(T, Q.subset(range(qlo, qhi)), liblq.oracle_opt(**opts))
    rows = [None] * world if rank == 0 else None
    dist.gather_object(table, rows, dst=0)
    if rank == 0:
        open(out, "wb").write(b"".join(rows))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["plain_pb", "tandem_parts"])
def test_two_rank_plan_reproduces_single_process_table(name, tmp_path):
    out = str(tmp_path / "merged.tsv")
    mp.spawn(_worker, args=(2, _free_port(), name, out), nprocs=2, join=True)
    want = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".tsv"), "rb").read()
    assert open(out, "rb").read() == want
