"""Multi-GPU parity check (run under torchrun on N GPUs): the sharded job of longqc_b200/dist.py on a seeded case must
reproduce the golden table of the unmodified reference.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py parts tandem_parts plain_pb
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import longqc_b200 as L  # noqa: E402
from longqc_b200 import _lib, dist as lqd  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name in sys.argv[1:] or ["plain_pb"]:
        T, Q = cases.make_case(name)
        o = cases.opts(name)
        o["device"] = local
        opt = L.Opt(**o)
        lens = T.lengths().astype(np.int64)
        parts = lqd.part_boundaries(lens, int(opt.batch_size), int(opt.mini_batch_size))
        metas = [lqd._meta_struct(T.names[s:e], lens[s:e]) for s, e in parts]
        lo, hi = lqd.split_even(T.n, world)[rank]
        qlo, qhi = lqd.split_even(Q.n, world)[rank]
        myT, myQ = T.subset(range(lo, hi)), Q.subset(range(qlo, qhi))
        tk, qk = _lib.reads_struct(myT), _lib.reads_struct(myQ)
        with L.Coverage(opt) as cov:
            lqd.comm_init(cov, rank, world)
            table = lqd.run_job(cov, tk, myT.n, lo, parts, metas, qk, myQ.n, tk.st.seq, qk.st.seq, 0, rank, world)
        if rank == 0:
            want = open(os.path.join(ROOT, "tests", "golden", name + ".tsv"), "rb").read()
            same = table == want
            ok &= same
            print("dist_check %-18s world=%d parts=%d rows=%d  %s" % (name, world, len(parts), table.count(b"\n"), "IDENTICAL to the reference golden" if same else "DIFFERS"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
