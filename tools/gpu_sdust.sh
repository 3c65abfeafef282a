#!/bin/bash
# sdust visit: parity tests, then the sdust table over the bench workload (and over reads with repeats sprinkled in)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_cli.py -m gpu -x -q -k "sdust" ) > gpurun_out/pytest_sdust.log 2>&1
tail -4 gpurun_out/pytest_sdust.log
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/sdust_bench.log
import sys, time, json, argparse
import numpy as np
sys.path.insert(0, '.')
import bench, longqc_b200 as L
from longqc_b200 import synth
sys.argv = ['bench.py']
a = bench.parse()
T, Q = bench.global_workload(a, 1)
print("reads", T.n, "bases", T.n_bases)
out = bench.sdust_line(a, L, T)
print(json.dumps(out))
# repeats sprinkled in: every read gets a few microsatellites / homopolymers of 10..80 bases
rng = np.random.default_rng(3)
seq = T.seq.copy()
off = np.asarray(T.seq_off)
acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
n_ins = 0
for r in range(0, T.n):
    L0, L1 = int(off[r]), int(off[r + 1])
    for _ in range(max(1, (L1 - L0) // 3000)):
        ln = int(rng.integers(10, 80)); 
        if L1 - L0 <= ln + 2: continue
        at = int(rng.integers(L0, L1 - ln))
        u = acgt[rng.integers(0, 4, int(rng.integers(1, 4)))]
        seq[at:at + ln] = np.tile(u, ln)[:ln]; n_ins += 1
T2 = synth.ReadSet(seq, T.seq_off, T.qual, T.names)
print("inserted", n_ins, "repeats")
out2 = bench.sdust_line(a, L, T2)
print(json.dumps(out2))
PY
