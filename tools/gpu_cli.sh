#!/bin/bash
# timeline of the drop-in executable on the bench workload (FASTQ files in /dev/shm)
mkdir -p gpurun_out
python - <<'PY'
import sys, os, time, subprocess
sys.path.insert(0, '.')
import bench
class A: pass
a = A(); a.reads=100000; a.read_len=8000; a.err=0.15; a.queries=5000; a.seed=20260925
t0=time.time()
T,Q = bench.global_workload(a, 1)
f = bench.FastqFiles(T,Q)
print("files ready %.1fs" % (time.time()-t0), f.tf)
open('gpurun_out/cli_files.txt','w').write(f.tf+"\n"+f.qf+"\n")
PY
TF=$(sed -n 1p gpurun_out/cli_files.txt); QF=$(sed -n 2p gpurun_out/cli_files.txt)
for v in "" "LQCOV_FAST_EXIT=0" ""; do
  for i in 1 2 3; do
    echo "== $v run $i" >> gpurun_out/cli_timeline.log
    ( time env $v longqc_b200/bin/minimap2-coverage -Y -l 0 -q 160 -k 12 -w 5 -I 4G -p 160 -t 16 $TF $QF > /dev/shm/out.tsv ) 2>> gpurun_out/cli_timeline.log
  done
done
md5sum /dev/shm/out.tsv >> gpurun_out/cli_timeline.log
grep -E "table written|Real time|^== |real" gpurun_out/cli_timeline.log | tail -60; tail -1 gpurun_out/cli_timeline.log
