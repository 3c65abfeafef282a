#!/bin/bash
# walk statistics + timing of the bench workload at N GPUs (1: plain bench; >1: torchrun)
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  LQCOV_WALK_STATS=1 timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust --steps 1 --warmup 0 > gpurun_out/ws.log 2> gpurun_out/ws.err
  grep "walk stats" gpurun_out/ws.err | sort | uniq -c | sort -rn | head -20
  timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust > gpurun_out/bench.log 2> gpurun_out/bench.err
  F=gpurun_out/bench.log
else
  LQCOV_WALK_STATS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 1 --warmup 0 > gpurun_out/ws.log 2> gpurun_out/ws.err
  grep "walk stats" gpurun_out/ws.err | sort | uniq -c | sort -rn | head -20
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err
  F=gpurun_out/bench_$N.log
fi
python - <<PY
import json
for ln in open('$F'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("N=%d value %.3f e2e %.3f ms/step %.1f" % (b['n_gpus'], b['value'], b['e2e']['value'], b['ms_per_step']))
        print("parity", b['parity'])
        for k in b['kernels'][:14]: print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
PY
