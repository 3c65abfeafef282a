#!/bin/bash
# N-GPU bench only (gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err
python - <<PY
import json
for ln in open('gpurun_out/bench_$N.log'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("N=%d value %.3f e2e %.3f ms/step %.1f launches %d" % (b['n_gpus'], b['value'], b['e2e']['value'], b['ms_per_step'], b['gpu_launches']))
        print("parity", b['parity'])
        for k in b['kernels'][:16]: print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
        print({k: (round(v,1) if isinstance(v,float) else v) for k,v in b['stats'].items()})
PY
tail -3 gpurun_out/bench_$N.err
