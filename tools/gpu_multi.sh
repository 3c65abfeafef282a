#!/bin/bash
# N-GPU visit (gpurun --gpus N): multi-GPU parity tests, then the bench line at N (torchrun) with the 1-GPU table check
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt 2>&1
( time timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/pytest_multi.log 2>&1
tail -15 gpurun_out/pytest_multi.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err
python - <<PY
import json
for ln in open('gpurun_out/bench_$N.log'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("N=%d value %.3f e2e %.3f ms/step %.1f launches %d" % (b['n_gpus'], b['value'], b['e2e']['value'], b['ms_per_step'], b['gpu_launches']))
        print("parity", b['parity'])
        for k in b['kernels'][:26]: print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
        print({k: (round(v,1) if isinstance(v,float) else v) for k,v in b['stats'].items()})
PY
tail -5 gpurun_out/bench_$N.err
