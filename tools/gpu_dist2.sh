#!/bin/bash
# Two-GPU visit: the sharded job against the reference goldens (NCCL), then the 2-GPU bench line.
mkdir -p gpurun_out
N=${NGPU:-2}
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py parts tandem_parts plain_pb ) > gpurun_out/dist_check_$N.log 2>&1
grep -a "dist_check\|Error\|error" gpurun_out/dist_check_$N.log | tail -8
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 ) > gpurun_out/bench_$N.log 2> gpurun_out/bench_$N.err
tail -c 1500 gpurun_out/bench_$N.err
python - <<PY
import json
for ln in open('gpurun_out/bench_$N.log'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("N", b['n_gpus'], "value %.3f e2e %.3f ms/step %.1f" % (b['value'], b['e2e']['value'], b['ms_per_step']), b['config']['parallelism'][:40], b['parity'])
        print({k: (round(v,1) if isinstance(v,float) else v) for k,v in b['stats'].items()})
PY
