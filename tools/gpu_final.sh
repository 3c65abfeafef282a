#!/bin/bash
# Round-end visit: the whole GPU test tier, the bench line, the reference arm, the ncu launch list of one job.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.log').readline())
print("value %.3f e2e %.3f ms/step %.1f e2e_ms %.1f launches %d" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['e2e']['ms_per_step'], b['gpu_launches']))
print("parity", b['parity']); print("roofline", b['roofline']); print("clocks", b['clocks'])
PY
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.log 2>&1
tail -2 gpurun_out/bench_ref.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
bash tools/gpu_launches.sh
