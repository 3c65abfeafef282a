#!/bin/bash
# Round 2, first visit: the whole GPU test tier (new goldens: covt_gate, ultralong, c5_small, many_targets) and the bench line with cli_e2e + sdust.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.log').readline())
print("value %.3f e2e %.3f ms/step %.1f e2e_ms %.1f launches %d" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['e2e']['ms_per_step'], b['gpu_launches']))
print("cpu", b['cpu_baseline']); print("cli", b['cli_e2e']); print("sdust", b['sdust']); print("parity", b['parity']); print("roofline", b['roofline'])
for k in b['kernels'][:24]: print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
print({k: (round(v,1) if isinstance(v,float) else v) for k,v in b['stats'].items()})
PY
tail -3 gpurun_out/bench.err
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 3 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
cut -c1-600 gpurun_out/bench_ref.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' ) 2>&1 | tail -2
bash tools/gpu_launches.sh
