#!/bin/bash
# kernel iteration visit: the sort/seed parity tests, then a short bench without the CPU/CLI/sdust legs
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_cli.py tests/test_zz_c1_consumer.py -m gpu -x -q -k "${QUICK_K:-abi_table or seeds_and_sort or many_targets or small_seed}" ) > gpurun_out/pytest_quick.log 2>&1
tail -4 gpurun_out/pytest_quick.log
( time timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust ${BENCH_ARGS} ) > gpurun_out/bench.log 2> gpurun_out/bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.log').readline())
print("value %.3f e2e %.3f ms/step %.1f e2e_ms %.1f launches %d" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['e2e']['ms_per_step'], b['gpu_launches']))
print("parity", b['parity'])
for k in b['kernels'][:30]: print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
print({k: (round(v,1) if isinstance(v,float) else v) for k,v in b['stats'].items()})
PY
tail -3 gpurun_out/bench.err
