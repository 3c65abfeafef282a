#!/bin/bash
# fused two-digit sort levels: parity, then the bench workload with (default) and without (LQCOV_AFB_BITW=0) the shared-memory bitmap
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_cli.py tests/test_zz_c1_consumer.py -m gpu -x -q ) > gpurun_out/pytest_sort.log 2>&1
tail -4 gpurun_out/pytest_sort.log
for m in 24576 0; do
  LQCOV_AFB_BITW=$m timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust > gpurun_out/bench_bw$m.log 2> gpurun_out/bench_bw$m.err
  python - <<PY
import json
for ln in open('gpurun_out/bench_bw$m.log'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("BITW=$m value %.3f e2e %.3f ms/step %.1f parity %s" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['parity'].get('md5')))
        for k in b['kernels'][:12]: print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
PY
  tail -3 gpurun_out/bench_bw$m.err
done
