#!/bin/bash
# sketch kernels: parity of the three forms, then the bench workload with each of them
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/pytest_parity.log 2>&1
tail -5 gpurun_out/pytest_parity.log
for m in 1 2 0; do
  LQCOV_SKETCH_PK=$m timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust > gpurun_out/bench_pk$m.log 2> gpurun_out/bench_pk$m.err
  python - <<PY
import json
for ln in open('gpurun_out/bench_pk$m.log'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("PK=$m value %.3f e2e %.3f ms/step %.1f parity %s" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['parity'].get('md5')))
        for k in b['kernels'][:16]:
            if k['name'] in ('sketch','pack') : print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
PY
  tail -3 gpurun_out/bench_pk$m.err
done
