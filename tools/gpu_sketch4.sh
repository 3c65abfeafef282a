#!/bin/bash
# packed-key sketch with 32-record rows: parity, then the bench workload at 7 / 8 / 9 CTAs per SM (72 / 64 / 56 registers)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sketch or table" ) > gpurun_out/pytest_sketch.log 2>&1
tail -4 gpurun_out/pytest_sketch.log
for m in 8 9 7; do
  LQCOV_SKETCH_MINB=$m timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust > gpurun_out/bench_mb$m.log 2> gpurun_out/bench_mb$m.err
  python - <<PY
import json
for ln in open('gpurun_out/bench_mb$m.log'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("MINB=$m value %.3f e2e %.3f ms/step %.1f parity %s" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['parity'].get('md5')))
        for k in b['kernels'][:24]:
            if k['name'] in ('sketch','seed_sort_s40') : print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
PY
  tail -3 gpurun_out/bench_mb$m.err
done
