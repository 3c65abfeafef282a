#!/bin/bash
# index radix sort with ranks computed in the count pass: parity, then the bench workload
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_index_dump.py -m gpu -x -q -k "table or dump or index" ) > gpurun_out/pytest_radix.log 2>&1
tail -4 gpurun_out/pytest_radix.log
timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust > gpurun_out/bench_rx.log 2> gpurun_out/bench_rx.err
python - <<PY
import json
for ln in open('gpurun_out/bench_rx.log'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("value %.3f e2e %.3f ms/step %.1f parity %s" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['parity'].get('md5')))
        for k in b['kernels'][:8]: print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
PY
tail -3 gpurun_out/bench_rx.err
