#!/bin/bash
# one-variable experiments on the bench workload: VAR=name VALUES="a b c" bash tools/gpu_tune.sh
mkdir -p gpurun_out
VAR=${VAR:-LQCOV_AFB_CTAS}
for v in ${VALUES:-1 2 4 8}; do
  env $VAR=$v timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust > gpurun_out/tune_$v.log 2> gpurun_out/tune_$v.err
  python - <<PY
import json
b=json.loads(open('gpurun_out/tune_$v.log').readline())
ks={k['name']:k['ms_per_step'] for k in b['kernels']}
print("$VAR=$v  ms/step %.1f  value %.3f  parity %s" % (b['ms_per_step'], b['value'], b['parity'].get('table_md5','')[:8] if isinstance(b['parity'],dict) else ''))
print("   ", " ".join("%s=%.2f" % (k, ks[k]) for k in sorted(ks) if k.startswith('seed_sort') or k.startswith('seed_walk') or k.startswith('seed_place')))
PY
done
