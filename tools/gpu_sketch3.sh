#!/bin/bash
# warp-tile packed-key sketch: parity (library + executables), bench with bulk copies (1) / plain loads (2), ncu --set full
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k sketch ) > gpurun_out/pytest_sketch.log 2>&1
tail -4 gpurun_out/pytest_sketch.log
for m in 1; do
  LQCOV_SKETCH_PK=$m timeout 600 python bench.py --no-cpu-baseline --no-cli --no-sdust > gpurun_out/bench_pk$m.log 2> gpurun_out/bench_pk$m.err
  python - <<PY
import json
for ln in open('gpurun_out/bench_pk$m.log'):
    if ln.startswith('{'):
        b=json.loads(ln)
        print("PK=$m value %.3f e2e %.3f ms/step %.1f parity %s" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['parity'].get('md5')))
        for k in b['kernels'][:24]:
            if k['name'] in ('sketch','pack') : print("  %-22s %8.3f ms  %5.1f%%  %7.1f GB/s" % (k['name'], k['ms_per_step'], 100*k['share'], k['achieved_gbs']))
PY
  tail -3 gpurun_out/bench_pk$m.err
done
for m in 1; do
  REP=gpurun_out/prof_sketch_pkw$m
  LQCOV_SKETCH_PK=$m timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lq_sketch_pk_k" --launch-skip 2 --launch-count 2 \
    -f -o $REP python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-cli --no-sdust > gpurun_out/ncu_sketch_pkw$m.log 2>&1
  tail -2 gpurun_out/ncu_sketch_pkw$m.log | cut -c1-200
  ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/prof_sketch_pkw${m}_raw.csv 2> gpurun_out/ncu_export.err
  ncu -i $REP.ncu-rep --page source --csv > gpurun_out/prof_sketch_pkw${m}_src.csv 2>> gpurun_out/ncu_export.err
  gzip -f gpurun_out/prof_sketch_pkw${m}_src.csv
  SZ=$(stat -c %s $REP.ncu-rep); if [ "$SZ" -gt 20000000 ]; then rm -f $REP.ncu-rep; echo "rep too big ($SZ), removed"; fi
done
ls -la gpurun_out | grep prof_sketch_pkw
