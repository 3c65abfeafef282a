#!/bin/bash
# Round-end visit (r02e): the whole GPU test tier, the bench line, the reference arm, smoke, the ncu launch list of one job,
# and ncu --set full of the kernels this round changed (sketch, two-digit sort levels, radix scatter) plus the seed filter.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/bench.log').readline())
print("value %.3f e2e %.3f ms/step %.1f e2e_ms %.1f launches %d" % (b['value'], b['e2e']['value'], b['ms_per_step'], b['e2e']['ms_per_step'], b['gpu_launches']))
print("parity", b['parity']); print("roofline", b['roofline']); print("clocks", b['clocks']); print("cpu_baseline", b.get('cpu_baseline')); print("cli_e2e", b.get('cli_e2e')); print("sdust", b.get('sdust')); print("sketch_kernel", b.get('sketch_kernel'))
PY
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.log 2>&1
tail -2 gpurun_out/bench_ref.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
bash tools/gpu_launches.sh
REP=gpurun_out/prof_r02e
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"lq_sketch_pk_k|lq_af_big_k|lq_filter_count_k|lq_rs_scatter_k" --launch-count 12 \
  -f -o $REP python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-cli --no-sdust > gpurun_out/ncu_r02e.log 2>&1
tail -2 gpurun_out/ncu_r02e.log | cut -c1-200
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/prof_r02e_raw.csv 2> gpurun_out/ncu_export.err
ncu -i $REP.ncu-rep --page source --csv -k regex:"lq_sketch_pk_k" > gpurun_out/prof_r02e_src_sketch.csv 2>> gpurun_out/ncu_export.err
gzip -f gpurun_out/prof_r02e_src_sketch.csv gpurun_out/prof_r02e_raw.csv
SZ=$(stat -c %s $REP.ncu-rep); if [ "$SZ" -gt 20000000 ]; then rm -f $REP.ncu-rep; echo "rep too big ($SZ), removed"; fi
ls -la gpurun_out | grep prof_r02e
