#!/bin/bash
# One GPU-box visit: parity tests, bench line, reference arm, ncu launch list. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -c 6000 gpurun_out/bench.log
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.log 2>&1
tail -3 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches.csv
