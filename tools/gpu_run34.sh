#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_tail_fused -s 3 -c 1 -o gpurun_out/prof_ft5 python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_ft.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "=== full trace"; B200_TRACE=1 python tools/trace_run.py 2>&1 | grep "b200 trace" | tail -24
