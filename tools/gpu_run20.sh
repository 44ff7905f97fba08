#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_tail_fused -s 3 -c 1 -o gpurun_out/prof_ft1 python tools/trace_run.py nofft=1 overlap=0 tails=2 ft_threads=256 steps=5 > gpurun_out/ncu_ft.log 2>&1
ls -la gpurun_out/prof_ft1.ncu-rep
