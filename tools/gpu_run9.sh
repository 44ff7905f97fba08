#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_xd_pipe -s 4 -c 1 -o gpurun_out/prof_pipe_paired python tools/trace_run.py nofft=1 overlap=0 pair=1 s1=3 > gpurun_out/ncu_pipe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fir_cd|k_poly2|k_fir_c2|k_fir_r2" -s 8 -c 5 -o gpurun_out/prof_tails python tools/trace_run.py nofft=1 overlap=0 pair=1 s1=3 > gpurun_out/ncu_tails.log 2>&1
ls -la gpurun_out/*.ncu-rep
