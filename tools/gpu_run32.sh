#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_xd_pfb -s 3 -c 1 -o gpurun_out/prof_pfb1 python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_pfb.log 2>&1
ls -la gpurun_out/prof_pfb1.ncu-rep
