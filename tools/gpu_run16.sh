#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_xd_pipe -s 4 -c 1 -o gpurun_out/prof_pipe3 python tools/trace_run.py nofft=1 overlap=0 pair=1 s1=3 > gpurun_out/ncu_pipe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fft_p1|k_fft_p2" -s 4 -c 2 -o gpurun_out/prof_fft3 python tools/trace_run.py overlap=0 pair=1 s1=3 fft_async=0 > gpurun_out/ncu_fft.log 2>&1
ls -la gpurun_out/*.ncu-rep
