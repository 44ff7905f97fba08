#!/bin/bash
# ncu captures behind profiles/r02_*: run with `gpurun -- bash tools/gpu_profile.sh`, then tools/profiles_r02.sh here (no GPU needed)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# the dominant kernel alone (stage 1), one launch, full set
ncu --set full --clock-control none --import-source on -k regex:k_xd_tma -s 3 -c 1 -o gpurun_out/r02_xd_tma -f python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_r02_xd_tma.log 2>&1
# the kernels behind stage 1
ncu --set full --clock-control none --import-source on -k regex:"k_dfir_reg|k_poly_reg|k_fir_reg|k_firr_reg|k_quad" -s 12 -c 6 -o gpurun_out/r02_tails_reg -f python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_r02_tails.log 2>&1
# the spectrum branch
ncu --set full --clock-control none --import-source on -k regex:k_fftr -s 4 -c 2 -o gpurun_out/r02_fftr -f python tools/trace_run.py overlap=0 fft_async=0 steps=5 > gpurun_out/ncu_r02_fftr.log 2>&1
# launch list of the bench command itself (per-launch device time; cold cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 4 --no-cpu --c4 0 --c3 0 > gpurun_out/ncu_r02_bench.log 2>&1
echo "=== launch timeline"; B200_TRACE=1 python tools/trace_run.py 2>&1 | grep "b200 trace" | tail -30 > gpurun_out/r02_trace.txt; tail -12 gpurun_out/r02_trace.txt
