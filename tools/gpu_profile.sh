#!/bin/bash
# ncu captures behind profiles/: run with `gpurun -- bash tools/gpu_profile.sh`, then tools/ncu_summary.py on gpurun_out/*.ncu-rep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for k in k_xd_pfb k_tail_fused; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_$k python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_fftr -s 4 -c 2 -o gpurun_out/prof_k_fftr python tools/trace_run.py overlap=0 fft_async=0 steps=5 > gpurun_out/ncu_k_fftr.log 2>&1
# launch list of the bench command itself (per-launch device time, cold cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "=== launch timeline"; B200_TRACE=1 python tools/trace_run.py 2>&1 | grep "b200 trace" | tail -24
