#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for o in "ft_threads=256" "ft_threads=512" "ft_threads=256 ft_ob=600" "tails=1"; do
echo "=== $o"; B200_TRACE=1 python tools/trace_run.py nofft=1 overlap=0 $o 2>&1 | grep "b200 trace" | tail -6 | grep -E "tails|stage1"
done
echo "=== full (fft on, overlap on)"; B200_TRACE=1 python tools/trace_run.py 2>&1 | grep "b200 trace" | tail -16
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
