#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x -k "frontend or chunking or ragged" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -6 gpurun_out/pytest_quick.log
B200_FT_CLOCKS=1 python tools/trace_run.py nofft=1 overlap=0 steps=5 2>&1 | grep "ft clocks" | tail -3
B200_FT_CLOCKS=1 python tools/trace_run.py nofft=1 overlap=0 steps=5 ft_threads=512 2>&1 | grep "ft clocks" | tail -2
for o in "ft_threads=256" "ft_threads=512"; do
echo "=== $o"; B200_TRACE=1 python tools/trace_run.py nofft=1 overlap=0 $o 2>&1 | grep "b200 trace" | tail -6 | grep -E "tails"
done
python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/b.json 2>> gpurun_out/bench.err; python - <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print("-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f cf32 %.0f cs8 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["cf32"]["value"], d["e2e"]["cs8"]["value"]))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
