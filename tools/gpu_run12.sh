#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 40 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/bench.json"))
print("value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f cf32 %.0f cpu %.1f launches %d" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["cf32"]["value"], d["cpu_baseline"]["value"], d["gpu_launches"]))
PY
B200_TRACE=1 python tools/trace_run.py overlap=1 pair=1 2>&1 | grep "b200 trace" | tail -9
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
