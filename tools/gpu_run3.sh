#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
python tools/diag_e2e.py > gpurun_out/diag_e2e.txt 2>&1; cat gpurun_out/diag_e2e.txt
for v in 3 4; do python bench.py --steps 20 --warmup 3 --no-cpu --s1 $v > gpurun_out/bench_s1_$v.json 2>> gpurun_out/bench.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_s1_$v.json"))
print("s1=$v value %.0f MS/s step %.3f ms  s1 %.3f ms  e2e cs16 %.0f cf32 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["e2e"]["value"], d["e2e"]["cf32"]["value"]))
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/bench.err
