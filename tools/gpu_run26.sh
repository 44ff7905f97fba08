#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x -k "frontend or chunking or ragged" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -6 gpurun_out/pytest_quick.log
for o in "ft_threads=256" "ft_threads=512" "ft_threads=128"; do
echo "=== $o"; B200_TRACE=1 python tools/trace_run.py nofft=1 overlap=0 $o 2>&1 | grep "b200 trace" | tail -6 | grep -E "tails"
done
python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/b.json 2>> gpurun_out/bench.err; python - <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print("-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
ncu --set full --clock-control none --import-source on -k regex:k_tail_fused -s 3 -c 1 -o gpurun_out/prof_ft4 python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_ft.log 2>&1
