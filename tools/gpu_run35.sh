#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x -k "variants_100msps or c1_geometry or full_size" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -6 gpurun_out/pytest_quick.log
for o in "s1=7"; do
echo "=== $o"; timeout 120 env B200_TRACE=1 python tools/trace_run.py nofft=1 overlap=0 $o 2>&1 | grep "b200 trace" | tail -6 | grep -E "stage1"
done
run() { timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
a=d["roofline"].get("alone") or {}
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f alone %.3f ms frac %.3f e2e cs16 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], a.get("avg_launch_ms",0), a.get("frac",0), d["e2e"]["value"]))
PY
}

run --s1 7
