#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
run() { python bench.py --steps 40 --warmup 5 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f (%.3f ms) cf32 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step_wall"], d["e2e"]["cf32"]["value"]))
PY
}
run
run --no-clocks
run --overlap 0
run --offsets asym
B200_TRACE=1 python tools/trace_run.py overlap=1 pair=1 2>&1 | grep "b200 trace" | tail -9
