#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { python bench.py --steps 40 --warmup 5 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f (%.3f ms) cf32 %.0f (%.3f ms)" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step_wall"], d["e2e"]["cf32"]["value"], d["e2e"]["cf32"]["ms_per_step_wall"]))
PY
}
run
run
tail -3 gpurun_out/bench.err
