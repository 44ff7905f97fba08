#!/bin/bash
cd "$(dirname "$0")/.."
run() { timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu --quick "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
a=d["roofline"].get("alone") or {}
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms alone %.3f ms e2e cs16 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], a.get("avg_launch_ms",0), d["e2e"]["value"]))
PY
}
run
run --ft s1_cps=2
run --ft ft_smem_kb=72
run --ft fft=2
run --ft s1_cps=2,ft_smem_kb=72
run --overlap 0
