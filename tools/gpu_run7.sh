#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for o in "nofft=1 overlap=0 pair=0" "nofft=1 overlap=0 pair=1" "nofft=1 overlap=0 pair=1 s1_mt=128" "nofft=1 overlap=1 pair=1 s1_mt=128" "overlap=1 pair=1 s1_mt=128"; do
  echo "=== trace $o"; B200_TRACE=1 python tools/trace_run.py $o 2>&1 | grep "b200 trace" | tail -16
done > gpurun_out/trace.txt 2>&1
cat gpurun_out/trace.txt
run() { python bench.py --steps 20 --warmup 3 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f cf32 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["cf32"]["value"]))
PY
}
run --s1-mt 128
run --s1-mt 128 --overlap 0
run --s1-mt 128 --pair 0
