#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for o in "nofft=1 overlap=0 pair=1 s1=3" "nofft=1 overlap=0 pair=1 s1=5" "nofft=1 overlap=0 pair=1 s1=6" "nofft=1 overlap=0 pair=0 s1=6" "overlap=0 pair=1 s1=3 fft_async=0"; do
  echo "=== trace $o"; B200_TRACE=1 python tools/trace_run.py $o 2>&1 | grep "b200 trace" | tail -7 | grep -E "stage1|outputs"
done
run() { python bench.py --steps 40 --warmup 5 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f cf32 %.0f cs8 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["cf32"]["value"], d["e2e"]["cs8"]["value"]))
PY
}
run --s1 3
run --s1 5
run --s1 6
