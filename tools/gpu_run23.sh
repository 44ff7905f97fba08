#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -k "fft or c1_geometry or variants_100msps or full_size or int16" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -12 gpurun_out/pytest_quick.log
run() { python bench.py --steps 40 --warmup 5 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
}
run --tails 2
ncu --set full --clock-control none --import-source on -k regex:"k_fftr" -s 4 -c 2 -o gpurun_out/prof_fftr python tools/trace_run.py overlap=0 pair=1 fft_async=0 steps=5 > gpurun_out/ncu_fft.log 2>&1
B200_TRACE=1 python tools/trace_run.py overlap=0 fft_async=0 steps=4 2>&1 | grep "b200 trace" | tail -14
