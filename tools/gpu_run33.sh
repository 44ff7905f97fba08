#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_adapter.py -q --tb=short -p no:cacheprovider -x -k "frontend or chunking or ragged or block or full_size or adapter" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -6 gpurun_out/pytest_quick.log
run() { python bench.py --steps 40 --warmup 5 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
}
run --s1 7
run --s1 7 --pair 0
echo "=== s1 alone"; B200_TRACE=1 python tools/trace_run.py nofft=1 overlap=0 2>&1 | grep "b200 trace" | tail -6 | grep -E "stage1"
ncu --set full --clock-control none --import-source on -k regex:k_xd_pfb -s 3 -c 1 -o gpurun_out/prof_pfb2 python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_pfb.log 2>&1
