#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x -k "c1_geometry and 6 or ragged or chunking or mixed_modes" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -15 gpurun_out/pytest_quick.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q --tb=line -p no:cacheprovider -x -k "ragged or chunking" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -c "Invalid" gpurun_out/sanitizer.log; grep -m5 -A12 "Invalid" gpurun_out/sanitizer.log; tail -3 gpurun_out/sanitizer.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
run() { python bench.py --steps 40 --warmup 5 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
}
run --tails 1
run --tails 2
run --tails 2 --ft ft_threads=256
run --tails 2 --ft ft_smem_kb=72
run --tails 2 --ft ft_smem_kb=72,ft_threads=256
echo "=== trace"
B200_TRACE=1 python tools/trace_run.py nofft=1 overlap=0 tails=2 2>&1 | grep "b200 trace" | tail -12
B200_TRACE=1 python tools/trace_run.py nofft=1 overlap=0 tails=2 ft_threads=256 2>&1 | grep "b200 trace" | tail -6
