#!/bin/bash
# full validation on one B200: every GPU test, smoke(), the default bench line, a launch timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench_default.json
echo "=== full trace of one step"
B200_TRACE=1 python tools/trace_run.py overlap=1 pair=1 s1=6 2>&1 | grep "b200 trace" | tail -40
