#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -x -k "frontend or chunking or ragged" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -6 gpurun_out/pytest_quick.log
for o in "ft_threads=256" "ft_threads=512" "ft_threads=128" "ft_threads=512 ft_smem_kb=72"; do
echo "=== $o"; B200_TRACE=1 python tools/trace_run.py nofft=1 overlap=0 tails=2 $o steps=6 2>&1 | grep "b200 trace" | grep -E "tails" | tail -4
done
run() { python bench.py --steps 40 --warmup 5 --no-cpu "$@" > gpurun_out/b.json 2>> gpurun_out/bench.err; python - "$@" <<PY
import json,sys
d=json.load(open("gpurun_out/b.json"))
print(" ".join(sys.argv[1:]), "-> value %.0f MS/s step %.3f ms  s1 %.3f ms frac %.3f  e2e cs16 %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
}
run --ft ft_threads=256
run --ft ft_threads=512
run --ft ft_threads=512,fft=2
run --ft ft_threads=256,fft=2
