#!/bin/bash
# programmatic dependent launch of the chain behind stage 1: parity of the whole GPU suite with it on, then the bench line
# (device-resident value, chunk sweep) with it off and on
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_PDL=1 timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --durations=12 > gpurun_out/pytest_gpu_pdl.log 2>&1; echo "pytest(pdl) rc=$?"; tail -22 gpurun_out/pytest_gpu_pdl.log
for v in 0 1 0 1; do
  B200_PDL=$v timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu --c3 0 --c4 0 > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err; echo "bench pdl=$v rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pdl$v.json"))
cs=d["config"]["chunk_sweep"]
print("pdl=$v value=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), {k:round(x["value"]) for k,x in cs.items()}, [(g["group"],round(g["avg_ms"]*1e3,1)) for g in d["roofline"]["by_group"]], "alone", round(d["roofline"]["alone"]["avg_launch_ms"]*1e3,1))
PY
done
