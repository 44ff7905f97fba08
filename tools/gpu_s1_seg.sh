#!/bin/bash
# stage 1 with one ring slot per 128-byte segment of a tile (B200_S1_SEG=1, the default) against one slot per tile: the whole
# GPU suite, then the bench line both ways and with the deeper ring
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_seg.log 2>&1; echo "pytest(seg=1) rc=$?"; tail -8 gpurun_out/pytest_gpu_seg.log
for cfg in "0:2" "1:2" "0:3" "1:3" "1:2" "0:2"; do
  IFS=: read v st <<< "$cfg"
  B200_S1_SEG=$v timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu --c3 0 --c4 0 --s1-stages $st > gpurun_out/bench_seg$v$st.json 2> gpurun_out/bench_seg$v$st.err; echo "bench seg=$v stages=$st rc=$?"; tail -2 gpurun_out/bench_seg$v$st.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_seg$v$st.json"))
    cs=d["config"]["chunk_sweep"]
    r=d["roofline"]
    print("seg=$v stages=$st value=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), {k:round(x["value"]) for k,x in cs.items()}, [(g["group"],round(g["avg_ms"]*1e3,1)) for g in r["by_group"]], "alone us", round(r["alone"]["avg_launch_ms"]*1e3,1), "frac in situ %.3f alone %.3f"%(r["frac"], r["alone"]["frac"]))
except Exception as e:
    print("failed", e)
PY
done
