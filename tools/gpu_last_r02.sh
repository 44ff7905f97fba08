#!/bin/bash
# last GPU call of the round: evidence first (ncu reports stay on the box, only their summaries come back), then the split
# stage 1 (B200_S1_SPLIT=1): whole GPU suite with it on, bench line both ways
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=/tmp/ncu_r02; mkdir -p $R
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 4 --no-cpu --c4 0 --c3 0 > gpurun_out/ncu_r02_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_rds_demod -s 1 -c 1 -o $R/r02_rds_demod -f python tools/rds_run.py > gpurun_out/ncu_r02_rds.log 2>&1; echo "ncu rds rc=$?"
for sp in 0 1; do
  B200_S1_SPLIT=$sp timeout 90 ncu --set full --clock-control none --import-source on -k regex:k_xd_tma -s 3 -c 1 -o $R/r02_xd_tma_split$sp -f python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_r02_xd_tma$sp.log 2>&1; echo "ncu xd_tma split=$sp rc=$?"
  python tools/ncu_summary.py $R/r02_xd_tma_split$sp.ncu-rep > gpurun_out/r02_ncu_full_xd_tma_split$sp.txt 2>&1
  python tools/ncu_traffic.py $R/r02_xd_tma_split$sp.ncu-rep 16777216 gpurun_out/r02_traffic_split$sp.json > /dev/null 2>&1
done
python tools/ncu_summary.py $R/r02_rds_demod.ncu-rep > gpurun_out/r02_ncu_full_rds_demod.txt 2>&1
grep -h "gpu__time_duration.sum\|Kernel Name" gpurun_out/r02_ncu_full_xd_tma_split*.txt
B200_S1_SPLIT=1 timeout 400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu_split.log 2>&1; echo "pytest(split) rc=$?"; tail -8 gpurun_out/pytest_gpu_split.log
for v in 0 1; do
  B200_S1_SPLIT=$v timeout 120 python bench.py --steps 12 --warmup 3 --no-cpu --c3 0 --c4 0 > gpurun_out/bench_split$v.json 2> gpurun_out/bench_split$v.err; echo "bench split=$v rc=$?"; tail -2 gpurun_out/bench_split$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_split$v.json"))
    cs=d["config"]["chunk_sweep"]
    r=d["roofline"]
    print("split=$v value=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), {k:round(x["value"]) for k,x in cs.items()}, [(g["group"],round(g["avg_ms"]*1e3,1)) for g in r["by_group"]], "alone us", round(r["alone"]["avg_launch_ms"]*1e3,1), "frac in situ %.3f alone %.3f"%(r["frac"], r["alone"]["frac"]))
except Exception as e:
    print("failed", e)
PY
done
du -sh gpurun_out
