#!/bin/bash
# evidence capture of the round (ncu reports stay on the box -- gpurun brings back at most 64 MiB -- only their summaries come back)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=/tmp/ncu_r02; mkdir -p $R
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 4 --no-cpu --c4 0 --c3 0 > gpurun_out/ncu_r02_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_rds_demod -s 1 -c 1 -o $R/r02_rds_demod -f python tools/rds_run.py > gpurun_out/ncu_r02_rds.log 2>&1; echo "ncu rds rc=$?"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:k_xd_tma -s 3 -c 1 -o $R/r02_xd_tma -f python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_r02_xd_tma.log 2>&1; echo "ncu xd_tma rc=$?"
python tools/ncu_summary.py $R/r02_xd_tma.ncu-rep > gpurun_out/r02_ncu_full_xd_tma.txt 2>&1
python tools/ncu_traffic.py $R/r02_xd_tma.ncu-rep 16777216 gpurun_out/r02_traffic.json > /dev/null 2>&1
python tools/ncu_summary.py $R/r02_rds_demod.ncu-rep > gpurun_out/r02_ncu_full_rds_demod.txt 2>&1
timeout 300 python tools/s1_bounds.py > gpurun_out/s1_bounds.json 2> gpurun_out/s1_bounds.err; tail -6 gpurun_out/s1_bounds.err
du -sh gpurun_out
