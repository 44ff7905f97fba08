#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for v in 0 2; do python bench.py --steps 10 --warmup 3 --no-cpu --s1 $v > gpurun_out/bench_s1_$v.json 2>> gpurun_out/bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_xd_tile -s 3 -c 2 -o gpurun_out/prof_xd python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
