#!/bin/bash
# end-of-round validation and evidence: every GPU test, smoke, the default bench line; what bounds stage 1; ncu captures
# (launch list of the bench command, full sets of stage 1, the chain behind it and the RDS demodulator)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 1500 gpurun_out/bench_default.json; echo; tail -3 gpurun_out/bench.err
timeout 300 python tools/s1_bounds.py > gpurun_out/s1_bounds.json 2> gpurun_out/s1_bounds.err; echo "s1_bounds rc=$?"; cat gpurun_out/s1_bounds.err | tail -14
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_xd_tma -s 3 -c 1 -o gpurun_out/r02_xd_tma -f python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_r02_xd_tma.log 2>&1; echo "ncu xd_tma rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_dfir_reg|k_poly_reg|k_fir_reg|k_firr_reg|k_quad" -s 12 -c 6 -o gpurun_out/r02_tails_reg -f python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_r02_tails.log 2>&1; echo "ncu tails rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_rds_demod -s 1 -c 1 -o gpurun_out/r02_rds_demod -f python tools/rds_run.py > gpurun_out/ncu_r02_rds.log 2>&1; echo "ncu rds rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --chunks-per-step 4 --no-cpu --c4 0 --c3 0 > gpurun_out/ncu_r02_bench.log 2>&1; echo "ncu launches rc=$?"
B200_TRACE=1 timeout 120 python tools/trace_run.py 2>&1 | grep "b200 trace" | tail -30 > gpurun_out/r02_trace.txt; tail -4 gpurun_out/r02_trace.txt
ls -la gpurun_out | head -40
