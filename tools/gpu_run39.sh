#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_rank.py -q --tb=short -p no:cacheprovider -m gpu > gpurun_out/pytest_shard.log 2>&1; echo "shard rc=$?"; tail -15 gpurun_out/pytest_shard.log
ncu --set full --clock-control none --import-source on -k regex:k_xd_pfb -s 3 -c 1 -o gpurun_out/prof_pfb3 python tools/trace_run.py nofft=1 overlap=0 steps=5 > gpurun_out/ncu_pfb.log 2>&1
