#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/diag_shard.py 2>&1 | head -4
timeout 900 python -m pytest tests/test_multi_rank.py tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider -m gpu -k "sharding or many_vfos or variants_100msps" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -5 gpurun_out/pytest_quick.log
