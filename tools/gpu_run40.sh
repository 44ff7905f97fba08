#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_rank.py -q --tb=short -p no:cacheprovider -m gpu -k "compressed or sharding or int16 or c1_geometry or fft_line" > gpurun_out/pytest_quick.log 2>&1; echo "quick rc=$?"; tail -15 gpurun_out/pytest_quick.log
