#!/bin/bash
# one pytest process per GPU test: a sticky CUDA error in one test cannot poison the others
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
./build/ubench > gpurun_out/ubench.txt 2>&1
python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" > gpurun_out/tests.txt
: > gpurun_out/pytest_each.log
while read -r t; do
  echo "=== $t" >> gpurun_out/pytest_each.log
  timeout 300 python -m pytest "$t" -q -x --tb=short -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/pytest_each.log
done < gpurun_out/tests.txt
grep -E "^(===|FAILED|ERROR)|passed|failed" gpurun_out/pytest_each.log | tail -150
