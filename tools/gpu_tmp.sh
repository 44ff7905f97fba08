#!/bin/bash
cd "$(dirname "$0")/.."
for i in 1 2 3 4 5 6; do
timeout 300 python -m pytest tests/test_gpu_parity.py -q --tb=line -p no:cacheprovider -m gpu -k "af_chain" 2>&1 | tail -2 | head -1
done
