#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 compute-sanitizer --tool racecheck --print-limit 30 python -m pytest tests/test_gpu_parity.py -q --tb=line -p no:cacheprovider -x -k "c1_geometry and 7 or many_vfos" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -c "Race reported\|hazard" gpurun_out/racecheck.log; grep -m12 -B2 -A10 "hazard" gpurun_out/racecheck.log | head -80; tail -5 gpurun_out/racecheck.log
