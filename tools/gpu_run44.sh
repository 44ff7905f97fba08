#!/bin/bash
cd "$(dirname "$0")/.."
for o in "ft_direct=0" "ft_direct=1" "tails=1" "s1=6 ft_direct=0"; do
  echo "=== racecheck $o"
  timeout 400 compute-sanitizer --tool racecheck --print-limit 10 python tools/diag_race.py $o > gpurun_out/race_$$.log 2>&1
  grep -E "rel rms|RACECHECK SUMMARY|hazard" gpurun_out/race_$$.log | head -8
done
echo "=== plain"; python tools/diag_race.py 2>&1 | tail -2
