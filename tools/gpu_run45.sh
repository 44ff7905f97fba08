#!/bin/bash
cd "$(dirname "$0")/.."
echo "=== plain tails=1"; python tools/diag_race.py tails=1 2>&1 | tail -2
echo "=== racecheck tails=1 overlap=0"; timeout 400 compute-sanitizer --tool racecheck --print-limit 10 python tools/diag_race.py overlap=0 tails=1 2>&1 | grep -E "rel rms|SUMMARY|hazard" | head -6
echo "=== memcheck tails=1"; timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python tools/diag_race.py tails=1 2>&1 | grep -E "rel rms|SUMMARY|Invalid|Error" | head -8
echo "=== racecheck tails=0"; timeout 400 compute-sanitizer --tool racecheck --print-limit 10 python tools/diag_race.py tails=0 2>&1 | grep -E "rel rms|SUMMARY|hazard" | head -6
