#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_TRACE=1 python tools/trace_run.py host=1 steps=10 2>&1 | grep -E "b200 trace|numa" | tail -40
