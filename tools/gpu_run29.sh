#!/bin/bash
cd "$(dirname "$0")/.."
B200_FT_CLOCKS=1 python tools/trace_run.py nofft=1 overlap=0 steps=5 2>&1 | grep "ft clocks" | tail -2
B200_FT_CLOCKS=1 python tools/trace_run.py nofft=1 overlap=0 steps=5 ft_threads=128 2>&1 | grep "ft clocks" | tail -1
