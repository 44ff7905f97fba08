#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/diag_shard.py 2>&1 | tail -32
