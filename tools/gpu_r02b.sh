#!/bin/bash
# round 2, second session: new RDS demodulator tests first, then the whole GPU suite, smoke, the default bench line, and a
# small-chunk probe of the fused tail against the register chain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rds.py tests/test_host_adapter.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_rds.log 2>&1; echo "rds rc=$?"; tail -25 gpurun_out/pytest_rds.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 6000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench.err
timeout 300 python tools/small_chunk_variants.py > gpurun_out/small_chunk_variants.json 2> gpurun_out/small_chunk_variants.err; echo "probe rc=$?"; cat gpurun_out/small_chunk_variants.json | head -60
