#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "adapter or zoom" > gpurun_out/pytest_adapter.log 2>&1; echo "pytest adapter rc=$?"; tail -8 gpurun_out/pytest_adapter.log
for n in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"; tail -c 1500 gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 2 --warmup 3 --impl reference > gpurun_out/bench_ref_n$n.json 2> gpurun_out/bench_ref_n$n.err; echo "ref n=$n rc=$?"; tail -c 600 gpurun_out/bench_ref_n$n.json
done
