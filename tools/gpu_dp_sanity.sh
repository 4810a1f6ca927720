#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests"; timeout 300 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q --timeout=280 -x > gpurun_out/t_dp_graph.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_dp_graph.log
echo "== 2gpu full line"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"; tail -1 gpurun_out/bench_2gpu.json | cut -c1-220
