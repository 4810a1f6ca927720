#!/bin/bash
# 2-GPU visit: data-parallel tests, then same-box A/B of the data-parallel step (graph / weight-gradient stream switches).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests"; timeout 500 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q --timeout=400 -x > gpurun_out/t_dp_graph.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_dp_graph.log
for cfg in "A=1" "NLPS_WGRAD_STREAM=0" "NLPS_GRAPH_DP=0" "A=1"; do
  echo "== 2gpu bench $cfg"
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --quick 2>> gpurun_out/bench_2gpu_ab.err | tail -1 | tee -a gpurun_out/bench_2gpu_ab.jsonl
done
echo "== 2gpu full line"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -1 gpurun_out/bench_2gpu.json | cut -c1-200
echo "== 1gpu bench"; timeout 300 python bench.py --steps 20 --warmup 5 --quick 2> gpurun_out/bench_1gpu_q.err | tail -1 | tee -a gpurun_out/bench_2gpu_ab.jsonl
