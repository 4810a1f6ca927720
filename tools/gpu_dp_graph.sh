#!/bin/bash
# 2-GPU visit: deferred-step graph tests, then same-box A/B of the data-parallel step with and without the graph.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests"; timeout 500 python -m pytest tests/test_gpu_dp_nccl.py "tests/test_gpu_model.py::test_cuda_graph_of_the_deferred_step_with_external_bucket_events_is_bitwise_the_eager_one" -m gpu -q --timeout=400 -x > gpurun_out/t_dp_graph.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_dp_graph.log
for g in 1 0 1 0; do
  echo "== 2gpu bench NLPS_GRAPH_DP=$g"
  NLPS_GRAPH_DP=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --quick 2> gpurun_out/bench_2gpu_g$g.err | tail -1 | tee -a gpurun_out/bench_2gpu_ab.jsonl
done
echo "== 1gpu bench"; timeout 300 python bench.py --steps 20 --warmup 5 --quick 2> gpurun_out/bench_1gpu_q.err | tail -1 | tee -a gpurun_out/bench_2gpu_ab.jsonl
