#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout=400 -x -k "weight_gradient or graph or deferred or delta" > gpurun_out/t_wstream.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_wstream.log
for w in 1 0 1 0; do
  echo "== c4 NLPS_WGRAD_STREAM=$w"
  NLPS_WGRAD_STREAM=$w timeout 300 python bench.py --steps 30 --warmup 5 --quick 2>> gpurun_out/bench_wstream.err | tail -1
done
for w in 1 0; do
  echo "== c5 NLPS_WGRAD_STREAM=$w"
  NLPS_WGRAD_STREAM=$w timeout 300 python bench.py --workload c5 --steps 10 --warmup 4 --quick 2>> gpurun_out/bench_wstream.err | tail -1
  echo "== c3 NLPS_WGRAD_STREAM=$w"
  NLPS_WGRAD_STREAM=$w timeout 300 python bench.py --workload c3 --steps 30 --warmup 5 --quick 2>> gpurun_out/bench_wstream.err | tail -1
done
