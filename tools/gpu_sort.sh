#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "A=1" "NLPS_EARLY_SORT=0" "NLPS_WGRAD_PRIO=same" "A=1" "NLPS_EARLY_SORT=0" "NLPS_WGRAD_PRIO=same"; do
  echo "== c4 $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --quick 2>> gpurun_out/bench_sort.err | tail -1
done
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
