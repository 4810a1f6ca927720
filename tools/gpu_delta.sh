#!/bin/bash
# 1-GPU visit: delta-fusion and deferred-graph tests, then same-box A/B of the c4 step with and without the fused delta.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests"; timeout 500 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout=400 -k "delta or deferred or tf32" > gpurun_out/t_delta.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_delta.log
for f in 1 0 1 0; do
  echo "== c4 NLPS_FUSE_DELTA=$f"
  NLPS_FUSE_DELTA=$f timeout 300 python bench.py --steps 30 --warmup 5 --quick 2> gpurun_out/bench_delta_$f.err | tail -1 | tee -a gpurun_out/bench_delta_ab.jsonl
done
for f in 1 0; do
  echo "== c5 NLPS_FUSE_DELTA=$f"
  NLPS_FUSE_DELTA=$f timeout 300 python bench.py --workload c5 --steps 10 --warmup 4 --quick 2>> gpurun_out/bench_delta_$f.err | tail -1 | tee -a gpurun_out/bench_delta_ab.jsonl
done
