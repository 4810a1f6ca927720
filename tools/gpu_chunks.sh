#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout=400 -x -k "weight_gradient or graph or deferred" > gpurun_out/t_chunks.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_chunks.log
for ch in 4 1 8 2 4 1; do
  echo "== c4 NLPS_VOCAB_CHUNKS=$ch"
  NLPS_VOCAB_CHUNKS=$ch timeout 300 python bench.py --steps 30 --warmup 5 --quick 2>> gpurun_out/bench_chunks.err | tail -1
done
for ch in 4 1; do
  echo "== c3 NLPS_VOCAB_CHUNKS=$ch"
  NLPS_VOCAB_CHUNKS=$ch timeout 300 python bench.py --workload c3 --steps 30 --warmup 5 --quick 2>> gpurun_out/bench_chunks.err | tail -1
done
