#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 500 python -m pytest tests/test_gpu_model.py tests/test_gpu_tf32_gemm.py -m gpu -q --timeout=400 -k "delta or tf32" > gpurun_out/t_prefetch.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_prefetch.log
for d in 0 48 16 32 0 48; do
  echo "== c4 NLPS_GEMM_DEBUG=$d"
  NLPS_GEMM_DEBUG=$d timeout 300 python bench.py --steps 30 --warmup 5 --quick 2>> gpurun_out/bench_prefetch.err | tail -1
done
NLPS_GEMM_DEBUG=0 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_shapes.txt 2>&1; tail -18 gpurun_out/gemm_shapes.txt
