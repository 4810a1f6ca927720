#!/bin/bash
# One GPU-box visit: parity tests (FP32 path first, tensor-core path in separate processes so a
# faulting kernel cannot poison the others), smoke, short benches.  Logs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== fp32 tests" ; timeout 900 python -m pytest tests -m gpu -q -k "not tf32" --timeout=300 -x -q > gpurun_out/t_fp32.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_fp32.log
echo "== tf32 gemm tests" ; timeout 600 python -m pytest tests/test_gpu_tf32_gemm.py -m gpu -q --timeout=120 > gpurun_out/t_tf32_gemm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_tf32_gemm.log
echo "== tf32 model tests" ; timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -k "tf32" --timeout=120 > gpurun_out/t_tf32_model.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/t_tf32_model.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "== bench tiny fp32"; timeout 300 python bench.py --workload tiny --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tiny_fp32.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_tiny_fp32.log | cut -c1-600
echo "== bench c4 tf32"; timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_tf32.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_c4_tf32.log | cut -c1-1500
