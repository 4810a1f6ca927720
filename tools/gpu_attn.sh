#!/bin/bash
# attention iteration visit: parity of the attention paths, timing alone, short c4 bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -k "attention or tf32" --timeout=200 > gpurun_out/t_attn.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/t_attn.log
timeout 120 python tools/attn_bench.py 64 512 8 64 20 2>&1 | tail -3
NLPS_ATTN_TRACE=1 timeout 120 python tools/attn_bench.py 64 512 8 64 5 2>&1 | tail -12
timeout 120 python tools/attn_bench.py 8 2048 12 64 10 2>&1 | tail -3
timeout 120 python tools/attn_bench.py 16 320 8 32 20 2>&1 | tail -3
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_tf32.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_c4_tf32.log | cut -c1-400
