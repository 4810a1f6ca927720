#!/bin/bash
# attention iteration visit: parity of the attention paths, timing alone
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" --timeout=200 -x > gpurun_out/t_attn.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_attn.log
timeout 120 python tools/attn_bench.py 64 512 8 64 20 2>&1 | tail -3
NLPS_ATTN_TRACE=${TRACE:-2} timeout 120 python tools/attn_bench.py 64 512 8 64 5 2>&1 | tail -22
timeout 120 python tools/attn_bench.py 8 2048 12 64 10 2>&1 | tail -3
timeout 120 python tools/attn_bench.py 16 320 8 32 20 2>&1 | tail -3
} 2>&1 | tee gpurun_out/attn.log
