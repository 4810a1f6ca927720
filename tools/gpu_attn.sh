#!/bin/bash
# attention iteration visit: parity of the attention paths, timing alone (ss = first-generation kernels as the box-speed reference)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" --timeout=200 -x > gpurun_out/t_attn.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_attn.log
for shape in "64 512 8 64 20" "8 2048 12 64 10" "16 320 8 32 20"; do
  echo "ts: $(timeout 120 python tools/attn_bench.py $shape 2>&1 | tail -1)"
  true
done
NLPS_ATTN_TRACE=${TRACE:-2} timeout 120 python tools/attn_bench.py 64 512 8 64 5 2>&1 | tail -32
} 2>&1 | tee gpurun_out/attn.log
