#!/bin/bash
# Round evidence in one visit: full bench line, ncu launch list of the c4 step, ncu --set full of the dominant kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --workload c4 --steps 20 --warmup 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_c4.json
timeout 300 python tools/attn_bench.py 64 512 8 64 20 > gpurun_out/attention_c4.txt 2>&1; timeout 300 python tools/attn_bench.py 8 2048 12 64 10 >> gpurun_out/attention_c4.txt 2>&1; timeout 300 python tools/attn_bench.py 16 320 8 32 20 >> gpurun_out/attention_c4.txt 2>&1; cat gpurun_out/attention_c4.txt; bash tools/gpu_attn_split.sh > gpurun_out/attention_split_c4.txt 2>&1; cat gpurun_out/attention_split_c4.txt
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_shapes.txt 2>&1; tail -3 gpurun_out/gemm_shapes.txt
CMD="python bench.py --workload c4 --steps 2 --warmup 1 --quick"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_c4.log 2>&1; echo "launch list rc=$?"
for k in attn_bwd_dkdv_ps attn_bwd_dq_ps attn_fwd_ps; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/$k env NLPS_BENCH_FAST=1 python tools/attn_bench.py 64 512 8 64 1 > gpurun_out/ncu_$k.log 2>&1; echo "$k rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair_kernel -s 3 -c 1 -f -o gpurun_out/gemm_ffn1 python tools/gemm_bench.py "fwd ffn1" 3 > gpurun_out/ncu_gemm.log 2>&1; echo "gemm rc=$?"
bash tools/gpu_attn_fwd_split.sh >> gpurun_out/attention_split_c4.txt 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -2 gpurun_out/pytest_gpu.log
