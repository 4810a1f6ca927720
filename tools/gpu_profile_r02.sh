#!/bin/bash
# Round-2 evidence: ncu launch lists (tf32 and 3xTF32 c4 steps), ncu --set full of the dominant GEMM and attention kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for prec in tf32 tf32x3; do
  CMD="python bench.py --workload c4 --steps 2 --warmup 1 --quick --precision $prec"
  timeout 300 $CMD > gpurun_out/r02_plain_$prec.json 2>/dev/null; echo "plain $prec rc=$? $(cut -c1-120 gpurun_out/r02_plain_$prec.json)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r02_launches_c4_$prec.csv $CMD > gpurun_out/r02_ncu_$prec.log 2>&1; echo "launch list $prec rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair_kernel -s 3 -c 1 -f -o gpurun_out/r02_gemm_ffn1 python tools/gemm_bench.py "fwd ffn1" 3 > gpurun_out/r02_ncu_gemm.log 2>&1; echo "gemm rc=$?"
for k in attn_bwd_dkdv_ps attn_bwd_dq_ps attn_fwd_ps; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r02_$k env NLPS_BENCH_FAST=1 python tools/attn_bench.py 64 512 8 64 1 > gpurun_out/r02_ncu_$k.log 2>&1; echo "$k rc=$?"
done
timeout 300 python tools/gemm_bench.py > gpurun_out/r02_gemm_shapes.txt 2>&1; tail -20 gpurun_out/r02_gemm_shapes.txt
timeout 300 python tools/attn_bench.py 64 512 8 64 20 > gpurun_out/r02_attention_c4.txt 2>&1; timeout 300 python tools/attn_bench.py 8 2048 12 64 10 >> gpurun_out/r02_attention_c4.txt 2>&1; cat gpurun_out/r02_attention_c4.txt
ls -la gpurun_out/r02_*.ncu-rep
