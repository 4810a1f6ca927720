#!/bin/bash
# Round-2 evidence after the GEMM producer change: launch lists (tf32, tf32x3), ncu --set full of the FFN GEMM and of a weight-gradient GEMM, shape table.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for prec in tf32 tf32x3; do
  CMD="python bench.py --workload c4 --steps 2 --warmup 1 --quick --precision $prec"
  timeout 300 $CMD > gpurun_out/r02_plain_$prec.json 2>/dev/null; echo "plain $prec rc=$? $(cut -c1-120 gpurun_out/r02_plain_$prec.json)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c4_$prec.csv $CMD > gpurun_out/r02_ncu_$prec.log 2>&1; echo "launch list $prec rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair_kernel -s 3 -c 1 -f -o gpurun_out/r02_gemm_ffn1 python tools/gemm_bench.py "fwd ffn1" 3 > gpurun_out/r02_ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair_kernel -s 3 -c 1 -f -o gpurun_out/r02_gemm_dw1 python tools/gemm_bench.py "wgrad dW1" 3 > gpurun_out/r02_ncu_gemm_dw1.log 2>&1; echo "gemm dw1 rc=$?"
timeout 300 python tools/gemm_bench.py > gpurun_out/r02_gemm_shapes.txt 2>&1; tail -16 gpurun_out/r02_gemm_shapes.txt
