#!/bin/bash
# Round-end evidence in one 1-GPU visit: bench line, reference arm, ncu launch list, ncu DRAM metrics of the
# bandwidth kernels, the GPU test suite and smoke().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_c4.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
CMD="python bench.py --workload c4 --steps 2 --warmup 1 --quick"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c4.csv $CMD > gpurun_out/ncu_c4.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none \
  -k regex:"ln_fwd_kernel|ln_bwd_kernel|xent|adam_kernel|sumsq_partial|colsum_partial|attn_delta|embed_fwd|scatter_sum|splitk_reduce" -c 80 --csv --log-file gpurun_out/bw_kernels.csv $CMD > gpurun_out/ncu_bw.log 2>&1; echo "bw kernels rc=$?"
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
