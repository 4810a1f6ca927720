#!/bin/bash
# N-GPU A/B of the data-parallel backends and their knobs (device-resident timing only).  usage: gpu_dp_ab.sh N
N=${1:-2}
cd "$(dirname "$0")/.."
run() {  # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --quick > gpurun_out/r02_dp${N}_$name.json 2> gpurun_out/r02_dp${N}_$name.err
  echo "$name: $(tail -1 gpurun_out/r02_dp${N}_$name.json | cut -c1-140)"
}
python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | tail -1 | cut -c1-140
run native NLPS_DP_BACKEND=native
run native_low NLPS_DP_BACKEND=native NLPS_COMM_PRIO=low
run native_cta8 NLPS_DP_BACKEND=native NCCL_MAX_CTAS=8
run native_low_cta8 NLPS_DP_BACKEND=native NLPS_COMM_PRIO=low NCCL_MAX_CTAS=8
run native_nograph NLPS_DP_BACKEND=native NLPS_GRAPH=0
run torch NLPS_DP_BACKEND=torch
python bench.py --steps 20 --warmup 5 --quick 2>/dev/null | tail -1 | cut -c1-140
grep -h "nranks" gpurun_out/r02_dp${N}_native.err | head -4
grep -h -i "nvls\|channels\|algo" gpurun_out/r02_dp${N}_native.err | head -6
