#!/bin/bash
# N-GPU A/B of NCCL settings under the native data-parallel back end (device-resident timing)
N=${1:-2}
cd "$(dirname "$0")/.."
run() { name=$1; shift
  out=$(env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 6 --quick 2>/dev/null | tail -1 | cut -c55-90)
  echo "$name: $out"; }
echo "1 GPU: $(python bench.py --quick --steps 20 --warmup 6 2>/dev/null | tail -1 | cut -c55-90)"
run default NLPS_X=0
run min16 NCCL_MIN_CTAS=16
run min32 NCCL_MIN_CTAS=32
run max4 NCCL_MAX_CTAS=4
run max16 NCCL_MAX_CTAS=16
run ll128 NCCL_PROTO=LL128
run simple NCCL_PROTO=Simple
run bucket4M NLPS_DP_BUCKET=4194304
run bucket64M NLPS_DP_BUCKET=67108864
run default2 NLPS_X=0
echo "1 GPU: $(python bench.py --quick --steps 20 --warmup 6 2>/dev/null | tail -1 | cut -c55-90)"
