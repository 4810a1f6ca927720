#!/bin/bash
# N-GPU check of the driver's scaling command (native NCCL back end): full bench line + torch back end quick line
N=${1:-8}
cd "$(dirname "$0")/.."
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c4_${N}gpu.json 2> gpurun_out/r02_bench_c4_${N}gpu.err
echo "rc=$? $(cut -c1-400 gpurun_out/r02_bench_c4_${N}gpu.json)"
python - <<PY
import json
l=json.loads(open("gpurun_out/r02_bench_c4_${N}gpu.json").read().strip().splitlines()[-1])
print("ms", l["ms_per_step"], "value", l["value"], "e2e", l["e2e"]["value"], "replicas_identical", l.get("replicas_identical"), "other", {k:(round(v["ms_per_step"],2), round(v["value"])) for k,v in (l.get("other") or {}).items()})
PY
grep -c "NCCL INFO" gpurun_out/r02_bench_c4_${N}gpu.err; grep -m2 "nranks" gpurun_out/r02_bench_c4_${N}gpu.err | cut -c1-200; grep -m3 -i "NVLS" gpurun_out/r02_bench_c4_${N}gpu.err | cut -c1-200
NLPS_DP_BACKEND=torch python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus $N --steps 20 --warmup 5 --quick 2>/dev/null | tail -1 | cut -c1-150
