#!/bin/bash
# weak-scaling series on ONE 8-GPU box, device-resident timing only (the driver's own series runs the full line)
cd "$(dirname "$0")/.."
for N in 1 2 4 8 1; do
  if [ $N = 1 ]; then
    out=$(python bench.py --quick --steps 20 --warmup 6 2>/dev/null | tail -1)
  else
    out=$(python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 6 --quick 2>/dev/null | tail -1)
  fi
  echo "N=$N $out" | cut -c1-130
done | tee gpurun_out/r02_scale_series.txt
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv,noheader | head -8
