#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "" "NLPS_GRAPH=0" "NLPS_FUSE_DELTA=0" ""; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['e2e'], d['clocks'])"
done
