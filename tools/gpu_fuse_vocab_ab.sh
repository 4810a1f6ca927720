#!/bin/bash
# same-box A/B of the fused vocabulary path (NLPS_FUSE_VOCAB) on the configurations with a vocabulary head that matters
cd "$(dirname "$0")/.."
for wl in c3 c4 c4-32k c5; do
  for mb in 0 32 64 96; do
    if [ $mb = 0 ]; then envs="NLPS_FUSE_VOCAB=0"; else envs="NLPS_FUSE_VOCAB=1 NLPS_VOCAB_CHUNK_MB=$mb"; fi
    echo "$wl $envs: $(env $envs python bench.py --workload $wl --quick --steps 12 --warmup 6 2>/dev/null | tail -1 | cut -c1-100)"
  done
done
