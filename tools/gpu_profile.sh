#!/bin/bash
# ncu launch list of one bench command (per-launch device time; shares, not absolutes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WL=${1:-c4}
CMD="python bench.py --workload $WL --steps 2 --warmup 1 --quick"
$CMD > gpurun_out/plain_$WL.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$WL.csv $CMD > gpurun_out/ncu_$WL.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/plain_$WL.log | cut -c1-300; wc -l gpurun_out/launches_$WL.csv
