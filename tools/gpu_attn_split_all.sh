#!/bin/bash
# per-kernel device time of the attention kernels at one shape (ncu serialised times)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SHAPE=${SHAPE:-"64 512 8 64 2"}
NLPS_BENCH_FAST=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ -s 30 -c 80 --csv --log-file gpurun_out/attn_split.csv python tools/attn_bench.py $SHAPE > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/attn_split.csv 1 | sed 's/ms\/step/ms total/'
