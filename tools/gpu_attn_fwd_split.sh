#!/bin/bash
cd "$(dirname "$0")/.."
SHAPE=${SHAPE:-"64 512 8 64 2"}
NLPS_BENCH_FAST=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_fwd -s 5 -c 12 --csv --log-file gpurun_out/attn_fwd_split.csv python tools/attn_bench.py $SHAPE > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/attn_fwd_split.csv 1 | grep attn
