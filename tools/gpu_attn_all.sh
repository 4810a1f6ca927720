#!/bin/bash
# attention regression visit: kernel parity (all shapes), per-kernel times at c4 and c5 shapes, every bench workload once
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" --timeout=200 2>&1 | tail -3
SHAPE="64 512 8 64 2" bash tools/gpu_attn_split_all.sh
SHAPE="8 2048 12 64 2" bash tools/gpu_attn_split_all.sh
for w in c1 c2 c3 c5 c4; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$w.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'][:34], '| ms/step', round(d['ms_per_step'],3), '| value', round(d['value']))" || tail -2 gpurun_out/bench_$w.err; done
