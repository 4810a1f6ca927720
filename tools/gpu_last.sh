#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 75 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 40 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_nocpu.json 2> gpurun_out/bench_c4_nocpu.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_c4_nocpu.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_c4_nocpu.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'], d['clocks'], d['gpu_launches'])"
