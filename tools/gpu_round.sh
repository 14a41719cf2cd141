#!/bin/bash
# One GPU-box session: the GPU test suite, then short benches.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu -s 2>&1 | tail -120 > gpurun_out/t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
tail -n 12 gpurun_out/t_all.log
cut -c1-400 gpurun_out/bench_train.json
tail -n 3 gpurun_out/bench_train.err
