#!/bin/bash
# GPU box (last call of the round): the whole -m gpu suite, then - if the budget allows - the shipped-config trainer loop
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 100 python -m pytest tests/ -x -q -m gpu > gpurun_out/suite3.log 2>&1; echo "suite rc=$?"; tail -n 2 gpurun_out/suite3.log
timeout 40 python tools/trainer_loop_bench.py --rays 1024 --samples 32 --steps 20 --warmup 3 --kinds fp16x3 > gpurun_out/trainer_loop_shipped3.json 2> /dev/null; echo "loop rc=$?"; cut -c1-900 gpurun_out/trainer_loop_shipped3.json
