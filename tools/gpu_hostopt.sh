#!/bin/bash
# GPU box: full -m gpu suite, then the trainer loop at the bench size and at the shipped config
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/ -x -q -m gpu > gpurun_out/suite2.log 2>&1; echo "suite rc=$?"; tail -n 2 gpurun_out/suite2.log
timeout 200 python tools/trainer_loop_bench.py --kinds fp16x3,fp16x1 > gpurun_out/trainer_loop2.json 2> /dev/null; echo "loop rc=$?"; cut -c1-1200 gpurun_out/trainer_loop2.json
timeout 200 python tools/trainer_loop_bench.py --rays 1024 --samples 32 --steps 30 --warmup 5 --kinds fp16x3,fp16x1 > gpurun_out/trainer_loop_shipped2.json 2> /dev/null; echo "loop rc=$?"; cut -c1-1200 gpurun_out/trainer_loop_shipped2.json
