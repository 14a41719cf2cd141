#!/bin/bash
# GPU box: the whole -m gpu suite once with durations, then the 200-step loss-curve parity test twice more (spread)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/ -x -q -m gpu --durations=12 > gpurun_out/suite.log 2>&1; echo "suite rc=$?"; tail -n 22 gpurun_out/suite.log
for i in 1 2; do
  timeout 300 python -m pytest tests/test_gpu_trainer.py -q -m gpu -s -k loss_curve > gpurun_out/curve_$i.log 2>&1; echo "curve $i rc=$?"
  grep -E "loss-curve parity|window means|passed|failed" gpurun_out/curve_$i.log | cut -c1-900
done
