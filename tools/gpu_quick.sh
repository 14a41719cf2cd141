#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu 2>&1 | tail -5 > gpurun_out/t_train.log; tail -n 3 gpurun_out/t_train.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_train.json'))
print(round(j['value']), round(j['ms_per_step'],2), j['gpu_launches_per_step'], j['clocks'], json.dumps(j['roofline']['kernel_ms_per_step']))
PY
tail -n 3 gpurun_out/bench_train.err
