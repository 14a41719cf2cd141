#!/bin/bash
# GPU box: whole GPU test suite + training and forward benches (no reference arms)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/t_all.log; tail -n 4 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 600 python bench.py --mode forward --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_fwd.json 2> gpurun_out/bench_fwd.err
python - <<'PY'
import json
for n in ['train','fwd']:
    try:
        j=json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, round(j['value']), round(j['ms_per_step'],2), j['gpu_launches_per_step'], j['clocks']['sm_mhz'], round(j['roofline']['frac'],4), json.dumps(j['roofline']['kernel_ms_per_step']))
    except Exception as e:
        print(n, 'ERR', e); print(open(f'gpurun_out/bench_{n}.err').read()[-800:])
PY
