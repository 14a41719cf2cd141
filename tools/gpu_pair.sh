#!/bin/bash
# GPU session: suite on single CTAs, then the CTA-pair kernels (ES_PAIR=1): parity subset + training tests + benches.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -q -m gpu -s -rs 2>&1 | tail -150 > gpurun_out/t_all.log
tail -n 4 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
export ES_PAIR=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t_pair_parity.log
tail -n 4 gpurun_out/t_pair_parity.log
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_helpers.py -q -m gpu 2>&1 | tail -60 > gpurun_out/t_pair_train.log
tail -n 4 gpurun_out/t_pair_train.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_train_pair.json 2> gpurun_out/bench_train_pair.err
timeout 600 python bench.py --mode forward --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_fwd_pair.json 2> gpurun_out/bench_fwd_pair.err
python - <<'PY'
import json
for n in ['train','train_pair','fwd_pair']:
    try:
        j=json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, round(j['value']), round(j['ms_per_step'],2), j['gpu_launches_per_step'], j['clocks']['sm_mhz'], json.dumps(j['roofline']['kernel_ms_per_step']))
    except Exception as e:
        print(n, 'ERR', e); print(open(f'gpurun_out/bench_{n}.err').read()[-1500:])
PY
