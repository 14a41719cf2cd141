#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | grep -v "^  \|Warning" | tail -80 > gpurun_out/t_all.log; grep -n "passed\|failed\|first-step\|host sync\|Error\|^FAILED" gpurun_out/t_all.log | cut -c1-1500
for P in 0 1; do
export ES_PAIR=$P
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_train_p$P.json 2> gpurun_out/bench_train_p$P.err
timeout 600 python bench.py --mode forward --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_fwd_p$P.json 2> gpurun_out/bench_fwd_p$P.err
done
python - <<'PY'
import json
for n in ['train_p0','fwd_p0','train_p1','fwd_p1']:
    try:
        j=json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, round(j['value']), round(j['ms_per_step'],2), j['gpu_launches_per_step'], j['clocks']['sm_mhz'], json.dumps(j['roofline']['kernel_ms_per_step']))
    except Exception as e:
        print(n, 'ERR', e); print(open(f'gpurun_out/bench_{n}.err').read()[-800:])
PY
