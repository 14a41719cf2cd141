#!/bin/bash
# GPU box with N GPUs: bash tools/gpu_scale.sh N    -> gpurun_out/bench_train_n{N}[_8192].json
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_n$N.txt
run() {  # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference $2 > gpurun_out/$1.json 2> gpurun_out/$1.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/$1.json'))
    print('$1', round(j['value']), round(j['ms_per_step'],2), j['n_gpus'], j['config']['rays_per_step_per_gpu'], j['clocks'])
except Exception as e:
    print('$1 ERR', e); print(open('gpurun_out/$1.err').read()[-1500:])
PY
}
run bench_train_n$N ""
run bench_train_n${N}_8192 "--rays 8192"
