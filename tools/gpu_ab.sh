#!/bin/bash
# GPU box: A/B of an environment switch on the SAME box, interleaved: bash tools/gpu_ab.sh VAR "0 1" [mode]
var=$1; vals=$2; mode=${3:-train}
mkdir -p gpurun_out
for rep in 1 2; do for v in $vals; do
  env $var=$v timeout 600 python bench.py --mode $mode --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > /tmp/ab.json 2>/tmp/ab.err
  python - <<PY
import json
try:
    j=json.load(open('/tmp/ab.json'))
    print("$var=$v $mode", round(j['value']), round(j['ms_per_step'],2), j['clocks']['sm_mhz'], json.dumps(j['roofline']['kernel_ms_per_step']))
except Exception as e:
    print("$var=$v ERR", e, open('/tmp/ab.err').read()[-500:])
PY
done; done
