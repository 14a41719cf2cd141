#!/bin/bash
# GPU box: time the training step with parts of the pipeline switched off (ablation variant of the library, built in the
# build container: ES_NVCC_FLAGS=-DES_ABLATE python -m endosurf_b200.build --tag=ablate).  Results are numerically
# meaningless; only the per-kernel times matter.  flags: 1 no weight copies, 4 no MMAs, 8 no L2 gate prefetch,
# 16 no plane-record stores.
export ES_LIB_PATH=$PWD/endosurf_b200/libendosurf_b200_ablate.so
for f in ${ABLATE_FLAGS:-0 8 16 4 20}; do
  ES_DEBUG_FLAGS=$f timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference > /tmp/ab.json 2>/tmp/ab.err
  python - <<PY
import json
try:
    j=json.load(open('/tmp/ab.json'))
    print("flags $f", round(j['ms_per_step'],2), j['clocks']['sm_mhz'], json.dumps(j['roofline']['kernel_ms_per_step']))
except Exception as e:
    print("flags $f ERR", e, open('/tmp/ab.err').read()[-600:])
PY
done
