#!/bin/bash
# GPU box: rebuild with -DES_ABLATE and time the forward path with parts of the pipeline switched off
# (ES_DEBUG_FLAGS: 1 = no weight copies after the first ring fill, 4 = no tcgen05.mma).
ES_NVCC_FLAGS="-DES_ABLATE" python -m endosurf_b200.build --force > /dev/null || exit 1
for f in 0 1 4 5; do echo "== flags $f"; ES_DEBUG_FLAGS=$f python tools/quick_time.py 4096 2>&1 | tail -3; done
python -m endosurf_b200.build --force > /dev/null
