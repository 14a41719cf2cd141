#!/bin/bash
# GPU box: rebuild the library with different A-ring / weight-ring depths and time the forward path.
# usage: bash tools/sweep_rings.sh "3,6 2,8 4,4"
for v in $1; do
  ns=${v%,*}; nt=${v#*,}
  ES_NVCC_FLAGS="-DES_NSLOT=$ns -DES_NSTAGE=$nt" python -m endosurf_b200.build --force > /dev/null || { echo "build failed $v"; continue; }
  echo "== NSLOT=$ns NSTAGE=$nt"
  python tools/quick_time.py 4096 2>&1 | tail -3
done
python -m endosurf_b200.build --force > /dev/null
