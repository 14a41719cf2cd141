#!/bin/bash
# GPU box: pipeline traces of the training launches (instrumented library variant), then the quick test + bench.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for k in ${TRACE_KINDS:-geom_train rev_sdf rev_deform}; do
  ES_TRACE_EVENTS=${ES_TRACE_EVENTS:-900} timeout 300 python tools/trace_chain.py $k > gpurun_out/trace_$k.txt 2>&1
  tail -n 2 gpurun_out/trace_$k.txt
done
bash tools/gpu_quick.sh
