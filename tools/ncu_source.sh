#!/bin/bash
# GPU box: source-level ncu capture (warp-stall samples per line) of ONE fused-chain launch.
#   bash tools/ncu_source.sh fwd|train <skip> <tag>     (train: launches of one training step after 3 warm-up steps)
mode=${1:-fwd}; skip=${2:-3}; tag=${3:-src}
mkdir -p gpurun_out
if [ "$mode" = fwd ]; then CMD="python tools/run_geom_once.py 606208"; else CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference"; fi
ncu --set full --clock-control none --import-source on -k regex:mlp_chain_kernel -s $skip -c 1 -o /tmp/$tag -f $CMD > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/$tag.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${tag}_lines.csv 2>/dev/null
ncu -i /tmp/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
sz=$(stat -c %s /tmp/$tag.ncu-rep); echo rep bytes $sz; [ $sz -lt 45000000 ] && cp /tmp/$tag.ncu-rep gpurun_out/
python tools/ncu_lines.py gpurun_out/${tag}_lines.csv mlp_chain 45 | cut -c1-170
