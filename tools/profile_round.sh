#!/bin/bash
# GPU box: the ncu evidence of a round (summarised into profiles/ by tools/summarise_profiles.py afterwards).
#   bash tools/profile_round.sh <tag>        reports stay in /tmp on the box; only CSV exports come back
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
B="python bench.py --no-cpu-baseline --no-gpu-reference"
# launch lists (gpu__time_duration per launch): one training step after 3 warm-up steps, and the forward mode
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/${tag}_launches_train.csv \
    $B --steps 1 --warmup 3 > $out/${tag}_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches_forward.csv \
    $B --mode forward --steps 2 --warmup 3 >> $out/${tag}_ncu_launch.log 2>&1
# full capture of the tensor-core kernels of ONE training step (5 sampling chains, geometry, colour, 3 reverse chains,
# 3 input-adjoint launches, weight gradients = 15 launches; the first 3 steps are warm-up) and of the forward chains
ncu --set full --clock-control none --import-source on -k regex:'mlp_chain_kernel|wgrad_kernel' -s 48 -c 16 \
    -o /tmp/${tag}_train -f $B --steps 1 --warmup 3 > $out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_chain_kernel -s 21 -c 7 -o /tmp/${tag}_forward -f \
    $B --mode forward --steps 1 --warmup 3 >> $out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_train.ncu-rep --page raw --csv > $out/${tag}_train_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_forward.ncu-rep --page raw --csv > $out/${tag}_forward_raw.csv 2>/dev/null
ls -la $out | tail -12
