#!/bin/bash
# GPU box: the evidence files of a round (summarised into profiles/ by tools/summarise_profiles.py afterwards).
#   bash tools/profile_round.sh <tag>        reports stay in /tmp on the box; only CSV exports come back
tag=${1:-rX}
out=gpurun_out
# launch lists (gpu__time_duration per launch) of the default bench command with eager launches, and of the forward mode
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $out/${tag}_launches_train.csv \
    python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches_forward.csv \
    python bench.py --mode forward --steps 2 --warmup 3 --no-cpu-baseline >> $out/${tag}_ncu_launch.log 2>&1
# full capture of the fused chains at bench size: forward kernels, then the training (stash / reverse) kernels
ncu --set full --clock-control none -k regex:mlp_chain -s 24 -c 9 -o /tmp/${tag}_chains_forward -f \
    python bench.py --mode forward --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:mlp_chain -s 40 -c 12 -o /tmp/${tag}_chains_train -f \
    python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline >> $out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_chains_forward.ncu-rep --page raw --csv > $out/${tag}_chains_forward_raw.csv 2>/dev/null
ncu -i /tmp/${tag}_chains_train.ncu-rep --page raw --csv > $out/${tag}_chains_train_raw.csv 2>/dev/null
ls -la $out | tail -12
