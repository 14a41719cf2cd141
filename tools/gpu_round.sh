#!/bin/bash
# One GPU-box session: the GPU test suite, then the benches of every mode.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
ls -la oracle/_ref > gpurun_out/ref_ls.txt 2>&1
timeout 1800 python -m pytest tests -q -m gpu -s -rs 2>&1 | tail -150 > gpurun_out/t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 600 python bench.py --mode forward --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fwd.json 2> gpurun_out/bench_fwd.err
timeout 600 python bench.py --mode frame --steps 3 --warmup 3 > gpurun_out/bench_frame.json 2> gpurun_out/bench_frame.err
timeout 600 python bench.py --mode grid256 --steps 3 --warmup 3 > gpurun_out/bench_grid.json 2> gpurun_out/bench_grid.err
timeout 600 python bench.py --precision-terms 1 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_train_p1.json 2> gpurun_out/bench_train_p1.err
timeout 600 python bench.py --rays 8192 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_train_8192.json 2> gpurun_out/bench_train_8192.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -n 8 gpurun_out/t_all.log
for f in train fwd frame grid train_p1 train_8192 ref; do echo "== $f"; cut -c1-300 gpurun_out/bench_$f.json; tail -n 2 gpurun_out/bench_$f.err; done
