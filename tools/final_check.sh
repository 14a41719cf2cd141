mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -n 4 gpurun_out/final_smoke.log
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -n 3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_bench.json
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; cut -c1-300 gpurun_out/final_bench_ref.json
