#!/bin/bash
# GPU box: config-3 trainer loop, B2 reference latencies (frame / 256^3 grid), smoke of the rebuilt library
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 400 python tools/trainer_loop_bench.py > gpurun_out/trainer_loop.json 2> gpurun_out/trainer_loop.err; echo "loop rc=$?"; cut -c1-1500 gpurun_out/trainer_loop.json; tail -n 3 gpurun_out/trainer_loop.err
timeout 300 python bench.py --mode frame --steps 3 --warmup 3 > gpurun_out/bench_frame.json 2> gpurun_out/bench_frame.err; echo "frame rc=$?"
timeout 200 python bench.py --mode grid256 --steps 3 --warmup 3 > gpurun_out/bench_grid256.json 2> gpurun_out/bench_grid256.err; echo "grid rc=$?"
python - <<'PY'
import json
for f in ("frame", "grid256"):
    try:
        j = json.load(open(f"gpurun_out/bench_{f}.json"))
        print(f, round(j["value"], 2), j["unit"], "e2e", round(j["e2e"]["value"], 2), json.dumps(j.get("gpu_reference"))[:400])
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/bench_{f}.err").read()[-600:])
PY
