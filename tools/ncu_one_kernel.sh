#!/bin/bash
# GPU box: ncu --set full of ONE launch of the training-mode geometry chain (stash variant) at bench size; prints its DRAM bytes.
ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:mlp_chain_kernel<\(int\)0, \(bool\)1, \(bool\)0, \(bool\)1>' -s 1 -c 1 \
    -o /tmp/one_kernel -f python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/one_kernel.log 2>&1
ncu -i /tmp/one_kernel.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/one_kernel_raw.csv
python - <<'PY'
import csv, io
rows = list(csv.reader(io.StringIO("".join(l for l in open("gpurun_out/one_kernel_raw.csv") if l.startswith('"')))))
h, u, r = rows[0], rows[1], rows[2]
for k in ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]:
    i = h.index(k); print(k, u[i], r[i])
PY
