"""Summarise the CSV exports of tools/profile_round.sh into the committed evidence files under profiles/.
    python tools/summarise_profiles.py r1v2 r1_v2"""
import csv, io, os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def read_ncu_csv(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    return list(csv.reader(io.StringIO("".join(lines))))

def launch_list(path, out, what):
    rows = read_ncu_csv(path)
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, per_launch = OrderedDict(), []
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", "")); u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        k = r[ik]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
        if "mlp_chain" in k:
            per_launch.append((k, us))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {what}\n# ncu --metrics gpu__time_duration.sum --clock-control none; all launches of the command, aggregated per kernel\n")
        f.write(f"# (per-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes); total {tot/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
        f.write("share_pct,total_ms,launches,kernel\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            f.write(f"{100*us/tot:.2f},{us/1e3:.3f},{n},\"{k[:150]}\"\n")
        f.write("# every launch of the fused chains (us)\n")
        for k, us in per_launch:
            f.write(f"{us:.1f},\"{k[:90]}\"\n")
    return agg, tot

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]

def full_summary(path, f, what):
    rows = read_ncu_csv(path)
    hdr, units = rows[0], rows[1]
    f.write(f"\n## {what}\n")
    seen = {}
    for r in rows[2:]:
        k = r[hdr.index("Kernel Name")]
        seen[k] = seen.get(k, 0) + 1
        if seen[k] > 1:
            continue
        f.write(f"\nkernel: {k}   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                f.write(f"  {key} [{units[i]}] = {r[i]}\n")

launch_list(os.path.join(G, f"{tag}_launches_train.csv"), os.path.join(P, f"{name}_launches_train.csv"),
            "python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline  (default training bench, eager launches)")
launch_list(os.path.join(G, f"{tag}_launches_forward.csv"), os.path.join(P, f"{name}_launches_forward.csv"),
            "python bench.py --mode forward --steps 2 --warmup 3 --no-cpu-baseline")
with open(os.path.join(P, f"{name}_ncu_summary.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none, fused chains at bench size (4096 rays x 128 samples)\n")
    f.write("# per-launch DRAM bytes are the roofline.traffic figures; the chains are tensor/issue bound, not HBM bound\n")
    full_summary(os.path.join(G, f"{tag}_chains_forward_raw.csv"), f, "python bench.py --mode forward --steps 2 --warmup 3 (first launch of each chain)")
    full_summary(os.path.join(G, f"{tag}_chains_train_raw.csv"), f, "python bench.py --steps 1 --warmup 3 --graph 0 (training: stash / reverse chains)")
print(open(os.path.join(P, f"{name}_ncu_summary.txt")).read())
