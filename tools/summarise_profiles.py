"""Summarise the CSV exports of tools/profile_round.sh into the committed evidence files under profiles/.
    python tools/summarise_profiles.py r2 r2
writes profiles/<name>_launches_{train,forward}.csv, profiles/<name>_ncu_summary.txt and
profiles/<name>_ncu_traffic.json (the per-launch DRAM bytes bench.py reports as roofline.traffic)."""
import csv, io, json, os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
OURS = ("mlp_chain", "wgrad", "composite", "pack_jobs", "wn_", "smallm", "amax", "scale_from", "upsample", "merge_z",
        "coarse_z", "points_kernel", "eikonal", "sum_reduce", "point_adjoints", "grid_points", "fill_identity")


def read_ncu_csv(path):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    return list(csv.reader(io.StringIO("".join(lines))))


def launch_list(path, out, what, last_step_from=None):
    rows = read_ncu_csv(path)
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = []
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", "")); u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        launches.append((r[ik], us))
    if last_step_from is not None:
        # keep the launches of the LAST step: from the last occurrence of the step's first kernel
        idx = max(i for i, (k, _) in enumerate(launches) if last_step_from in k)
        launches = launches[idx:]
    agg = OrderedDict()
    for k, us in launches:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if any(o in k for o in OURS))
    with open(out, "w") as f:
        f.write(f"# {what}\n# ncu --metrics gpu__time_duration.sum --clock-control none; aggregated per kernel\n")
        f.write(f"# (per-launch times are cold-cache and serialised by the profiler: compare SHARES, not absolutes)\n")
        f.write(f"# total {tot/1e3:.2f} ms over {len(launches)} launches; this repo's kernels: {100*ours/tot:.1f} % of the time\n")
        f.write("share_pct,total_ms,launches,kernel\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{100*us/tot:.2f},{us/1e3:.3f},{n},\"{k[:160]}\"\n")
        f.write("# every launch in order (us)\n")
        for k, us in launches:
            f.write(f"{us:.1f},\"{k[:110]}\"\n")
    return agg, tot


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def full_summary(path, f, what):
    rows = read_ncu_csv(path)
    hdr, units = rows[0], rows[1]
    f.write(f"\n## {what}\n")
    out = []
    for n, r in enumerate(rows[2:]):
        k = r[hdr.index("Kernel Name")]
        f.write(f"\n[{n}] kernel: {k[:200]}   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
        rec = {"kernel": k}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                f.write(f"  {key} [{units[i]}] = {r[i]}\n")
                rec[key] = (r[i], units[i])
        out.append(rec)
    return out


launch_list(os.path.join(G, f"{tag}_launches_train.csv"), os.path.join(P, f"{name}_launches_train.csv"),
            "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference : the LAST (timed) training step",
            last_step_from="pack_jobs")
launch_list(os.path.join(G, f"{tag}_launches_forward.csv"), os.path.join(P, f"{name}_launches_forward.csv"),
            "python bench.py --mode forward --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference : all launches")
traffic = {}
with open(os.path.join(P, f"{name}_ncu_summary.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on, tensor-core kernels at bench size (4096 rays x 128 samples)\n")
    tr = full_summary(os.path.join(G, f"{tag}_train_raw.csv"), f,
                      "one training step (python bench.py --steps 1 --warmup 3): sampling chains, geometry + colour "
                      "(training variants), reverse chains, input-adjoint launches, weight-gradient kernel")
    fw = full_summary(os.path.join(G, f"{tag}_forward_raw.csv"), f, "one forward step (python bench.py --mode forward)")

    def dram(rec):
        return to_bytes(*rec["dram__bytes_read.sum"]) + to_bytes(*rec["dram__bytes_write.sum"])

    def is_geom(rec, stash):
        k = rec["kernel"]
        return "mlp_chain_kernel<0, 1, 0, " + ("1" if stash else "0") in k
    for mode, recs, stash in (("train", tr, True), ("forward", fw, False)):
        g = [r for r in recs if is_geom(r, stash)]
        if g:
            traffic[mode] = {"kernel": g[0]["kernel"][:120], "dram_bytes_per_launch": dram(g[0]),
                             "duration": g[0]["gpu__time_duration.sum"][0] + " " + g[0]["gpu__time_duration.sum"][1],
                             "tensor_pipe_pct": g[0]["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][0],
                             "source": f"profiles/{name}_ncu_summary.txt"}
with open(os.path.join(P, f"{name}_ncu_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print(open(os.path.join(P, f"{name}_ncu_summary.txt")).read()[:6000])
print(json.dumps(traffic, indent=1))
