#!/usr/bin/env python
"""Benchmark of the render_rays hot path (BASELINE.json: rays/sec, 512x512 frames, 64+64 samples).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this repo's CUDA path
    python bench.py --impl reference --steps 3 --warmup 1          # reference algorithm on the host CPU cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                     # one rank per GPU, weak scaling over ray batches

A step is one render_rays call on a batch of `--rays` synthetic rays of one 512x512 frame (configs[1] of
BASELINE.json: 4096 rays, 64 coarse + 64 fine samples, 4 up-sampling steps, 9x256 networks, fp32 parity mode).
Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for what every key means.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- algorithmic work per unit (1 MAC = 2 FLOP) -----------------------------------------------------
D_MAC, S_MAC, C_MAC = 459_520, 544_512, 638_208  # deform / sdf / colour MLP, MAC per point


def flops_per_ray(ns, ni, steps):
    m = ns + ni
    u = ns + (steps - 1) * ni // steps if ni > 0 else 0
    return 2.0 * (u * (D_MAC + S_MAC) + m * (4 * D_MAC + 2 * S_MAC + C_MAC))


NET_CFG = {
    "bound": 1.0, "use_deform": True,
    "deform_network": {"enc_pos_cfg": {"enc_type": "frequency", "input_dim": 3, "multires": 6},
                       "enc_time_cfg": {"enc_type": "frequency", "input_dim": 1, "multires": 6},
                       "n_layers": 9, "hidden_dim": 256, "skips": [4], "out_dim": 3},
    "sdf_network": {"enc_pos_cfg": {"enc_type": "frequency", "input_dim": 3, "multires": 6},
                    "n_layers": 9, "hidden_dim": 256, "skips": [4], "out_dim": 257, "geometric_init": True,
                    "geometric_init_bias": 0.8},
    "color_network": {"enc_pos_cfg": {"enc_type": "frequency", "input_dim": 3, "multires": 10},
                      "enc_dir_cfg": {"enc_type": "frequency", "input_dim": 3, "multires": 4},
                      "n_layers": 9, "hidden_dim": 256, "skips": [4], "feat_dim": 256, "out_dim": 3},
    "deviation_network": {"init_val": 0.3},
}  # == net: of configs/endosurf/baseline/base_pull.yml:40-82
RENDER_CFG = {"type": "endosurf", "net_chunk": 80000, "anneal_end": 50000, "n_samples": 64, "n_importance": 64,
              "important_begin_iter": 0, "up_sample_steps": 4, "perturb": True}
ITER_STEP = 50000


def make_rays(n_rays, frame, hw=512, n_frames=60, seed=0):
    """Pinhole rays from o=(0,0,-1.5) over a 512x512 image (focal 1.2*W), one frame per batch, time=frame/59."""
    g = torch.Generator().manual_seed(seed + 7919 * frame)
    pix = torch.randint(0, hw * hw, (n_rays,), generator=g)
    u = (pix % hw).float() + 0.5
    v = (pix // hw).float() + 0.5
    f = 1.2 * hw
    d = torch.stack([(u - hw / 2) / f, (v - hw / 2) / f, torch.ones_like(u)], -1)
    d = d / d.norm(dim=-1, keepdim=True)
    o = torch.tensor([0.0, 0.0, -1.5]).expand(n_rays, 3)
    return torch.cat([o, d, torch.zeros(n_rays, 2), torch.full((n_rays, 1), frame / (n_frames - 1))], -1).contiguous()


def make_targets(n_rays, frame, seed=0):
    g = torch.Generator().manual_seed(seed + 31 * frame + 5)
    return torch.rand(n_rays, 3, generator=g), 0.5 + 0.5 * torch.rand(n_rays, 1, generator=g)


def train_loss(o, color_gt, depth_gt):
    """L1 colour + L1 depth + 0.1 eikonal: the render_rays part of the reference loss (trainer_endosurf.py:132-162)."""
    return (o["color_map"] - color_gt).abs().mean() + (o["depth_map"] - depth_gt).abs().mean() + \
        0.1 * o["gradient_o_error"]


def seeded_state(renderer_module):
    """Random-init weights of the reference architecture: geometric SDF init + seeded noise (a deforming surface)."""
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        for n, p in renderer_module.named_parameters():
            s = 0.004 if "sdf_network" in n else 0.02
            p.add_(s * torch.randn(p.shape, generator=g).to(p.device))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.stop_flag, self.idx = [], False, gpu_index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([c.strip() for c in line.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.t.start()

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j.get("bf16_tflops_sustained", j.get("bf16_tflops"))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained 1.4 PFLOP/s)"


def oracle_cpu_rays_per_s(n_rays, repeats, threads, mode="train"):
    """Reference algorithm (oracle port; the reference is pure Python/PyTorch and cannot travel) on the host CPU."""
    from oracle import endosurf_oracle as orc  # checker / baseline only
    from endosurf_b200 import EndoSurfNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = EndoSurfNet(NET_CFG)
    seeded_state(model)
    train = mode == "train"
    ck = {k: {kk: vv.detach().clone().requires_grad_(train) for kk, vv in sd.items()}
          for k, sd in model.save_checkpoint().items()}
    net = orc.OracleNet(ck, NET_CFG)
    rc = copy.deepcopy(RENDER_CFG)
    params = [p for sd in ck.values() for p in sd.values()]
    opt = torch.optim.Adam(params, lr=5e-4) if train else None
    times = []
    for i in range(repeats + 1):
        rays = make_rays(n_rays, frame=i)
        cgt, dgt = make_targets(n_rays, frame=i)
        t0 = time.perf_counter()
        if train:
            opt.zero_grad()
            o = orc.render_rays(net, rc, rays, iter_step=ITER_STEP)
            loss = train_loss(o, cgt, dgt)
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                orc.render_rays(net, rc, rays, iter_step=ITER_STEP)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times[1:])) if repeats > 0 else times[0]
    return n_rays / t, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_rays
    from oracle import endosurf_oracle as orc  # noqa
    v, _ = oracle_cpu_rays_per_s(n, 0, threads, args.mode)  # warm the allocator / thread pool
    times = []
    for _ in range(max(args.warmup - 1, 0)):
        oracle_cpu_rays_per_s(n, 0, threads, args.mode)
    for _ in range(args.steps):
        _, t = oracle_cpu_rays_per_s(n, 0, threads, args.mode)
        times.append(t)
    ms = 1e3 * float(np.mean(times))
    val = n / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRICS[args.mode], "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.rays, mode=args.mode),
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{n} rays x (64+64 samples, 4 up-sampling steps) per step, {args.mode}, torch "
                                   f"{torch.__version__} CPU fp32, {threads} threads"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel at this workload's size, from the
# committed ncu --set full captures (profiles/r1_v2_ncu_summary.txt; training: the stash-writing variant, 524288 points,
# tools/ncu_one_kernel.sh -> profiles/r1_v2_ncu_stash_kernel.txt)
NCU_TRAFFIC_BYTES = {"forward": 11.487488e6 + 525.844480e6, "train": 0.696432896e9 + 26.990113e9}

METRICS = {"train": "training rays/sec (512x512 frames, 64+64 samples)",
           "forward": "render_rays forward rays/sec (512x512 frames, 64+64 samples)"}
METRIC = METRICS["train"]


def train_flops_per_ray(ns, ni, steps):
    """SURVEY 8d convention: up-sampling x1 + render_core x3 (forward + ~2x backward)."""
    m = ns + ni
    u = ns + (steps - 1) * ni // steps if ni > 0 else 0
    return 2.0 * (u * (D_MAC + S_MAC) + 3 * m * (4 * D_MAC + 2 * S_MAC + C_MAC))


def workload_config(n_rays, note=None, mode="train"):
    what = ("training step (render_rays forward + L1 colour/depth + eikonal loss, backward, Adam)" if mode == "train"
            else "render_rays forward")
    c = {"workload": f"{what}, {n_rays}-ray batch of one 512x512 frame, 64 coarse + 64 fine samples, "
                     "4 up-sampling steps, deform+sdf+colour 9x256 MLPs, fp32-parity (fp16 hi/lo x3) tensor-core mode",
         "rays_per_step_per_gpu": n_rays, "n_samples": 64, "n_importance": 64, "up_sample_steps": 4,
         "parallelism": ("rays sharded per rank; one NCCL all-reduce of the flat 1.65 M-float gradient bucket per step"
                         if mode == "train" else "rays sharded per rank, no data-path collective (forward)"),
         "l2_policy": "per-step working set (>= 1 GiB of per-point scratch) exceeds the 126 MB L2; ray batches rotate"}
    if mode == "train":
        c["precision"] = ("forward and activation-gradient chains: fp16 hi/lo x3 products (fp32 parity); weight/bias "
                          "gradients: own split-K tcgen05 kernel on the fp16 hi planes, fp32 accumulation; gradient "
                          "parity of this exact configuration against the oracle's autograd at 512 rays x 128 samples: "
                          "tests/test_gpu_training.py::test_timed_path_gradient_parity_512_rays")
    if note:
        c["note"] = note
    return c


def allreduce_gradients(params, world):
    """Data parallelism over ray batches: ONE all-reduce of the flat gradient bucket per step."""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from endosurf_b200 import EndoSurfRenderer

    train = args.mode == "train"
    torch.manual_seed(0)
    r = EndoSurfRenderer(copy.deepcopy(RENDER_CFG), NET_CFG, device=f"cuda:{local}")
    seeded_state(r.model)
    r.train(train)
    params = [p for v in r.get_train_params().values() for p in v]
    opt = torch.optim.Adam(params, lr=5e-4) if train else None
    R, K, W = args.rays, args.steps, args.warmup
    n_batches = min(K + W, 16)
    frames = [(rank * 17 + i) % 60 for i in range(n_batches)]
    host = [make_rays(R, frame=f, seed=rank).pin_memory() for f in frames]
    host_t = [tuple(x.pin_memory() for x in make_targets(R, frame=f, seed=rank)) for f in frames]
    devb = [h.to(dev) for h in host]
    devt = [(c.to(dev), d.to(dev)) for c, d in host_t]
    out_c = torch.empty(R, 3).pin_memory()
    out_d = torch.empty(R, 1).pin_memory()
    out_l = torch.empty(()).pin_memory()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph_err = ""

    def step(rays, cgt, dgt):
        if train:
            opt.zero_grad(set_to_none=True)
            o = r(rays, iter_step=ITER_STEP)
            loss = train_loss(o, cgt, dgt)
            loss.backward()
            if world > 1:
                allreduce_gradients(params, world)
            opt.step()
            return o, loss
        with torch.no_grad():
            return r.render_rays(rays, iter_step=ITER_STEP), None

    for i in range(W):
        step(devb[i % n_batches], *devt[i % n_batches])
    r.sync_check()
    # ---------------- per-kernel roofline: CUDA events around every fused-chain launch, eager steps of the same workload
    r.profile(True)
    r.profile_read()
    launches0 = r.launch_count()
    n_prof = min(K, 4)
    for i in range(n_prof):
        step(devb[(W + i) % n_batches], *devt[(W + i) % n_batches])
    prof = r.profile_read()
    launches_per_step = (r.launch_count() - launches0) / n_prof
    r.profile(False)
    graphed = None
    if graphed is None:  # eager steps: the kernel events are taken over the timed region itself
        r.profile(True)
        r.profile_read()
    # ---------------- device-resident timing (value)
    clocks = ClockSampler(local)
    sync_all()
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(devb[(W + i) % n_batches], *devt[(W + i) % n_batches])
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = launches_per_step * K
    if graphed is None:
        prof, n_prof = r.profile_read(), K
        r.profile(False)
    # ---------------- end to end: pinned host rays (+ targets) in, colour + depth (+ loss) back to the host, every step
    sync_all()
    t0 = time.perf_counter()
    for i in range(K):
        j = (W + i) % n_batches
        rays = host[j].to(dev, non_blocking=True)
        cgt, dgt = (x.to(dev, non_blocking=True) for x in host_t[j])
        o, loss = step(rays, cgt, dgt)
        out_c.copy_(o["color_map"].detach(), non_blocking=True)
        out_d.copy_(o["depth_map"].detach(), non_blocking=True)
        if loss is not None:
            out_l.copy_(loss.detach(), non_blocking=True)  # the reference reads loss.item() every step
        torch.cuda.current_stream().synchronize()
    sync_all()
    e2e_s = time.perf_counter() - t0
    r.sync_check()

    t = torch.tensor([ms_total, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t[0].item() / K
    e2e_ms_step = t[1].item() / K
    if rank == 0:
        value = world * R / (ms_step * 1e-3)
        e2e = world * R / (e2e_ms_step * 1e-3)
        peak, peak_src = measured_peak_tflops()
        g = prof["geometry_chain"]
        pts_per_launch = g["points"] / max(g["launches"], 1)
        alg_flops_launch = pts_per_launch * 2.0 * (4 * D_MAC + 2 * S_MAC)
        ms_launch = g["ms"] / max(g["launches"], 1)
        achieved = alg_flops_launch / (ms_launch * 1e-3) / 1e12 if ms_launch > 0 else 0.0
        kern_ms = {k: round(v["ms"] / n_prof, 4) for k, v in prof.items()}
        fpr = train_flops_per_ray(64, 64, 4) if train else flops_per_ray(64, 64, 4)
        h2d = R * 9 * 4 + (R * 4 * 4 if train else 0)
        d2h = R * 4 * 4 + (4 if train else 0)
        line = {
            "metric": METRICS[args.mode], "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16 hi/lo x3 tensor-core products, fp32 accumulate)", "data": "synthetic",
            "config": workload_config(R, mode=args.mode, note=(
                "forward+loss+backward replayed as one CUDA graph per step (GraphedTrainStep); Adam and the gradient "
                "all-reduce run eagerly after it" if graphed is not None else
                ("eager launches" + (f" (graph capture failed: {graph_err})" if train and args.graph else "")))),
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_step},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "mlp_chain_kernel<CHAIN_SDF,TANGENT> (forward geometry chain: "
                                                      "deform+sdf MLPs with 3 forward-mode tangent rows per point)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": NCU_TRAFFIC_BYTES.get(args.mode),
                         "algorithmic_flops_per_launch": alg_flops_launch, "points_per_launch": pts_per_launch,
                         "ms_per_launch": ms_launch,
                         "note": "algorithmic = 2*(4D+2S) FLOP/point (SURVEY 8d); the kernel issues 3 fp16 MMAs per "
                                 "product (hi/lo split) and 4D+4S+feat (forward-mode normals), i.e. ~3.6x the "
                                 "algorithmic MMA work, so frac <= ~0.28 by construction",
                         "kernel_ms_per_step": kern_ms,
                         "kernel_timing": ("CUDA events around every chain launch over the timed region" if graphed is None
                                           else f"CUDA events around every chain launch in {n_prof} eager steps of the same "
                                                "workload run inside this process before the graph-replayed timed region"),
                         "step_algorithmic_tflops": value / world * fpr / 1e12},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, tt = oracle_cpu_rays_per_s(args.ref_rays, 3, threads, args.mode)  # 1 warm-up + 3 timed passes, ~10 s
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.ref_rays} rays of the same workload ({args.mode}), oracle port "
                                              f"of the reference (PyTorch {torch.__version__} CPU fp32), median of 3 "
                                              f"passes of {tt:.1f} s after one warm-up pass"}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = sys.stdout


def main():
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner on the first collective)
    # are sent to stderr for the duration of the run
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=4096, help="rays per step per GPU (BASELINE configs[1])")
    ap.add_argument("--ref-rays", type=int, default=512,
                    help="bounded CPU sample per step for the reference arm / cpu_baseline (about 2.5 s of a 16-core host)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="train mode: replay forward+backward as one CUDA graph")
    ap.add_argument("--mode", default="train", choices=["train", "forward"],
                    help="train: BASELINE.json's metric (training rays/s); forward: inference render_rays")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
