#!/usr/bin/env python
"""Benchmark of the render_rays hot path (BASELINE.json: training rays/sec, 512x512 frames, 64+64 samples).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this repo's CUDA path, training step (default)
    python bench.py --mode forward                                 # inference render_rays
    python bench.py --mode frame                                   # BASELINE configs[4]: one full 512x512 frame, latency
    python bench.py --mode grid256                                 # BASELINE configs[4]: 256^3 SDF grid query, latency
    python bench.py --precision-terms 1                            # BASELINE configs[2]: single-pass fp16 training
    python bench.py --impl reference --steps 3 --warmup 1          # the UNMODIFIED reference on the host CPU cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench.py --gpus N --steps K --warmup W                     # one rank per GPU, ray shards (weak scaling)

A training step is: render_rays forward on a batch of `--rays` synthetic rays of one 512x512 frame (configs[1] of
BASELINE.json: 4096 rays, 64 coarse + 64 fine samples, 4 up-sampling steps, 9x256 networks, fp32-parity mode), the
reference trainer's masked-mean colour / depth losses + eikonal term, backward, (N > 1: the data-parallel gradient
all-reduce of endosurf_b200.distributed) and Adam.  Rank 0 prints ONE JSON line.  DESIGN.md section 7 explains every key.
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


def train_flops_per_ray(ns, ni, steps):
    """SURVEY 8d convention: up-sampling x1 + render_core x3 (forward + ~2x backward)."""
    m = ns + ni
    u = ns + (steps - 1) * ni // steps if ni > 0 else 0
    return 2.0 * (u * (D_MAC + S_MAC) + 3 * m * (4 * D_MAC + 2 * S_MAC + C_MAC))


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
HW = 512

METRICS = {"train": "training rays/sec (512x512 frames, 64+64 samples)",
           "forward": "render_rays forward rays/sec (512x512 frames, 64+64 samples)",
           "frame": "inference latency of one full 512x512 frame (64+64 samples)",
           "grid256": "latency of the 256^3 marching-cubes SDF grid query"}


def make_rays(n_rays, frame, hw=HW, n_frames=60, seed=0, all_pixels=False):
    """Pinhole rays from o=(0,0,-1.5) over a 512x512 image (focal 1.2*W), one frame per batch, time=frame/59."""
    if all_pixels:
        pix = torch.arange(hw * hw)
        n_rays = hw * hw
    else:
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
    """colour [R,3], depth [R,1], mask [R,1] (about 85 % of the pixels carry a valid colour/depth, like an endoscopic
    frame with tool masks)."""
    g = torch.Generator().manual_seed(seed + 31 * frame + 5)
    return (torch.rand(n_rays, 3, generator=g), 0.5 + 0.5 * torch.rand(n_rays, 1, generator=g),
            (torch.rand(n_rays, 1, generator=g) < 0.85).float())


def train_loss(o, color_gt, depth_gt, mask):
    """The render_rays part of the reference loss with its masked-mean normalisation (trainer_endosurf.py:131-152,
    weights of configs/endosurf/baseline/base_pull.yml): L1 colour + L1 depth + 0.1 eikonal."""
    ce = (o["color_map"] - color_gt) * mask
    de = (o["depth_map"] - depth_gt) * mask
    den = mask.sum() + 1e-10
    return ce.abs().sum() / den + de.abs().sum() / den + 0.1 * o["gradient_o_error"]


def seeded_state(module):
    """Random-init weights of the reference architecture: geometric SDF init + seeded noise (a deforming surface)."""
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        for n, p in module.named_parameters():
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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return (float(j.get("bf16_tflops_sustained", j.get("bf16_tflops"))),
                "measured (MEASURED_PEAKS.json bf16_tflops_sustained)", float(j.get("hbm_gbs", 6500.0)))
    return 1400.0, "fallback (B200_PROFILING.md sustained 1.4 PFLOP/s)", 6500.0


def ncu_capture(mode, key="dram_bytes_per_launch"):
    """One figure of the committed `ncu --set full` capture of the roofline kernel at this workload's size
    (profiles/r2_ncu_traffic.json, extracted from profiles/r2_ncu_summary.txt): `dram_bytes_per_launch` =
    dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, `tensor_pipe_pct` =
    sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed; null when no capture has been committed."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(p) as f:
            v = json.load(f).get(mode, {}).get(key)
        return None if v is None else float(v)
    except Exception:
        return None


def ncu_traffic(mode):
    return ncu_capture(mode, "dram_bytes_per_launch")


# ------------------------------------------------------------------------------------------------ the reference itself
def _reference_renderer(device, train=True):
    """The UNMODIFIED reference renderer (byte-compiled under oracle/_ref, SURVEY 8c shims) with the same seeded
    random-init state as our arm; falls back to the oracle port when oracle/_ref has not been built."""
    from oracle import ref_shims  # baseline infrastructure only
    torch.manual_seed(0)
    if ref_shims.available():
        mod = ref_shims.load_reference()
        r = mod.EndoSurfRenderer(copy.deepcopy(RENDER_CFG), copy.deepcopy(NET_CFG), device)
        seeded_state(r.model)
        r.train(train)
        params = [p for v in r.get_train_params().values() for p in v]
        return "reference", (lambda rays: r(rays, iter_step=ITER_STEP)), params
    from oracle import endosurf_oracle as orc
    from endosurf_b200 import EndoSurfNet
    model = EndoSurfNet(NET_CFG)
    seeded_state(model)
    ck = {k: {kk: vv.detach().clone().to(device).requires_grad_(train) for kk, vv in sd.items()}
          for k, sd in model.save_checkpoint().items()}
    net = orc.OracleNet(ck, NET_CFG)
    rc = copy.deepcopy(RENDER_CFG)
    return "port", (lambda rays: orc.render_rays(net, rc, rays, iter_step=ITER_STEP)), \
        [p for sd in ck.values() for p in sd.values()]


def reference_rays_per_s(n_rays, steps, warmup, device, mode="train", threads=None):
    """Time the reference's own PyTorch path (stock ops, fp32, TF32 off) on `device`: rays/s over `steps` steps."""
    if threads:
        torch.set_num_threads(threads)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    train = mode == "train"
    kind, render, params = _reference_renderer(device, train)
    opt = torch.optim.Adam(params, lr=5e-4) if train else None
    times = []
    for i in range(warmup + steps):
        rays = make_rays(n_rays, frame=i).to(device)
        cgt, dgt, msk = (x.to(device) for x in make_targets(n_rays, frame=i))
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if train:
            opt.zero_grad()
            o = render(rays)
            loss = train_loss(o, cgt, dgt, msk)
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                render(rays)
        if device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return n_rays / t, t, kind


def reference_latency_ms(mode, device, chunk=2048, net_chunk=80000, hw=HW, res=256):
    """BASELINE.md run B2: the UNMODIFIED reference on `device` (stock PyTorch, fp32, TF32 off) doing what our latency
    modes do - `frame`: one full 512x512 frame through its renderer in `chunk`-ray chunks with colour / depth / normal
    of every chunk converted to host numpy (trainer_endosurf.py:228-240); `grid256`: its own extract_fields over the
    256^3 grid with run_fn_split at `net_chunk` (utils.py:139-157, endosurf.py:493-494).  One warm-up chunk / 128^3
    block, then ONE timed pass (a frame takes ~18 s there)."""
    import importlib
    from oracle import ref_shims  # baseline infrastructure only
    if not ref_shims.available():
        raise RuntimeError("oracle/_ref has not been built")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    mod = ref_shims.load_reference()
    r = mod.EndoSurfRenderer(copy.deepcopy(RENDER_CFG), copy.deepcopy(NET_CFG), device)
    seeded_state(r.model)
    r.eval()

    def sync():
        if device != "cpu":
            torch.cuda.synchronize()
    if mode == "frame":
        rays_all = make_rays(0, frame=7, hw=hw, all_pixels=True).to(device)

        def run(n_rays):
            out = []
            for rs in rays_all[:n_rays].split(chunk):
                o = r(rs, iter_step=ITER_STEP)
                nrm = (o["gradients_o"] * o["weights"][:, :128, None]).sum(dim=1)
                out.append([x.detach().cpu().numpy() for x in (o["color_map"], o["depth_map"], nrm)])
                del o
            return out
        run(chunk)
        sync()
        t0 = time.perf_counter()
        run(hw * hw)
        sync()
        return (time.perf_counter() - t0) * 1e3
    ut = importlib.import_module("src.renderer.utils")
    t = torch.tensor([0.5], device=device)

    def query(pts):
        return ut.run_fn_split(lambda p: r.model.get_sdf_from_observed_space(p, t), pts, net_chunk, cpu=True)
    ut.extract_fields([-1.0] * 3, [1.0] * 3, min(64, res), query, device)
    sync()
    t0 = time.perf_counter()
    u = ut.extract_fields([-1.0] * 3, [1.0] * 3, res, query, device)
    sync()
    assert u.shape == (res, res, res)
    return (time.perf_counter() - t0) * 1e3


def reference_sample_rays(steps, warmup):
    """Bounded CPU sample: about 12k rays of CPU work in total (2-3 minutes on a 16-core host), <= 1024 rays per step
    (the reference keeps every chunk's autograd graph: 4096 rays x 128 samples do not fit a host's memory budget)."""
    per = 12288 // max(steps + warmup, 1)
    return int(min(1024, max(128, per // 128 * 128)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.ref_rays or reference_sample_rays(args.steps, args.warmup)
    mode = args.mode if args.mode in ("train", "forward") else "forward"
    val, t, kind = reference_rays_per_s(n, args.steps, max(args.warmup, 1), "cpu", mode, threads)
    line = {
        "impl": "reference", "metric": METRICS[mode], "value": val, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.rays, mode, args.precision_terms),
        "sample": {"rays_per_step": n, "note": "each step is a bounded sample of the workload in `config` (same "
                   "frame geometry, samples per ray, networks, loss); rays/s is a rate, the batch size only bounds "
                   "the CPU time and memory of the run"},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": kind,
                         "sample": f"{n} rays x (64+64 samples, 4 up-sampling steps) per step, {mode}, "
                                   f"{'the unmodified reference (oracle/_ref, byte-compiled)' if kind == 'reference' else 'oracle port'}"
                                   f", torch {torch.__version__} CPU fp32, {threads} threads"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def workload_config(n_rays, mode="train", precision_terms=3, world=1, note=None):
    what = {"train": "training step (render_rays forward + masked-mean L1 colour/depth + eikonal loss, backward, Adam)",
            "forward": "render_rays forward",
            "frame": "inference render of one full 512x512 frame",
            "grid256": "256^3 SDF grid query (extract_fields)"}[mode]
    prec = ("fp32-parity (fp16 hi/lo x3) tensor-core mode" if precision_terms == 3 else
            "single-pass fp16 tensor-core mode (precision_terms=1)")
    c = {"workload": f"{what}, {n_rays}-ray batch of one 512x512 frame, 64 coarse + 64 fine samples, "
                     f"4 up-sampling steps, deform+sdf+colour 9x256 MLPs, {prec}",
         "rays_per_step_per_gpu": n_rays, "n_samples": 64, "n_importance": 64, "up_sample_steps": 4,
         "parallelism": ("rays sharded per rank; all-reduce of the loss denominators (3 floats) before the backward and "
                         "ONE NCCL all-reduce of the flat 1.65 M-float gradient bucket after it "
                         "(endosurf_b200.distributed)" if mode == "train" else
                         "rays sharded per rank, no data-path collective"),
         "l2_policy": "per-step working set (>= 1 GiB of per-point scratch, 20 GB of plane records in training) exceeds "
                      "the 126 MB L2; ray batches rotate"}
    if mode == "train":
        c["precision"] = ("forward and activation-gradient chains: fp16 hi/lo x3 products (fp32 parity); weight/bias "
                          "gradients: own split-K tcgen05 kernel on the fp16 hi planes, fp32 accumulation; gradient "
                          "parity of this exact configuration against the oracle's autograd at 512 rays x 128 samples: "
                          "tests/test_gpu_training.py::test_timed_path_gradient_parity_512_rays") if precision_terms == 3 \
            else "every tensor-core product single-pass fp16 (tests/test_gpu_training.py::test_single_pass_fp16_*)"
    if note:
        c["note"] = note
    return c


# ------------------------------------------------------------------------------------------------ our arm
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
    from endosurf_b200 import distributed as dp

    mode = args.mode
    train = mode == "train"
    torch.manual_seed(0)
    r = EndoSurfRenderer(copy.deepcopy(RENDER_CFG), NET_CFG, device=f"cuda:{local}",
                         precision_terms=args.precision_terms)
    seeded_state(r.model)
    r.train(train)
    if mode in ("frame", "grid256"):
        return run_latency(args, r, dev, world, rank)
    params = [p for v in r.get_train_params().values() for p in v]
    opt = torch.optim.Adam(params, lr=5e-4) if train else None
    bucket = dp.FlatGradBucket(params).bind(r) if train else None
    R, K, W = args.rays, args.steps, args.warmup
    n_batches = min(K + W, 16)
    frames = [(rank * 17 + i) % 60 for i in range(n_batches)]
    host = [make_rays(R, frame=f, seed=rank).pin_memory() for f in frames]
    host_t = [tuple(x.pin_memory() for x in make_targets(R, frame=f, seed=rank)) for f in frames]
    devb = [h.to(dev) for h in host]
    devt = [tuple(x.to(dev) for x in t) for t in host_t]
    out_c = torch.empty(R, 3).pin_memory()
    out_d = torch.empty(R, 1).pin_memory()
    out_l = torch.empty(()).pin_memory()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(rays, cgt, dgt, msk):
        if train:
            o = r(rays, iter_step=ITER_STEP)
            terms, eps = dp.render_loss_terms(r, o, cgt, dgt, msk, msk)
            logs = dp.dp_backward(bucket, terms, eps)  # world == 1: no collective, same code path
            opt.step()
            return o, logs["loss"]
        with torch.no_grad():
            return r.render_rays(rays, iter_step=ITER_STEP), None

    for i in range(W):
        step(devb[i % n_batches], *devt[i % n_batches])
    r.sync_check()
    # ---------------- device-resident timing (value); CUDA events around every library kernel of the timed region
    r.profile(True)
    r.profile_read()
    launches0 = r.launch_count()
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
    launches = r.launch_count() - launches0
    prof = r.profile_read()
    r.profile(False)
    # ---------------- end to end: pinned host rays (+ targets) in, colour + depth (+ loss) back to the host, every step
    sync_all()
    t0 = time.perf_counter()
    for i in range(K):
        j = (W + i) % n_batches
        rays = host[j].to(dev, non_blocking=True)
        tg = tuple(x.to(dev, non_blocking=True) for x in host_t[j])
        o, loss = step(rays, *tg)
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
        peak, peak_src, _ = measured_peaks()
        g = prof["geometry_chain"]
        pts_per_launch = g["points"] / max(g["launches"], 1)
        alg_flops_launch = pts_per_launch * 2.0 * (4 * D_MAC + 2 * S_MAC)
        ms_launch = g["ms"] / max(g["launches"], 1)
        achieved = alg_flops_launch / (ms_launch * 1e-3) / 1e12 if ms_launch > 0 else 0.0
        kern_ms = {k: round(v["ms"] / K, 4) for k, v in prof.items()}
        fpr = train_flops_per_ray(64, 64, 4) if train else flops_per_ray(64, 64, 4)
        h2d = R * 9 * 4 + (R * 5 * 4 if train else 0)
        d2h = R * 4 * 4 + (4 if train else 0)
        terms_note = ("3 fp16 MMAs per product (hi/lo split)" if args.precision_terms == 3 else "1 fp16 MMA per product")
        line = {
            "metric": METRICS[mode], "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 (fp16 hi/lo x3 tensor-core products, fp32 accumulate)" if args.precision_terms == 3
                      else "fp16 (single-pass tensor-core products, fp32 accumulate)"),
            "data": "synthetic",
            "config": workload_config(R, mode, args.precision_terms, world),
            "launch_mode": "eager library launches, no CUDA graph",
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_step},
            "gpu_launches": int(launches),
            "gpu_launches_per_step": launches / K,
            "clocks": clk,
            "roofline": {"bound": "tensor",
                         "kernel": "mlp_chain_kernel<CHAIN_SDF,TANGENT,fwd" + (",STASH>" if train else ">") +
                                   " (forward geometry chain: deform+sdf MLPs with 3 forward-mode tangent rows per "
                                   "point" + ("; training variant keeping the plane records" if train else "") + ")",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": ncu_traffic(mode),
                         "tensor_pipe_pct_ncu": ncu_capture(mode, "tensor_pipe_pct"),
                         "algorithmic_flops_per_launch": alg_flops_launch, "points_per_launch": pts_per_launch,
                         "ms_per_launch": ms_launch,
                         "note": f"algorithmic = 2*(4D+2S) FLOP/point (SURVEY 8d); the kernel issues {terms_note} "
                                 "and 4D+4S+feat (forward-mode normals)",
                         "kernel_ms_per_step": kern_ms,
                         "kernel_ms_sum": round(sum(kern_ms.values()), 3),
                         "kernel_timing": "CUDA events around every library kernel group over the timed region itself",
                         "step_algorithmic_tflops": value / world * fpr / 1e12},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rm = mode
            v, tt, kind = reference_rays_per_s(args.ref_rays or 512, 3, 1, "cpu", rm, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": "rays/s", "cores": threads, "kind": kind,
                "sample": f"{args.ref_rays or 512} rays of the same workload ({rm}), "
                          f"{'the unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port'}, PyTorch "
                          f"{torch.__version__} CPU fp32, mean of 3 steps of {tt:.1f} s after one warm-up step"}
        if world == 1 and not args.no_gpu_reference:
            # the north-star denominator: the reference's own PyTorch path on this same GPU (stock ops, fp32, TF32 off)
            del devb, devt
            r.release_workspace()
            torch.cuda.empty_cache()
            try:
                v, tt, kind = reference_rays_per_s(R, 3, 1, f"cuda:{local}", mode)
                line["gpu_reference"] = {
                    "value": v, "unit": "rays/s", "ms_per_step": 1e3 * tt, "rays_per_step": R, "kind": kind,
                    "what": f"{'the unmodified reference (oracle/_ref)' if kind == 'reference' else 'oracle port'} on "
                            f"cuda:{local}, stock PyTorch {torch.__version__} ops, fp32, TF32 off, same rays / state / "
                            "loss, 1 warm-up + 3 timed steps",
                    "speedup_value": value / v, "speedup_e2e": e2e / v}
            except Exception as e:  # never lose the main line to the baseline
                line["gpu_reference"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_latency(args, r, dev, world, rank):
    """BASELINE configs[4] (latency mode, one GPU): `frame` = a full 512x512 frame rendered in ray chunks with the
    per-chunk outputs copied to the host, as the reference's eval loop does (trainer_endosurf.py:230-240,
    eval.ray_chunk = 2048 in base_pull.yml:37); `grid256` = the 256^3 SDF grid of extract_fields (utils.py:139-157)."""
    K, W = args.steps, args.warmup
    peak, peak_src, _ = measured_peaks()
    clocks = ClockSampler(dev.index or 0)
    line = {"metric": METRICS[args.mode], "n_gpus": 1, "steps": K, "warmup": W, "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "dtype": "f32 (fp16 hi/lo x3 tensor-core products, fp32 accumulate)"}
    if args.mode == "frame":
        rays_all = make_rays(0, frame=7, all_pixels=True).to(dev)
        results = {}
        for chunk in (2048, args.frame_chunk):
            host_rgb = torch.empty(HW * HW, 3).pin_memory()
            host_dep = torch.empty(HW * HW, 1).pin_memory()
            host_nrm = torch.empty(HW * HW, 3).pin_memory()

            def frame():
                with torch.no_grad():
                    for r0 in range(0, HW * HW, chunk):
                        o = r.render_rays(rays_all[r0:r0 + chunk], iter_step=ITER_STEP, perturb_overwrite=False)
                        nrm = (o["gradients_o"] * o["weights"][:, :, None]).sum(1)
                        host_rgb[r0:r0 + chunk].copy_(o["color_map"], non_blocking=True)
                        host_dep[r0:r0 + chunk].copy_(o["depth_map"], non_blocking=True)
                        host_nrm[r0:r0 + chunk].copy_(nrm, non_blocking=True)
                torch.cuda.synchronize()
            for _ in range(W):
                frame()
            if chunk == 2048:
                clocks.start()
            t0 = time.perf_counter()
            for _ in range(K):
                frame()
            results[chunk] = (time.perf_counter() - t0) / K * 1e3
        clk = clocks.stop()
        ms = results[2048]
        fl = HW * HW * flops_per_ray(64, 64, 4)
        line.update({"value": ms, "unit": "ms/frame", "ms_per_step": ms, "clocks": clk,
                     "config": workload_config(2048, "frame"),
                     "e2e": {"value": ms, "unit": "ms/frame", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": HW * HW * 7 * 4,
                             "note": "rays resident on the device like Dataset.rays (dataset.py:107); colour, depth and "
                                     "normal of every chunk copied to pinned host memory inside the timed region"},
                     "rays_per_s": HW * HW / (ms * 1e-3),
                     "large_chunk": {"rays_per_chunk": args.frame_chunk, "ms_per_frame": results[args.frame_chunk],
                                     "rays_per_s": HW * HW / (results[args.frame_chunk] * 1e-3)},
                     "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peak,
                                  "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / peak, "peak_source": peak_src,
                                  "traffic": None, "note": "whole frame, algorithmic forward FLOPs (SURVEY 8d)"},
                     "gpu_launches": int(r.launch_count())})
    else:
        res = 256
        t = torch.tensor([0.5], device=dev)
        lo, hi = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
        for _ in range(W):
            r.extract_fields(t, lo, hi, res, cpu=False)
        torch.cuda.synchronize()
        clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            u = r.extract_fields(t, lo, hi, res, cpu=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        t0 = time.perf_counter()
        for _ in range(K):
            u_host = r.extract_fields(t, lo, hi, res, cpu=True)  # volume to the host, what marching cubes consumes
        ms_e2e = (time.perf_counter() - t0) / K * 1e3
        clk = clocks.stop()
        fl = float(res) ** 3 * 2.0 * (D_MAC + S_MAC)
        line.update({"value": ms, "unit": "ms/grid", "ms_per_step": ms, "clocks": clk,
                     "config": {"workload": "256^3 = 16,777,216-point SDF grid at one time value, deform+sdf 9x256 MLPs, "
                                            "fp32-parity tensor-core mode, points generated on the device"},
                     "e2e": {"value": ms_e2e, "unit": "ms/grid", "h2d_bytes_per_step": 28,
                             "d2h_bytes_per_step": res ** 3 * 4, "note": "volume copied to host numpy"},
                     "roofline": {"bound": "tensor", "achieved": fl / (ms * 1e-3) / 1e12, "peak": peak,
                                  "unit": "TFLOP/s", "frac": fl / (ms * 1e-3) / 1e12 / peak, "peak_source": peak_src,
                                  "traffic": None, "algorithmic_flops_per_launch": fl,
                                  "note": "33.69 TFLOP per grid (SURVEY 8d); sdf-query chain, 3 fp16 MMAs per product"},
                     "gpu_launches": int(r.launch_count()), "sdf_range": [float(u_host.min()), float(u_host.max())]})
    if not args.no_gpu_reference:
        r.release_workspace()
        torch.cuda.empty_cache()
        try:
            ref_ms = reference_latency_ms(args.mode, f"cuda:{dev.index or 0}")
            line["gpu_reference"] = {
                "value": ref_ms, "unit": line["unit"], "kind": "reference",
                "what": "the unmodified reference (oracle/_ref) on the same GPU, stock PyTorch "
                        f"{torch.__version__} ops, fp32, TF32 off, one timed pass after a warm-up chunk: " +
                        ("its renderer over the same 512x512 frame in 2048-ray chunks, per-chunk colour / depth / normal "
                         "to host numpy (trainer_endosurf.py:228-240)" if args.mode == "frame" else
                         "its extract_fields over the 256^3 grid, run_fn_split at net_chunk 80000, blocks to host numpy "
                         "(utils.py:139-157)"),
                "speedup_value": ref_ms / line["value"], "speedup_e2e": ref_ms / line["e2e"]["value"]}
        except Exception as e:  # never lose the main line to the baseline
            line["gpu_reference"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


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
    ap.add_argument("--rays", type=int, default=4096,
                    help="rays per step per GPU (BASELINE configs[1]: 4096; configs[3]: 8192 per GPU on 8 GPUs)")
    ap.add_argument("--ref-rays", type=int, default=0,
                    help="rays per step of the bounded CPU sample (0: sized from --steps/--warmup)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--precision-terms", type=int, default=3, choices=[1, 3],
                    help="3: fp32-parity hi/lo split (default); 1: single-pass fp16 (BASELINE configs[2])")
    ap.add_argument("--frame-chunk", type=int, default=16384, help="--mode frame: second, larger ray chunk")
    ap.add_argument("--mode", default="train", choices=["train", "forward", "frame", "grid256"],
                    help="train: BASELINE.json's metric (training rays/s); forward: inference render_rays; "
                         "frame / grid256: configs[4] latencies")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
