#!/usr/bin/env python
"""BASELINE configs[2]: the full EndoSurf training loop on a synthetic 512x512x60-frame scene, one GPU.

The UNMODIFIED reference trainer (byte-compiled under oracle/_ref; ``EndoSurfTrainer.train_step`` +
``update_learning_rate``, reference src/trainer/trainer_endosurf.py:94-181) is driven three times on the synthetic
``Dataset`` stand-in of ``endosurf_b200.harness`` at full size (60 frames of 512x512): with this repository's renderer
in the fp32-parity mode, in the single-pass fp16 mode (``precision_terms=1``, the config's "bf16" arm) and with the
reference's own renderer (stock PyTorch fp32 on the same GPU).  One iteration = everything the trainer does per step:
ray sampling from the dataset, render_rays, the colour / depth / eikonal losses, ``errorondepth`` (sdf + angle losses),
``surface_neighbour_error``, backward, Adam, the 14 ``.item()`` logging syncs.  Wall-clock per iteration, device
synchronised on both sides.  This is measurement tooling (it executes the reference as the caller and as the baseline);
the product never imports it.

    python tools/trainer_loop_bench.py [--rays 4096] [--samples 64] [--steps 10] > gpurun_out/trainer_loop.json
"""
import argparse
import copy
import importlib
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class _Recorder:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, global_step):
        self.scalars[tag] = float(value.item() if torch.is_tensor(value) else value)


def make_trainer(kind, rays, ns, ni, n_frames, hw, workdir):
    from conftest import load_cfg
    from oracle import ref_shims
    ref_shims.install_shims()
    tb = importlib.import_module("src.trainer.trainer_basic")
    te = importlib.import_module("src.trainer.trainer_endosurf")
    from endosurf_b200.harness import patch_reference_trainer
    import endosurf_b200
    patch_reference_trainer(te, tb, n_frames=n_frames, hw=(hw, hw))
    if kind == "fp16x1":
        te.EndoSurfRenderer = lambda *a, **k: endosurf_b200.EndoSurfRenderer(*a, precision_terms=1, **k)
    elif kind == "reference":
        te.EndoSurfRenderer = importlib.import_module("src.renderer.endosurf").EndoSurfRenderer
    base = load_cfg()
    cfg = {
        "exp": {"project_name": "endosurf", "exp_name": f"loop_{kind}", "exp_dir": os.path.join(workdir, kind)},
        "data": {"info_dir": "synthetic", "normalize_time": True},
        "render": dict(copy.deepcopy(base["render"]), n_samples=ns, n_importance=ni),
        # loss weights / optimiser of configs/endosurf/baseline/base_pull.yml:19-38
        "train": {"n_iter": 100000, "ray_batch": rays, "mask_guided_ray_sampling": True, "color_loss_weight": 1.0,
                  "depth_loss_weight": 1.0, "sdf_loss_weight": 1.0, "angle_loss_weight": 0.1,
                  "eikonal_loss_weight": 0.1, "surf_neig_loss_weight": 0.1, "surf_neig_rad": 0.1, "resume": False,
                  "optim": {"lr": 5e-4, "lr_alpha": 0.05, "warm_up_end": 5000}, "eval": {"ray_chunk": 2048}},
        "net": copy.deepcopy(base["net"]),
        "log": {"summary_writer": {"type": "tensorboard"}, "i_eval": 0, "i_save": 0},
    }
    os.makedirs(cfg["exp"]["exp_dir"], exist_ok=True)
    path = os.path.join(workdir, f"cfg_{kind}.yml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    torch.manual_seed(0)
    np.random.seed(0)
    return te.EndoSurfTrainer(path)


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=64, help="coarse = fine samples per ray")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--ref-steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--hw", type=int, default=512)
    ap.add_argument("--kinds", default="fp16x3,fp16x1,reference")
    args = ap.parse_args()
    json_out = os.fdopen(os.dup(1), "w")  # the reference trainer prints banners on stdout: they go to stderr
    os.dup2(2, 1)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    work = tempfile.mkdtemp(prefix="es_loop_")
    res, init = {}, None
    for kind in args.kinds.split(","):
        tr = make_trainer(kind, args.rays, args.samples, args.samples, args.frames, args.hw, work)
        if init is None:
            init = copy.deepcopy(tr.renderer.save_checkpoint())
        else:
            tr.renderer.load_checkpoint(init)
        tr.writer = _Recorder()
        tr.renderer.train()
        steps = args.ref_steps if kind == "reference" else args.steps
        warm = 1 if kind == "reference" else args.warmup
        losses = []
        for it in range(1, warm + steps + 1):
            if it == warm + 1:
                _sync()
                t0 = time.perf_counter()
            torch.manual_seed(1000 + it)
            np.random.seed(1000 + it)
            losses.append(float(tr.train_step(global_step=it)))
            tr.update_learning_rate(it)
        _sync()
        dt = (time.perf_counter() - t0) / steps
        if kind != "reference":
            tr.renderer.sync_check()
        res[kind] = {"ms_per_iteration": 1e3 * dt, "rays_per_s": args.rays / dt, "iterations_timed": steps,
                     "warmup": warm, "loss_first": losses[0], "loss_last": losses[-1]}
        del tr
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    line = {"metric": "full EndoSurf training-loop iterations (unmodified reference trainer), rays/sec",
            "config": {"workload": f"EndoSurfTrainer.train_step + update_learning_rate, synthetic {args.hw}x{args.hw}x"
                                   f"{args.frames}-frame scene, ray_batch {args.rays}, {args.samples}+{args.samples} "
                                   "samples, all six loss terms of base_pull.yml, Adam, logging syncs included"},
            "unit": "rays/s", "results": res, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0) if torch.cuda.is_available() else None}
    if "reference" in res:
        for k in res:
            if k != "reference":
                res[k]["speedup_vs_reference_renderer"] = res[k]["rays_per_s"] / res["reference"]["rays_per_s"]
    print(json.dumps(line), file=json_out, flush=True)


if __name__ == "__main__":
    main()
