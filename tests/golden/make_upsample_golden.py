"""Pin the hierarchical-sampling trace to the UNMODIFIED reference's own ``up_sample`` / ``cat_z_vals``
(src/renderer/endosurf.py:221-287), replayed exactly as ``render_rays`` chains them (:85-110).

    python tests/golden/make_upsample_golden.py       # build container only (needs /root/reference)

Writes ``upsample_ref.npz``: for every up-sampling step i the inputs (z_vals, sdf) and the reference's new samples, plus
the final z_vals.  Also checks the oracle's trace against it and refuses to write if they disagree.
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from make_golden import import_reference  # noqa: E402
from conftest import load_cfg, load_ckpt  # noqa: E402
from oracle import endosurf_oracle as orc  # noqa: E402


def main():
    Renderer = import_reference()
    cfg, ckpt = load_cfg(), load_ckpt()
    rc = copy.deepcopy(cfg["render"])
    rc.update(n_samples=32, n_importance=32, perturb=False)
    torch.manual_seed(0)
    r = Renderer(rc, cfg["net"], device="cpu")
    r.load_checkpoint(ckpt)
    g = dict(np.load(os.path.join(HERE, "render_r32_s32_i32_it25k.npz")))
    rays = torch.from_numpy(g["rays"])
    n_rays, ns, ni, steps = rays.shape[0], 32, 32, rc["up_sample_steps"]
    rays_o, rays_d, time = rays[..., :3], rays[..., 3:6], rays[..., 8]
    rays_d_z = rays_d / (rays_d[..., 2:] + 1e-6)
    onet = orc.OracleNet(ckpt, cfg["net"])
    z_vals = orc.coarse_z_vals(rays, ns)  # (pinned against the reference by make_golden.py through render_rays)
    save = {"rays": rays.numpy()}
    with torch.no_grad():
        pts = (rays_o[:, None, :] + rays_d_z[:, None, :] * z_vals[..., :, None]).reshape(-1, 3)
        t = time[..., None, None].expand(n_rays, ns, 1).reshape(-1, 1)
        sdf = r.model.get_sdf_from_observed_space(pts, t).reshape(n_rays, ns)
        for i in range(steps):
            new_z = r.up_sample(rays_o, rays_d, z_vals, sdf, ni // steps, 64 * 2 ** i)
            save[f"up{i}_z"], save[f"up{i}_sdf"], save[f"up{i}_new_z"] = z_vals.numpy(), sdf.numpy(), new_z.numpy()
            z_vals, sdf = r.cat_z_vals(rays_o, rays_d, time, z_vals, new_z, sdf, last=(i + 1 == steps))
    save["z_final"] = z_vals.numpy()
    # the oracle's restatement on the SAME per-step inputs
    worst = 0.0
    for i in range(steps):
        z, s = torch.from_numpy(save[f"up{i}_z"]), torch.from_numpy(save[f"up{i}_sdf"])
        o = orc.up_sample(rays_o, rays_d, z, s, ni // steps, 64 * 2 ** i)
        worst = max(worst, (o - torch.from_numpy(save[f"up{i}_new_z"])).abs().max().item())
    print(f"oracle up_sample vs reference on the reference's own inputs: max abs err {worst:.3e}")
    assert worst < 1e-5, "oracle/endosurf_oracle.py::up_sample disagrees with the reference"
    np.savez(os.path.join(HERE, "upsample_ref.npz"), **save)
    with open(os.path.join(HERE, "ORACLE_PIN.txt"), "a") as f:
        f.write(f"{'up_sample (reference trace, 4 steps)':40s} {worst:.3e}\n")


if __name__ == "__main__":
    main()
