"""CPU check of the synthetic Dataset stand-in against the UNMODIFIED reference trainer and the
reference's own renderer: if the reference trains on it on the CPU, the member surface is right, and the GPU test
(tests/test_gpu_trainer.py) only swaps the renderer."""
import copy
import importlib

import numpy as np
import pytest
import torch
import yaml

from conftest import load_cfg


def test_synthetic_dataset_surface():
    from endosurf_b200.harness import SyntheticDataset
    d = SyntheticDataset({"normalize_time": True, "n_frames": 5, "w": 24, "h": 16, "device": "cpu"})
    assert d.rays.shape == (5, 16, 24, 9) and d.colors.shape == (5, 16, 24, 3) and d.depths.shape == (5, 16, 24, 1)
    assert d.poses.shape == (5, 4, 4) and d.intrinsics.shape == (5, 4, 4) and d.bbox_minmax.shape == (5, 3, 2)
    assert set(d.list_train).isdisjoint(d.list_test) and d.n_train + d.n_test == 5
    assert 0.0 < d.near < d.far < 2.0
    b = d.get_train_batch_data_by_index(ray_batch=40)
    assert {k: tuple(v.shape) for k, v in b.items()} == {
        "color": (40, 3), "rays": (40, 9), "depth": (40, 1), "mask": (40, 1), "color_mask": (40, 1),
        "depth_mask": (40, 1)}
    # rays: unit directions from the camera centre, time in [0, 1], z-depth consistent with the renderer's o + d/d_z * z
    r = b["rays"]
    assert torch.allclose(r[:, 3:6].norm(dim=-1), torch.ones(40), atol=1e-5)
    p = r[:, :3] + r[:, 3:6] / r[:, 5:6] * b["depth"]
    hit = b["mask"][:, 0] > 0
    t = r[:, 8]
    rad = 0.8 + 0.03 * torch.sin(2 * np.pi * t)
    assert torch.allclose(p.norm(dim=-1)[hit], rad[hit], atol=1e-4)
    f = d.get_frame_data_by_index(d.list_test[:1])
    assert f["rays"].shape == (1, 16, 24, 9)


def test_reference_trainer_accepts_the_synthetic_dataset(tmp_path):
    from oracle import ref_shims
    if not ref_shims.available():
        pytest.skip("oracle/_ref (byte-compiled reference) has not been built")
    if torch.cuda.is_available():
        pytest.skip("covered on the GPU by tests/test_gpu_trainer.py")
    ref_shims.install_shims()
    tb = importlib.import_module("src.trainer.trainer_basic")
    te = importlib.import_module("src.trainer.trainer_endosurf")
    ref_renderer = importlib.import_module("src.renderer.endosurf").EndoSurfRenderer
    from endosurf_b200.harness import patch_reference_trainer
    patch_reference_trainer(te, tb, n_frames=4, hw=(24, 24))
    te.EndoSurfRenderer = ref_renderer  # CPU: the reference's own renderer; only the dataset is the stand-in
    base = load_cfg()
    cfg = {
        "exp": {"project_name": "endosurf", "exp_name": "harness", "exp_dir": str(tmp_path / "logs")},
        "data": {"info_dir": "synthetic", "normalize_time": True},
        "render": dict(copy.deepcopy(base["render"]), n_samples=8, n_importance=8),
        "train": {"n_iter": 10, "ray_batch": 32, "mask_guided_ray_sampling": True, "color_loss_weight": 1.0,
                  "depth_loss_weight": 1.0, "sdf_loss_weight": 1.0, "angle_loss_weight": 0.1,
                  "eikonal_loss_weight": 0.1, "surf_neig_loss_weight": 0.1, "surf_neig_rad": 0.1, "resume": False,
                  "optim": {"lr": 5e-4, "lr_alpha": 0.05, "warm_up_end": 5}, "eval": {"ray_chunk": 2048}},
        "net": copy.deepcopy(base["net"]),
        "log": {"summary_writer": {"type": "tensorboard"}, "i_eval": 0, "i_save": 0},
    }
    path = tmp_path / "cfg.yml"
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    torch.manual_seed(0)
    np.random.seed(0)
    trainer = te.EndoSurfTrainer(str(path))
    losses = [trainer.train_step(global_step=i) for i in (1, 2)]
    assert all(np.isfinite(losses)), losses
