"""Synthetic stand-in for the reference ``Dataset`` so that its trainer runs unchanged on this renderer.

The reference trainer (``src/trainer/trainer_endosurf.py``) needs, besides the renderer, a ``Dataset`` built from a
pre-processed EndoNeRF / SCARED scene on disk (``src/dataset/dataset.py:22-181``; images via imageio, point clouds via
open3d).  There is no dataset in this environment, so :class:`SyntheticDataset` fabricates a scene with the same
member surface and tensor shapes: a pinhole endoscope at (0, 0, -1.5) looking down +z at a slowly deforming sphere
inside the unit ball, ``n_frames`` frames of ``h x w`` pixels with colour, z-depth and masks, rays ``[n_frames, h, w, 9]
= (o, d, near, far, time)`` exactly as ``Dataset.rays`` (``dataset.py:96-107``).

Members mirrored (same names, shapes and dtypes): ``dset_name, scene_name, n_frames, w, h, depth_scale, intrinsics,
poses, bbox_minmax, colors, depths, depth_masks, color_masks, masks, rays, near, far, list_train, list_test, n_train,
n_test, ray_importance_maps, get_train_batch_data_by_index(), get_frame_data_by_index()``.

``patch_reference_trainer`` swaps this class and ``endosurf_b200.EndoSurfRenderer`` into an imported reference trainer
module; it touches nothing else of the trainer.
"""
from __future__ import annotations

import math

import numpy as np
import torch


class SyntheticDataset(object):
    """Drop-in for ``src.dataset.Dataset`` (reference dataset.py:22-181) on a synthetic deforming-sphere scene."""

    def __init__(self, dset_cfg, device=None):
        self.dset_cfg = dset_cfg
        self.device = device or dset_cfg.get("device", "cuda" if torch.cuda.is_available() else "cpu")
        self.dset_name = "endonerf"
        self.scene_name = dset_cfg.get("scene_name", "synthetic_sphere")
        self.n_frames = int(dset_cfg.get("n_frames", 60))
        self.w = int(dset_cfg.get("w", 512))
        self.h = int(dset_cfg.get("h", 512))
        self.depth_scale = 1.0
        dev = self.device
        f = 1.2 * self.w
        K = torch.eye(4, device=dev)
        K[0, 0] = K[1, 1] = f
        K[0, 2], K[1, 2] = self.w / 2.0, self.h / 2.0
        pose = torch.eye(4)
        pose[2, 3] = -1.5
        self.intrinsics = K[None].expand(self.n_frames, 4, 4).contiguous()
        self.poses = pose[None].expand(self.n_frames, 4, 4).contiguous()  # camera to world (kept on the host, :53)
        self.bbox_minmax = np.tile(np.array([[-1.0, 1.0]] * 3)[None], (self.n_frames, 1, 1))

        rays = self.get_rays(self.intrinsics, self.poses.to(dev), self.w, self.h)  # [n, h, w, 6]
        normalize_time = dset_cfg.get("normalize_time", True)
        ts = torch.linspace(0.0, 1.0, self.n_frames, device=dev) if normalize_time else \
            torch.arange(self.n_frames, device=dev).float()
        o, d = rays[..., :3], rays[..., 3:6]
        # analytic scene: sphere of radius r(t) around the origin, z-depth along d / d_z like the reference's depth maps
        r_t = 0.8 + 0.03 * torch.sin(2.0 * math.pi * ts)[:, None, None]
        b = (o * d).sum(-1)
        disc = b * b - ((o * o).sum(-1) - r_t * r_t)
        hit = disc > 0
        s = -b - torch.sqrt(disc.clamp_min(0.0))
        depth = torch.where(hit, s * d[..., 2], torch.zeros_like(s))[..., None]
        p = o + s[..., None] * d
        normal = p / p.norm(dim=-1, keepdim=True).clamp_min(1e-6)
        shade = (0.55 + 0.45 * normal[..., 2:3].abs())
        tex = 0.5 + 0.5 * torch.sin(6.0 * p + torch.tensor([0.0, 2.0, 4.0], device=dev))
        self.colors = torch.where(hit[..., None], (shade * tex).clamp(0, 1), torch.zeros_like(tex))
        self.depths = depth / self.depth_scale
        dvals = self._tensor2array(self.depths[hit[..., None].expand_as(self.depths)])
        self.near = float(np.percentile(dvals, 3.0))
        self.far = float(np.percentile(dvals, 99.5))
        self.depth_masks = torch.bitwise_and(self.depths > self.near, self.depths < self.far) * 1.0
        self.color_masks = hit[..., None].float()
        self.masks = self.depth_masks * self.color_masks
        bounds = torch.tensor([self.near, self.far], device=dev)
        self.bds = bounds[None, None, None, :].expand(self.n_frames, self.h, self.w, 2)
        self.ts = ts[:, None, None, None].expand(self.n_frames, self.h, self.w, 1)
        self.rays = torch.cat([rays, self.bds, self.ts], -1)  # [n_frames, h, w, 9]
        ids = list(range(self.n_frames))
        self.list_test = ids[::8]
        self.list_train = [i for i in ids if i not in self.list_test]
        self.n_train, self.n_test = len(self.list_train), len(self.list_test)
        self.ray_importance_maps = self._ray_sampling_importance_from_masks(self.masks)
        self.vcam = None
        self.render_option = None

    # ---- the two accessors the trainer calls (dataset.py:120-181) --------------------------------------------------
    def get_train_batch_data_by_index(self, id_train=None, ray_batch=1024, mask_guided_ray_sampling=True):
        if id_train is None:
            id_train = int(np.random.choice(self.list_train))
        else:
            assert id_train in self.list_train, f"ID {id_train} is not in training list!"
        color_masks = self.color_masks[id_train]
        valid = torch.nonzero(color_masks[..., 0].flatten() == 1.0).squeeze(-1)
        if mask_guided_ray_sampling:
            w = self.ray_importance_maps[id_train][..., 0].flatten()[valid]
            inds = self._importance_sampling_coords(w.unsqueeze(0), ray_batch, device=self.device).squeeze(0)
            inds = inds.clamp(0, valid.numel() - 1)
        else:
            inds = torch.as_tensor(np.random.choice(valid.numel(), size=[ray_batch], replace=False), device=self.device)
        pix = valid[inds]

        def take(x):
            return x[id_train].reshape(self.h * self.w, -1)[pix]

        return {"color": take(self.colors), "rays": take(self.rays), "depth": take(self.depths),
                "mask": take(self.masks), "color_mask": take(self.color_masks), "depth_mask": take(self.depth_masks)}

    def get_frame_data_by_index(self, id):
        return {"color": self.colors[id], "rays": self.rays[id], "depth": self.depths[id], "mask": self.masks[id],
                "color_mask": self.color_masks[id], "depth_mask": self.depth_masks[id]}

    # ---- helpers with the reference's names -------------------------------------------------------------------------
    def get_rays(self, intrinsics, poses, w, h):
        """Pixel-centre-free pinhole rays, direction normalised before the pose rotation (dataset.py:217-235)."""
        inv = torch.inverse(intrinsics)
        px, py = torch.meshgrid(torch.linspace(0, w - 1, w, device=self.device),
                                torch.linspace(0, h - 1, h, device=self.device), indexing="ij")
        p = torch.stack([px.t(), py.t(), torch.ones(h, w, device=self.device)], -1)  # [h, w, 3]
        out = []
        for i in range(intrinsics.shape[0]):
            d = torch.matmul(inv[i, None, None, :3, :3], p[..., None]).squeeze(-1)
            d = d / torch.linalg.norm(d, ord=2, dim=-1, keepdim=True)
            d = torch.matmul(poses[i, None, None, :3, :3], d[..., None]).squeeze(-1)
            o = poses[i, None, None, :3, 3].expand(d.shape)
            out.append(torch.cat([o, d], -1))
        return torch.stack(out, 0)

    @staticmethod
    def _importance_sampling_coords(weights, N_samples, det=False, device="cuda"):
        weights = weights + 1e-5
        cdf = torch.cumsum(weights / torch.sum(weights, -1, keepdim=True), -1)
        if det:
            u = torch.linspace(0.0, 1.0, steps=N_samples, device=device).expand(list(cdf.shape[:-1]) + [N_samples])
        else:
            u = torch.rand(list(cdf.shape[:-1]) + [N_samples], device=device)
        return torch.searchsorted(cdf, u.contiguous(), right=True)

    @staticmethod
    def _ray_sampling_importance_from_masks(masks):
        freq = (1.0 - masks).sum(0)
        p = freq / torch.sqrt((torch.pow(freq, 2)).sum()).clamp_min(1e-12)
        return masks * (1.0 + p)

    @staticmethod
    def _array2tensor(array, device="cuda", dtype=torch.float32):
        return torch.tensor(array, dtype=dtype, device=device)

    @staticmethod
    def _tensor2array(tensor):
        return tensor.detach().cpu().numpy()


def patch_reference_trainer(trainer_module, basic_module=None, n_frames=None, hw=None):
    """Point an imported reference trainer module (``src.trainer.trainer_endosurf``) at this package's renderer and at
    the synthetic dataset.  The trainer code itself is untouched."""
    from .renderer import EndoSurfRenderer
    trainer_module.EndoSurfRenderer = EndoSurfRenderer
    ds = SyntheticDataset
    if n_frames is not None or hw is not None:
        class _Sized(SyntheticDataset):
            def __init__(self, dset_cfg, device=None):
                cfg = dict(dset_cfg)
                if n_frames is not None:
                    cfg["n_frames"] = n_frames
                if hw is not None:
                    cfg["h"], cfg["w"] = hw
                super().__init__(cfg, device)
        ds = _Sized
    for mod in (trainer_module, basic_module):
        if mod is not None and hasattr(mod, "Dataset"):
            mod.Dataset = ds
    return ds
