"""CPU/PyTorch restatement of EndoSurf's per-ray volume-rendering path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  The product path (``endosurf_b200``) never does.

It restates, in functional form and on plain tensors, the algorithm of the
reference renderer (all citations relative to the reference repository):

* ``src/renderer/encoder.py:40-54``      -> :func:`freq_encode`
* ``src/renderer/utils.py:57-58``        -> :func:`fold_weight_norm` (old-style weight norm)
* ``src/renderer/endosurf.py:724-738``   -> :func:`deform_mlp`
* ``src/renderer/endosurf.py:773-786``   -> :func:`sdf_mlp`
* ``src/renderer/endosurf.py:828-842``   -> :func:`color_mlp`
* ``src/renderer/endosurf.py:570-689``   -> :class:`OracleNet` (sdf / gradients / jacobian / forward)
* ``src/renderer/utils.py:194-210``      -> :func:`sphere_intersection`
* ``src/renderer/utils.py:160-191``      -> :func:`sample_pdf_det`
* ``src/renderer/endosurf.py:221-266``   -> :func:`up_sample`
* ``src/renderer/endosurf.py:268-287``   -> :func:`cat_z_vals`
* ``src/renderer/endosurf.py:134-213``   -> :func:`render_core`
* ``src/renderer/endosurf.py:60-132``    -> :func:`render_rays`
* ``src/renderer/endosurf.py:289-449``   -> :func:`errorondepth`, :func:`ray_marching`, :func:`secant`

Parity pin: the reference ships no tests / golden vectors, so
the pin is the reference itself, imported in the build container by
``tests/golden/make_golden.py``; that script checks this restatement against the
reference on the same seeded inputs and writes the fixtures under
``tests/golden/`` that travel to the GPU box.

Gradients are taken with ``torch.autograd.grad(create_graph=True)`` exactly like the
reference, so ``render_rays`` here is differentiable to all parameters and serves as
the training oracle as well.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

SQRT2 = float(np.sqrt(2))


# --------------------------------------------------------------------------- primitives
def freq_encode(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]; ref encoder.py:40-54."""
    out = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, dim=-1)


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """Old-API ``nn.utils.weight_norm`` (dim=0): W = g * v / ||v||_row; ref utils.py:57-58."""
    return v * (g / torch.linalg.norm(v, dim=1, keepdim=True))


def _layers(sd: Dict[str, torch.Tensor]) -> List[tuple]:
    n = 0
    while f"net.{n}.bias" in sd:
        n += 1
    return [(fold_weight_norm(sd[f"net.{l}.weight_g"], sd[f"net.{l}.weight_v"]), sd[f"net.{l}.bias"])
            for l in range(n)]


def _mlp(layers, inp, skips, act, linear=F.linear):
    """NeRF/IDR style skip MLP body shared by the three nets: at a skip layer the
    running activation is concatenated with the *network input* and divided by sqrt(2)
    (ref endosurf.py:732-737, 778-783, 835-840)."""
    h = inp
    n = len(layers)
    for l, (w, b) in enumerate(layers):
        if l in skips:
            h = torch.cat([h, inp], -1) / SQRT2
        h = linear(h, w, b)
        if l != n - 1:
            h = act(h)
    return h


def softplus100(x):
    return F.softplus(x, beta=100)


class OracleNet:
    """Functional EndoSurfNet over the reference's checkpoint layout
    ``{"deform_network","sdf_network","color_network","deviation_network"}``."""

    def __init__(self, ckpt: Dict[str, Dict[str, torch.Tensor]], net_cfg: dict, linear=F.linear):
        self.ckpt = ckpt
        self.cfg = net_cfg
        self.use_deform = bool(net_cfg["use_deform"])
        self.linear = linear
        c = net_cfg
        if self.use_deform:
            self.d_Lx = c["deform_network"]["enc_pos_cfg"]["multires"]
            self.d_Lt = c["deform_network"]["enc_time_cfg"]["multires"]
            self.d_skips = list(c["deform_network"]["skips"])
        self.s_Lx = c["sdf_network"]["enc_pos_cfg"]["multires"]
        self.s_skips = list(c["sdf_network"]["skips"])
        self.c_Lx = c["color_network"]["enc_pos_cfg"]["multires"]
        self.c_Ld = c["color_network"]["enc_dir_cfg"]["multires"]
        self.c_skips = list(c["color_network"]["skips"])

    # -- the three MLPs ------------------------------------------------------------
    def deform_mlp(self, x, t):
        inp = torch.cat([freq_encode(x, self.d_Lx), freq_encode(t, self.d_Lt)], -1)
        return _mlp(_layers(self.ckpt["deform_network"]), inp, self.d_skips, F.relu, self.linear)

    def sdf_mlp(self, x_c):
        return _mlp(_layers(self.ckpt["sdf_network"]), freq_encode(x_c, self.s_Lx), self.s_skips,
                    softplus100, self.linear)

    def color_mlp(self, x_c, n, d, feat):
        inp = torch.cat([freq_encode(x_c, self.c_Lx), n, freq_encode(d, self.c_Ld), feat], -1)
        return torch.sigmoid(_mlp(_layers(self.ckpt["color_network"]), inp, self.c_skips, F.relu,
                                  self.linear))

    def inv_s(self):
        """ref endosurf.py:850-852 + :168."""
        return torch.exp(self.ckpt["deviation_network"]["variance"] * 10.0).clip(1e-6, 1e6)

    # -- composite queries (ref endosurf.py:570-658) -------------------------------
    def canonical(self, x, t):
        return x + self.deform_mlp(x, t) if self.use_deform else x

    def sdf_from_observed(self, x, t):
        return self.sdf_mlp(self.canonical(x, t))[..., :1]

    def sdf_grad_observed(self, x, t):
        with torch.enable_grad():
            x = x.detach().requires_grad_(True) if not x.requires_grad else x
            y = self.sdf_mlp(self.canonical(x, t))[..., :1]
            return torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)[0]

    def sdf_grad_canonical(self, x_c):
        with torch.enable_grad():
            if not x_c.requires_grad:
                x_c = x_c.detach().requires_grad_(True)
            y = self.sdf_mlp(x_c)[..., :1]
            return torch.autograd.grad(y, x_c, torch.ones_like(y), create_graph=True, retain_graph=True)[0]

    def deform_jacobian(self, x, t):
        with torch.enable_grad():
            if not x.requires_grad:
                x = x.detach().requires_grad_(True)
            x_c = self.canonical(x, t)
            rows = []
            for i in range(3):
                y = x_c[:, i]
                rows.append(torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True,
                                                retain_graph=True)[0].unsqueeze(1))
            return torch.cat(rows, 1)  # [n, out, in]

    def forward(self, inputs):
        """ref endosurf.py:660-689 -> [sdf, rgb]."""
        x, d, t = inputs[..., :3], inputs[..., 3:6], inputs[..., 6:]
        with torch.enable_grad():
            if not x.requires_grad:
                x = x.detach().requires_grad_(True)
            x_c = self.canonical(x, t)
            h = self.sdf_mlp(x_c)
            sdf, feat = h[..., :1], h[..., 1:]
            g_c = self.sdf_grad_canonical(x_c)
            jac = self.deform_jacobian(x, t)
            d_c = torch.bmm(jac, d.unsqueeze(-1)).squeeze(-1)
            d_c = d_c / (torch.linalg.norm(d_c, ord=2, dim=-1, keepdim=True) + 1e-10)
            rgb = self.color_mlp(x_c, g_c, d_c, feat)
        return torch.cat([sdf, rgb], -1)

    def forward_parts(self, inputs):
        """Same as :meth:`forward` but returns every intermediate (stage goldens)."""
        x, d, t = inputs[..., :3], inputs[..., 3:6], inputs[..., 6:]
        with torch.enable_grad():
            x = x.detach().requires_grad_(True)
            x_c = self.canonical(x, t)
            h = self.sdf_mlp(x_c)
            sdf, feat = h[..., :1], h[..., 1:]
            g_c = self.sdf_grad_canonical(x_c)
            jac = self.deform_jacobian(x, t)
            d_c = torch.bmm(jac, d.unsqueeze(-1)).squeeze(-1)
            d_c = d_c / (torch.linalg.norm(d_c, ord=2, dim=-1, keepdim=True) + 1e-10)
            rgb = self.color_mlp(x_c, g_c, d_c, feat)
            g_o = self.sdf_grad_observed(x, t)
        return {k: v.detach() for k, v in dict(x_c=x_c, sdf=sdf, feat=feat, g_c=g_c, jac=jac, d_c=d_c,
                                                rgb=rgb, g_o=g_o).items()}


def _split_apply(fn, inputs, chunk):
    """ref utils.py:113-127 (tensor branch)."""
    return torch.cat([fn(part) for part in torch.split(inputs, chunk, 0)], 0)


# --------------------------------------------------------------------------- sampling
def sphere_intersection(rays_o, rays_d, r: float = 1.0):
    """ref utils.py:194-210."""
    d1 = -torch.sum(rays_d * rays_o, dim=-1) / torch.sum(rays_d * rays_d, dim=-1)
    p = rays_o + d1.unsqueeze(-1) * rays_d
    tmp = r * r - torch.sum(p * p, dim=-1)
    mask = tmp > 0.0
    d2 = torch.sqrt(torch.clamp(tmp, min=0.0)) / torch.norm(rays_d, dim=-1)
    near = torch.clamp(d1 - d2, min=0.0)
    far = d1 + d2
    return near[..., None], far[..., None], mask[..., None]


def sample_pdf_det(bins, weights, n_samples):
    """Deterministic inverse-CDF sampling; ref utils.py:160-191 with det=True."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.linspace(0.0 + 0.5 / n_samples, 1.0 - 0.5 / n_samples, steps=n_samples).to(cdf.device)
    u = u.expand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b)


def _rays_d_z(rays_d):
    return rays_d / (rays_d[..., 2:] + 1e-6)


def up_sample(rays_o, rays_d, z_vals, sdf, n_importance, inv_s):
    """ref endosurf.py:221-266."""
    n_rays, n_samples = z_vals.shape
    pts = rays_o[:, None, :] + _rays_d_z(rays_d)[:, None, :] * z_vals[..., :, None]
    radius = torch.linalg.norm(pts, ord=2, dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    sdf = sdf.reshape(n_rays, n_samples)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z_vals[:, :-1], z_vals[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * 0.5
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + 1e-6)
    prev_cos = torch.cat([torch.zeros([n_rays, 1], dtype=z_vals.dtype, device=z_vals.device), cos_val[:, :-1]], -1)
    cos_val = torch.minimum(prev_cos, cos_val)
    cos_val = cos_val.clip(-1e3, 0.0) * inside
    dist = next_z - prev_z
    prev_cdf = torch.sigmoid((mid_sdf - cos_val * dist * 0.5) * inv_s)
    next_cdf = torch.sigmoid((mid_sdf + cos_val * dist * 0.5) * inv_s)
    alpha = (prev_cdf - next_cdf + 1e-6) / (prev_cdf + 1e-6)
    ones = torch.ones([n_rays, 1], dtype=z_vals.dtype, device=z_vals.device)
    weights = alpha * torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return sample_pdf_det(z_vals, weights, n_importance).detach()


def cat_z_vals(net: OracleNet, rays_o, rays_d, time, z_vals, new_z_vals, sdf, last=False):
    """ref endosurf.py:268-287."""
    n_rays, n_samples = z_vals.shape
    n_imp = new_z_vals.shape[1]
    z_all, index = torch.sort(torch.cat([z_vals, new_z_vals], -1), dim=-1)
    if not last:
        pts = rays_o[:, None, :] + _rays_d_z(rays_d)[:, None, :] * new_z_vals[..., :, None]
        t = time[:, None, None].expand(n_rays, n_imp, 1)
        new_sdf = net.sdf_from_observed(pts.reshape(-1, 3), t.reshape(-1, 1)).reshape(n_rays, n_imp)
        sdf = torch.gather(torch.cat([sdf, new_sdf], -1), 1, index)
    return z_all, sdf


def cos_anneal_ratio(iter_step, anneal_end):
    """ref endosurf.py:215-219."""
    if anneal_end == 0.0:
        return 1.0
    return float(np.min([1.0, iter_step / anneal_end]))


# --------------------------------------------------------------------------- rendering
def render_core(net: OracleNet, rays_o, rays_d, time, z_vals, sample_dist, cos_ratio=0.0, net_chunk=80000):
    """ref endosurf.py:134-213."""
    n_rays, n_samples = z_vals.shape
    dev, dt = z_vals.device, z_vals.dtype
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], sample_dist)], -1)
    mid_z = z_vals + dists * 0.5
    pts = (rays_o[:, None, :] + _rays_d_z(rays_d)[:, None, :] * mid_z[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(n_rays, n_samples, 3).reshape(-1, 3)
    t = time[:, None, None].expand(n_rays, n_samples, 1).reshape(-1, 1)

    raw = _split_apply(net.forward, torch.cat([pts, dirs, t], -1), net_chunk)
    sdf = raw[..., :1]
    sampled_color = raw[..., 1:4].reshape(n_rays, n_samples, 3)
    g_o = _split_apply(lambda x: net.sdf_grad_observed(x[..., :3], x[..., 3:4]), torch.cat([pts, t], -1), net_chunk)

    inv_s = net.inv_s().reshape(1, 1).expand(n_rays * n_samples, 1)
    true_cos = (dirs * g_o).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_ratio) + F.relu(-true_cos) * cos_ratio)
    est_next = sdf + iter_cos * dists.reshape(-1, 1) * 0.5
    est_prev = sdf - iter_cos * dists.reshape(-1, 1) * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    p, c = prev_cdf - next_cdf, prev_cdf
    alpha = ((p + 1e-6) / (c + 1e-6)).reshape(n_rays, n_samples).clip(0.0, 1.0)

    pts_norm = torch.linalg.norm(pts, ord=2, dim=-1, keepdim=True).reshape(n_rays, n_samples)
    relax = (pts_norm < 1.2).to(dt).detach()
    ones = torch.ones([n_rays, 1], dtype=dt, device=dev)
    weights = alpha * torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    depth_map = torch.sum(weights * mid_z, -1, keepdim=True)
    color = (sampled_color * weights[:, :, None]).sum(dim=1)
    g_err = (torch.linalg.norm(g_o.reshape(n_rays, n_samples, 3), ord=2, dim=-1) - 1.0) ** 2
    g_err = (relax * g_err).sum() / (relax.sum() + 1e-6)
    return {
        "color_map": color,
        "depth_map": depth_map,
        "gradients_o": g_o.reshape(n_rays, n_samples, 3),
        "gradient_o_error": g_err,
        "cdf": c.reshape(n_rays, n_samples),
        "weights": weights,
        "s_val": 1.0 / inv_s,
        "sdf": sdf.reshape(n_rays, n_samples),          # extra (stage checks)
        "sampled_color": sampled_color,                 # extra (stage checks)
        "mid_z_vals": mid_z,                            # extra (stage checks)
    }


def coarse_z_vals(rays, n_samples, perturb=False, t_rand=None):
    """ref endosurf.py:63-82.  ``t_rand`` lets a test inject the one-per-ray jitter."""
    rays_o, rays_d = rays[..., :3], rays[..., 3:6]
    near, far, _ = sphere_intersection(rays_o, rays_d)
    t_vals = torch.linspace(0.0, 1.0, n_samples, device=rays.device)
    z_vals = near + (far - near) * t_vals[None, :]
    if perturb:
        if t_rand is None:
            t_rand = torch.rand([rays.shape[0], 1], device=rays.device) - 0.5
        z_vals = z_vals + t_rand * (2.0 / n_samples)
    return z_vals


def hierarchical_z_vals(net: OracleNet, rays, z_vals, n_importance, up_sample_steps, return_trace=False):
    """ref endosurf.py:85-110 (the no-grad up-sampling loop)."""
    rays_o, rays_d, time = rays[..., :3], rays[..., 3:6], rays[..., 8]
    n_rays, n_samples = z_vals.shape
    trace = []
    with torch.no_grad():
        pts = rays_o[:, None, :] + _rays_d_z(rays_d)[:, None, :] * z_vals[..., :, None]
        t = time[..., None, None].expand(n_rays, n_samples, 1)
        sdf = net.sdf_from_observed(pts.reshape(-1, 3), t.reshape(-1, 1)).reshape(n_rays, n_samples)
        for i in range(up_sample_steps):
            if return_trace:
                trace.append((z_vals.clone(), sdf.clone()))
            new_z = up_sample(rays_o, rays_d, z_vals, sdf, n_importance // up_sample_steps, 64 * 2 ** i)
            z_vals, sdf = cat_z_vals(net, rays_o, rays_d, time, z_vals, new_z, sdf,
                                     last=(i + 1 == up_sample_steps))
            if return_trace:
                trace[-1] = trace[-1] + (new_z.clone(),)
    return (z_vals, trace) if return_trace else z_vals


def render_rays(net: OracleNet, render_cfg: dict, rays, iter_step=0, perturb_overwrite=None, t_rand=None,
                z_vals_override=None):
    """ref endosurf.py:60-132.  Returns the reference's 8-key dict (+ ``z_vals``)."""
    n_rays = rays.shape[0]
    rays_o, rays_d, time = rays[..., :3], rays[..., 3:6], rays[..., 8]
    n_samples = render_cfg["n_samples"]
    n_importance = render_cfg["n_importance"]
    perturb = render_cfg["perturb"] if perturb_overwrite is None else perturb_overwrite
    sample_dist = 2.0 / n_samples
    if z_vals_override is not None:
        z_vals = z_vals_override
    else:
        z_vals = coarse_z_vals(rays, n_samples, perturb, t_rand)
        if iter_step >= render_cfg["important_begin_iter"] and n_importance > 0:
            z_vals = hierarchical_z_vals(net, rays, z_vals, n_importance, render_cfg["up_sample_steps"])
    fine = render_core(net, rays_o, rays_d, time, z_vals, sample_dist,
                       cos_ratio=cos_anneal_ratio(iter_step, render_cfg["anneal_end"]),
                       net_chunk=render_cfg["net_chunk"])
    m = z_vals.shape[1]
    return {
        "color_map": fine["color_map"],
        "depth_map": fine["depth_map"],
        "gradients_o": fine["gradients_o"],
        "gradient_o_error": fine["gradient_o_error"],
        "weights": fine["weights"],
        "weight_max": torch.max(fine["weights"], dim=-1, keepdim=True)[0],
        "cdf": fine["cdf"],
        "s_val": fine["s_val"].reshape(n_rays, m).mean(dim=-1, keepdim=True),
        "z_vals": z_vals,
        "sdf": fine["sdf"],
        "sampled_color": fine["sampled_color"],
    }


# --------------------------------------------------------------------------- helpers (section 8f "next" rows)
def errorondepth(net: OracleNet, rays, d_gt, mask):
    """ref endosurf.py:289-317."""
    rays_o, rays_d, time = rays[..., :3], rays[..., 3:6], rays[..., 8]
    pts = (rays_o + _rays_d_z(rays_d) * d_gt).reshape(-1, 3)
    t = time.reshape(-1, 1)
    sdf = net.sdf_from_observed(pts, t)
    g_o = net.sdf_grad_observed(pts, t)
    relu_cos = F.relu((rays_d * g_o).sum(-1, keepdim=True))
    inside = (torch.linalg.norm(pts.detach(), ord=2, dim=-1, keepdim=True) < 1.0).to(rays.dtype) * mask
    denom = inside.sum() + 1e-6
    return (inside * sdf).abs().sum() / denom, relu_cos.abs().sum() / denom, inside


def secant(net: OracleNet, f_low, f_high, d_low, d_high, n_steps, rays, tau=0.0):
    """ref endosurf.py:422-449 (note: d/d.z without epsilon, :427)."""
    rays_o, rays_d, time = rays[..., :3], rays[..., 3:6], rays[..., 8:]
    d_z = rays_d / rays_d[..., 2:]
    d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
    for _ in range(n_steps):
        p_mid = rays_o + d_pred.unsqueeze(-1) * d_z
        with torch.no_grad():
            f_mid = net.sdf_from_observed(p_mid, time)[..., 0] - tau
        low = f_mid < 0
        d_low = torch.where(low, d_pred, d_low)
        f_low = torch.where(low, f_mid, f_low)
        d_high = torch.where(low, d_high, d_pred)
        f_high = torch.where(low, f_high, f_mid)
        d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
    return d_pred


def ray_marching(net: OracleNet, rays, tau=0.0, n_steps=128, n_secant_steps=8):
    """ref endosurf.py:344-420 with the (always 128) randint draw fixed."""
    n_rays = rays.shape[0]
    rays_o, rays_d, time = rays[..., :3], rays[..., 3:6], rays[..., 8:]
    near, far, _ = sphere_intersection(rays_o, rays_d)
    t_vals = torch.linspace(0.0, 1.0, steps=n_steps, device=rays.device)
    d_prop = near * (1.0 - t_vals) + far * t_vals
    pts = rays_o[:, None, :] + d_prop[..., None] * _rays_d_z(rays_d)[:, None, :]
    t = time[:, None, :].expand(n_rays, n_steps, 1)
    with torch.no_grad():
        val = net.sdf_from_observed(pts.reshape(-1, 3), t.reshape(-1, 1)).view(n_rays, n_steps) - tau
    val = -val
    mask0 = val[:, 0] < 0
    sign = torch.cat([torch.sign(val[:, :-1] * val[:, 1:]), torch.ones(n_rays, 1, device=rays.device)], -1)
    cost = sign * torch.arange(n_steps, 0, -1, device=rays.device).float()
    values, idx = torch.min(cost, -1)
    ar = torch.arange(n_rays, device=rays.device)
    mask = (values < 0) & (val[ar, idx] < 0) & mask0
    idx_hi = torch.clamp(idx + 1, max=n_steps - 1)
    out = torch.ones(n_rays, device=rays.device)
    if mask.any():
        d_pred = secant(net, val[ar, idx][mask], val[ar, idx_hi][mask], d_prop[ar, idx][mask],
                        d_prop[ar, idx_hi][mask], n_secant_steps, rays[mask], tau)
        out[mask] = d_pred
    out[~mask] = float("inf")
    out[~mask0] = 0
    return out.unsqueeze(-1)


# --------------------------------------------------------------------------- synthetic inputs
def synthetic_rays(n_rays: int, frame: int = 0, n_frames: int = 60, hw: int = 512, seed: int = 0,
                   device="cpu") -> torch.Tensor:
    """Pinhole rays from o=(0,0,-1.5) over an hw x hw image, focal 1.2*hw, unit directions,
    cols 6-7 zero, time = frame/(n_frames-1).  Pixels are drawn with a seeded generator."""
    g = torch.Generator().manual_seed(seed + 7919 * frame)
    pix = torch.randint(0, hw * hw, (n_rays,), generator=g)
    u = (pix % hw).float() + 0.5
    v = (pix // hw).float() + 0.5
    f = 1.2 * hw
    d = torch.stack([(u - hw / 2) / f, (v - hw / 2) / f, torch.ones_like(u)], -1)
    d = d / d.norm(dim=-1, keepdim=True)
    o = torch.tensor([0.0, 0.0, -1.5]).expand(n_rays, 3)
    rays = torch.cat([o, d, torch.zeros(n_rays, 2), torch.full((n_rays, 1), frame / max(n_frames - 1, 1))], -1)
    return rays.to(device)
