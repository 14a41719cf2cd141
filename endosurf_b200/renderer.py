"""Host-side mirror of the reference's renderer interface on top of the sm_100a kernels.

``EndoSurfRenderer(render_cfg, net_cfg, device)`` has the member surface the reference trainer uses
(reference ``src/trainer/trainer_endosurf.py:51,65-68,81,88,130,140,155,230,336,427,453``; class at
``src/renderer/endosurf.py:14-521``): same constructor, ``forward``/``render_rays`` returning the same 8-key
dict, ``n_samples``/``n_importance``, ``get_train_params``, ``save_checkpoint``/``load_checkpoint`` with the
reference's state-dict keys (``net.{l}.bias|weight_g|weight_v``, ``variance``).

Parameters live in ordinary ``nn.Parameter`` objects; the only PyTorch arithmetic on the hot path is folding
weight norm (``W = g v/||v||``, 1.65 M elements).  Everything per-sample -- encodings, the three MLPs, normals,
the deformation Jacobian, hierarchical sampling and compositing -- runs in ``libendosurf_b200.so`` through the
C ABI (``include/endosurf_b200.h``).  There is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import weakref
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib


# ------------------------------------------------------------------------------------------------ parameters
class WNLinear(nn.Module):
    """Parameters of one weight-normalised linear layer in the reference's (old-API) layout:
    ``bias``, ``weight_g`` [out,1], ``weight_v`` [out,in] -- registered in that order so that
    ``parameters()`` / Adam state indices line up with the reference (utils.py:57-58)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor):
        super().__init__()
        self.bias = nn.Parameter(bias.clone())
        self.weight_g = nn.Parameter(weight.norm(dim=1, keepdim=True).clone())
        self.weight_v = nn.Parameter(weight.clone())

    def effective_weight(self) -> torch.Tensor:
        return self.weight_v * (self.weight_g / torch.linalg.norm(self.weight_v, dim=1, keepdim=True))


def _default_linear_init(dim_in: int, dim_out: int):
    lin = nn.Linear(dim_in, dim_out)  # kaiming-uniform(a=sqrt(5)) weight, uniform bias: torch's default
    return lin.weight.detach(), lin.bias.detach()


def _make_mlp(n_layers, hidden, in_dim, out_dim, skips, style, geometric_bias=None):
    """Layer shapes and initialisation of the reference's builders.

    style "nerf": skip layer takes hidden+in_dim inputs (utils.py:11-60)
    style "idr" : the layer before a skip emits hidden-in_dim outputs (utils.py:63-111)
    geometric_bias (float) switches on the SDF sphere initialisation (utils.py:38-56)."""
    layers = []
    for l in range(n_layers):
        if style == "nerf":
            d0 = in_dim if l == 0 else (hidden + in_dim if l in skips else hidden)
            d1 = out_dim if l == n_layers - 1 else hidden
        else:
            d0 = in_dim if l == 0 else hidden
            d1 = out_dim if l == n_layers - 1 else (hidden - in_dim if (l + 1) in skips else hidden)
        w, b = _default_linear_init(d0, d1)
        if geometric_bias is not None:
            if l == n_layers - 1:
                w = torch.empty(d1, d0).normal_(mean=math.sqrt(math.pi) / math.sqrt(d0), std=1e-4)
                b = torch.full((d1,), -float(geometric_bias))
            elif l == 0:
                b = torch.zeros(d1)
                w = torch.zeros(d1, d0)
                w[:, :3].normal_(0.0, math.sqrt(2) / math.sqrt(d1))
            elif l in skips:
                b = torch.zeros(d1)
                w = torch.empty(d1, d0).normal_(0.0, math.sqrt(2) / math.sqrt(d1))
                w[:, -(in_dim - 3):] = 0.0
            else:
                b = torch.zeros(d1)
                w = torch.empty(d1, d0).normal_(0.0, math.sqrt(2) / math.sqrt(d1))
        layers.append(WNLinear(w, b))
    return nn.ModuleList(layers)


def _enc_dim(cfg) -> int:
    if cfg["enc_type"] != "frequency":
        raise NotImplementedError("endosurf_b200 kernels implement the frequency encoder only")
    return cfg["input_dim"] * (1 + 2 * cfg["multires"])


class _Net(nn.Module):
    def __init__(self, net: nn.ModuleList):
        super().__init__()
        self.net = net


class _Variance(nn.Module):
    def __init__(self, init_val):
        super().__init__()
        self.variance = nn.Parameter(torch.tensor(float(init_val)))


class EndoSurfNet(nn.Module):
    """Parameter container with the reference's module/parameter names (endosurf.py:524-568)."""

    def __init__(self, net_cfg: dict):
        super().__init__()
        self.bound = net_cfg["bound"]
        self.use_deform = bool(net_cfg["use_deform"])
        if self.use_deform:
            c = net_cfg["deform_network"]
            in_dim = _enc_dim(c["enc_pos_cfg"]) + _enc_dim(c["enc_time_cfg"])
            self.deform_network = _Net(_make_mlp(c["n_layers"], c["hidden_dim"], in_dim, c["out_dim"],
                                                 list(c["skips"]), "idr"))
        c = net_cfg["sdf_network"]
        self.sdf_network = _Net(_make_mlp(c["n_layers"], c["hidden_dim"], _enc_dim(c["enc_pos_cfg"]), c["out_dim"],
                                          list(c["skips"]), "nerf",
                                          geometric_bias=c.get("geometric_init_bias", 0.8)
                                          if c.get("geometric_init", True) else None))
        c = net_cfg["color_network"]
        in_dim = _enc_dim(c["enc_pos_cfg"]) + 3 + _enc_dim(c["enc_dir_cfg"]) + c["feat_dim"]
        self.color_network = _Net(_make_mlp(c["n_layers"], c["hidden_dim"], in_dim, c["out_dim"], list(c["skips"]),
                                            "nerf"))
        self.deviation_network = _Variance(net_cfg["deviation_network"]["init_val"])

    def get_train_params(self) -> Dict[str, List[nn.Parameter]]:
        p = {}
        if self.use_deform:
            p["deform_network"] = list(self.deform_network.parameters())
        p["sdf_network"] = list(self.sdf_network.parameters())
        p["color_network"] = list(self.color_network.parameters())
        p["deviation_network"] = list(self.deviation_network.parameters())
        return p

    def save_checkpoint(self):
        ck = {}
        if self.use_deform:
            ck["deform_network"] = self.deform_network.state_dict()
        ck["sdf_network"] = self.sdf_network.state_dict()
        ck["color_network"] = self.color_network.state_dict()
        ck["deviation_network"] = self.deviation_network.state_dict()
        return ck

    def load_checkpoints(self, ckpt):
        if self.use_deform:
            self.deform_network.load_state_dict(ckpt["deform_network"])
        self.sdf_network.load_state_dict(ckpt["sdf_network"])
        self.color_network.load_state_dict(ckpt["color_network"])
        self.deviation_network.load_state_dict(ckpt["deviation_network"])

    # ---- the reference's query methods (endosurf.py:570-689).  The parameters live here, the kernels in the renderer
    # that owns this module; every query is one fused chain launch (differentiable when grad is enabled).
    def _owner(self):
        r = self.__dict__.get("_renderer_ref")
        r = r() if r is not None else None
        if r is None:
            raise RuntimeError("EndoSurfNet queries need the EndoSurfRenderer that owns this module")
        return r

    def _fields(self, x, d, t):
        r = self._owner()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            x = x.reshape(-1, 3)
            if d is None:  # geometry-only query: the colour chain still runs on a dummy view direction
                d = torch.zeros_like(x)
                d[:, 2] = 1.0
            return r.point_field(x, d.reshape(-1, 3), t.reshape(-1, 1))
        o = r.point_forward(x, d, t)
        return o["sdf"], o["g_c"], o["jac"], o.get("rgb")

    def get_sdf_from_observed_space(self, x, t):
        """endosurf.py:570-579 -> [n,1]"""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._fields(x, None, t)[0]
        return self._owner().sdf_from_observed_space(x, t)

    def get_sdf_grad_from_observed_space(self, x, t):
        """endosurf.py:581-601 -> [n,3]: d sdf / d x = J^T g_c"""
        _, g_c, jac, _ = self._fields(x, None, t)
        return (jac * g_c[:, :, None]).sum(1)

    def get_deform_grad_from_observed_space(self, x, t):
        """endosurf.py:621-658 -> [n,3,3]: J[i,j] = d x_c_i / d x_j"""
        return self._fields(x, None, t)[2]

    def forward(self, inputs):
        """endosurf.py:660-689: inputs [n,7] = (x, d, t) -> cat[sdf, rgb] [n,4]"""
        sdf, _, _, rgb = self._fields(inputs[:, :3], inputs[:, 3:6], inputs[:, 6:7])
        return torch.cat([sdf, rgb], dim=-1)


def _net_config_struct(net_cfg: dict, precision_terms: int) -> _lib.EsNetConfig:
    s, c = net_cfg["sdf_network"], net_cfg["color_network"]
    d = net_cfg.get("deform_network", None) if net_cfg["use_deform"] else None
    nets = [n for n in (d, s, c) if n is not None]
    n_layers = {n["n_layers"] for n in nets}
    hidden = {n["hidden_dim"] for n in nets}
    skips = {tuple(n["skips"]) for n in nets}
    if len(n_layers) != 1 or len(hidden) != 1 or len(skips) != 1 or len(next(iter(skips))) > 1:
        raise NotImplementedError("endosurf_b200 kernels need the three networks to share n_layers/hidden_dim and "
                                  "a single skip layer (every shipped reference config does)")
    if s["out_dim"] != 257 or c["feat_dim"] != 256 or c["out_dim"] != 3 or (d is not None and d["out_dim"] != 3):
        raise NotImplementedError("unsupported output widths")
    sk = next(iter(skips))
    return _lib.EsNetConfig(
        use_deform=int(bool(net_cfg["use_deform"])), n_layers=n_layers.pop(), skip_layer=sk[0] if sk else -1,
        hidden_dim=hidden.pop(),
        multires_deform_pos=d["enc_pos_cfg"]["multires"] if d else 6,
        multires_deform_time=d["enc_time_cfg"]["multires"] if d else 6,
        multires_sdf_pos=s["enc_pos_cfg"]["multires"],
        multires_color_pos=c["enc_pos_cfg"]["multires"], multires_color_dir=c["enc_dir_cfg"]["multires"],
        precision_terms=precision_terms)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# CTA pairs are the default; ES_PAIR=0 selects the single-CTA kernels (A/B comparisons, debugging)
PAIR_MODE_DEFAULT = os.environ.get("ES_PAIR", "0") != "0"


class EndoSurfRenderer(nn.Module):
    """Drop-in for the reference ``EndoSurfRenderer`` (endosurf.py:14-132) backed by the CUDA library."""

    def __init__(self, render_cfg, net_cfg, device="cuda", precision_terms: int = 3):
        super().__init__()
        self.render_cfg = render_cfg
        self.net_cfg = net_cfg
        self.device = device
        self.dtype = torch.get_default_dtype()
        if self.dtype != torch.float32:
            raise NotImplementedError("endosurf_b200 computes in fp32 storage (fp16 hi/lo split tensor-core products)")
        self.model = EndoSurfNet(net_cfg).to(device)
        self.model.__dict__["_renderer_ref"] = weakref.ref(self)  # (not a sub-module: the queries run in this object)
        self.anneal_end = render_cfg["anneal_end"]
        self.n_samples = render_cfg["n_samples"]
        self.perturb = render_cfg["perturb"]
        self.n_importance = render_cfg["n_importance"]
        self.important_begin_iter = render_cfg["important_begin_iter"]
        self.up_sample_steps = render_cfg["up_sample_steps"]
        self.net_chunk = render_cfg["net_chunk"]  # accepted for compatibility; the fused kernels need no chunking
        self._cfg_struct = _net_config_struct(net_cfg, precision_terms)
        self._ctx = None
        self._packed_version = None
        self._consts = {}

    # ------------------------------------------------------------------ reference surface
    def get_train_params(self):
        return self.model.get_train_params()

    def load_checkpoint(self, ckpt):
        self.model.load_checkpoints(ckpt)
        self._packed_version = None

    def save_checkpoint(self):
        return self.model.save_checkpoint()

    def get_cos_anneal_ratio(self, iter_step):
        """endosurf.py:215-219"""
        if self.anneal_end == 0.0:
            return 1.0
        return float(np.min([1.0, iter_step / self.anneal_end]))

    def forward(self, rays, **kwargs):
        return self.render_rays(rays, **kwargs)

    # ------------------------------------------------------------------ library plumbing
    def _context(self):
        if self._ctx is None:
            if not torch.cuda.is_available():
                raise RuntimeError("endosurf_b200 needs a CUDA sm_100 device; there is no CPU fallback")
            lib = _lib.load()
            dev = torch.device(self.device)
            torch.cuda.set_device(dev if dev.index is not None else torch.cuda.current_device())
            ctx = C.c_void_p()
            rc = lib.es_create(C.byref(ctx), C.byref(self._cfg_struct))
            if rc != 0:
                raise _lib.EsError(f"es_create failed: {_lib.ES_E.get(rc, rc)}")
            self._ctx = ctx
            self.set_pair_mode(PAIR_MODE_DEFAULT)
            if os.environ.get("ES_STORE_HINT"):  # A/B experiments: L2 evict_first hint on the plane-record stores
                lib.es_debug_set(ctx, 8, int(os.environ["ES_STORE_HINT"]))
            if os.environ.get("ES_FEAT_REC"):  # A/B experiments: 0 = fp32 feature rows between the chains
                lib.es_debug_set(ctx, 7, int(os.environ["ES_FEAT_REC"]))
            if os.environ.get("ES_DEBUG_FLAGS"):  # ablation experiments; only -DES_ABLATE builds of the library look at it
                lib.es_debug_set(ctx, 0, int(os.environ["ES_DEBUG_FLAGS"]))
        return self._ctx

    def set_pair_mode(self, on: bool):
        """Run the 256-wide chains on CTA pairs (tcgen05 cta_group::2: two SMs share every weight unit) or on single
        CTAs.  The packed weights change layout, so they are handed to the library again on the next call."""
        lib, ctx = _lib.load(), self._context()
        _lib.check(ctx, lib.es_debug_set(ctx, 4, int(bool(on))), "es_debug_set(pair)")
        self._packed_version = None

    def __del__(self):
        try:
            if getattr(self, "_ctx", None) is not None:
                _lib.load().es_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _fast_params(self):
        """Every parameter in ``self.model.parameters()`` order (bias, weight_g, weight_v per layer per network, then
        ``variance``) read straight from the live ``_parameters`` dicts: nn.Module's recursive traversal costs about
        270 us for the 82 parameters, and every library call starts with a version check (13 per trainer step)."""
        mods = self.model._modules
        out = []
        for name in ("deform_network", "sdf_network", "color_network"):
            net = mods.get(name)
            if net is not None:
                for layer in net._modules["net"]._modules.values():
                    out.extend(layer._parameters.values())
        out.extend(mods["deviation_network"]._parameters.values())
        return out

    def _params_version(self):
        ps = self._fast_params()
        return tuple([p._version for p in ps] + [p.data_ptr() for p in ps])

    def _sync_weights(self):
        """Hand the parameters to the library when any of them changed: the weight-norm fold W = g v/|v| (reference
        utils.py:57-58) and the fp16 hi/lo operand packing both run in the library (es_load_network_wn)."""
        ver = self._params_version()
        if ver == self._packed_version:
            return
        lib, ctx = _lib.load(), self._context()
        nets = []
        if self.model.use_deform:
            nets.append((0, self.model.deform_network))
        nets += [(1, self.model.sdf_network), (2, self.model.color_network)]
        for net_id, mod in nets:
            for l in mod.net:
                for p in (l.weight_v, l.weight_g, l.bias):
                    if not p.is_contiguous() or p.dtype != torch.float32:
                        raise ValueError("endosurf_b200 parameters must be contiguous fp32 tensors")
            n = len(mod.net)
            vp = (C.c_void_p * n)(*[l.weight_v.data_ptr() for l in mod.net])
            gp = (C.c_void_p * n)(*[l.weight_g.data_ptr() for l in mod.net])
            bp = (C.c_void_p * n)(*[l.bias.data_ptr() for l in mod.net])
            _lib.check(ctx, lib.es_load_network_wn(ctx, net_id, vp, gp, bp, self._stream()), "es_load_network_wn")
        self._packed_version = ver

    def _poll_device_error(self):
        """Non-blocking check of the device-side watchdog word: raises if an EARLIER launch of this context recorded
        a barrier time-out (results after that are garbage; training must not continue silently)."""
        lib, ctx = _lib.load(), self._context()
        _lib.check(ctx, lib.es_poll_error(ctx, self._stream()), "device-side watchdog")

    def _const(self, key, fn):
        if key not in self._consts:
            self._consts[key] = fn()
        return self._consts[key]

    # ------------------------------------------------------------------ differentiable (training) path
    train_ray_chunk = 16384  # rays per library call: bounds the plane records (about 4.4 MB per ray at 64+64 samples)

    def point_field(self, x, d, t):
        """Differentiable EndoSurfNet.forward + gradient queries on explicit points:
        (sdf [n,1], g_c [n,3], jac [n,3,3], rgb [n,3]); gradients flow to every network parameter."""
        from .training import PointFieldFn, param_hub
        params, hub = param_hub(self)
        return PointFieldFn.apply(self, x, d, t, hub, params)

    def _sample_z(self, rays, iter_step, perturb_overwrite):
        """Coarse + hierarchical sampling only (no grad, CUDA): z_vals [R,M]."""
        self._sync_weights()
        lib, ctx = _lib.load(), self._context()
        dev, R = rays.device, rays.shape[0]
        ns, ni = int(self.n_samples), int(self.n_importance)
        perturb = self.perturb if perturb_overwrite is None else perturb_overwrite
        do_up = bool(iter_step >= self.important_begin_iter and ni > 0)
        M = ns + (ni if do_up else 0)
        steps = int(self.up_sample_steps)
        t_vals = self._const(("t", ns, dev), lambda: torch.linspace(0.0, 1.0, ns, device=dev))
        u_vals = None
        if do_up:
            k = ni // steps
            u_vals = self._const(("u", k, dev),
                                 lambda: torch.linspace(0.0 + 0.5 / k, 1.0 - 0.5 / k, steps=k, device=dev))
        t_rand = (torch.rand([R, 1], device=dev) - 0.5).reshape(-1).contiguous() if perturb else None
        z = torch.empty(R, M, device=dev)
        prm = _lib.EsRenderParams(
            n_samples=ns, n_importance=ni, up_sample_steps=steps, do_upsample=int(do_up), cos_anneal_ratio=0.0,
            variance=self.model.deviation_network.variance.data_ptr(), t_vals=t_vals.data_ptr(),
            u_vals=u_vals.data_ptr() if u_vals is not None else None,
            t_rand=t_rand.data_ptr() if t_rand is not None else None, z_override=None)
        o = _lib.EsRenderOut(z_vals=z.data_ptr())
        _lib.check(ctx, lib.es_render_rays(ctx, _ptr(rays), R, C.byref(prm), C.byref(o), self._stream()),
                   "es_render_rays(sampling)")
        return z

    def _render_rays_train(self, rays, iter_step, perturb_overwrite, z_vals_override=None, return_extras=False,
                           cos_ratio=None):
        """render_rays with autograd (endosurf.py:60-213): sampling runs without grad exactly as in the reference
        (:86); everything after it - points, the three MLPs with normals and Jacobian, compositing - is one fused
        library forward and one library backward (training.RenderFn)."""
        from .training import RenderFn, param_hub
        rays = rays.detach().contiguous().float()
        R = rays.shape[0]
        with torch.no_grad():
            z = z_vals_override.detach().contiguous().float() if z_vals_override is not None else \
                self._sample_z(rays, iter_step, perturb_overwrite)
        cos_ratio = self.get_cos_anneal_ratio(iter_step) if cos_ratio is None else float(cos_ratio)
        params, hub = param_hub(self)
        variance = self.model.deviation_network.variance
        names = ["color_map", "depth_map", "gradients_o", "gradient_o_error", "weights", "cdf", "sdf",
                 "sampled_color", "weight_max", "s_val", "eikonal_den"]
        if R <= self.train_ray_chunk:
            vals = RenderFn.apply(self, rays, z, cos_ratio, variance, hub, params)
            out = dict(zip(names, vals))
        else:
            # very large batches: several library calls; the eikonal mean is re-normalised over the whole batch
            parts = [RenderFn.apply(self, rays[r0:r0 + self.train_ray_chunk], z[r0:r0 + self.train_ray_chunk],
                                    cos_ratio, variance, hub, params) for r0 in range(0, R, self.train_ray_chunk)]
            out = {k: torch.cat([p[i] for p in parts]) for i, k in enumerate(names)
                   if k not in ("gradient_o_error", "eikonal_den")}
            dens = torch.stack([p[10].reshape(()) for p in parts])           # sum(relax) + 1e-6 of every chunk
            total = (dens - 1e-6).sum() + 1e-6
            out["gradient_o_error"] = (torch.stack([p[3] for p in parts]) * dens).sum() / total
            out["eikonal_den"] = total.reshape(1)
        self.last_eikonal_den = out.pop("eikonal_den")
        if not return_extras:
            out.pop("sdf")
            out.pop("sampled_color")
        else:
            out["z_vals"] = z
        return out

    # ------------------------------------------------------------------ helper methods of the reference renderer
    # (SURVEY 8f "next" rows: the callers either side of render_rays; all network work goes through the CUDA chains)
    @staticmethod
    def _split_rays(rays):
        rays_o, rays_d = rays[..., :3], rays[..., 3:6]
        return rays_o, rays_d, rays_d / (rays_d[..., 2:] + 1e-6)

    def _point_sdf_and_normal(self, pts, dirs, t):
        """(sdf [n,1], g_o [n,3]) at explicit points; differentiable when grad is enabled
        (get_sdf_from_observed_space + get_sdf_grad_from_observed_space, endosurf.py:570-601)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()):
            sdf, g_c, jac, _ = self.point_field(pts, dirs, t)
            return sdf, (jac * g_c[:, :, None]).sum(1)
        o = self.point_forward(pts, dirs, t)
        return o["sdf"], o["g_o"]

    def errorondepth(self, rays, d_gt, mask, iter_step=0):
        """endosurf.py:289-317: SDF and normal-angle error at the ground-truth depth of every ray."""
        rays = rays.float()
        rays_o, rays_d, rays_d_z = self._split_rays(rays)
        time = rays[..., 8]
        pts = (rays_o + rays_d_z * d_gt).reshape(-1, 3)
        sdf, g_o = self._point_sdf_and_normal(pts, rays_d.reshape(-1, 3), time.reshape(-1, 1))
        relu_cos = torch.relu((rays_d * g_o).sum(-1, keepdim=True))
        pts_norm = torch.linalg.norm(pts.detach(), ord=2, dim=-1, keepdim=True)
        inside = (pts_norm < 1.0).to(self.dtype) * mask
        denom = inside.sum() + 1e-6
        return (inside * sdf).abs().sum() / denom, relu_cos.abs().sum() / denom, inside

    def secant(self, f_low, f_high, d_low, d_high, n_secant_steps, rays, tau=0.0, max_points=None):
        """endosurf.py:422-449 without host round trips (the data-dependent masks become torch.where)."""
        rays_o, rays_d, time = rays[..., :3], rays[..., 3:6], rays[..., 8:]
        rays_d_z = rays_d / rays_d[..., 2:]  # no epsilon here in the reference (:427)
        d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
        for _ in range(n_secant_steps):
            p_mid = rays_o + d_pred.unsqueeze(-1) * rays_d_z
            with torch.no_grad():
                f_mid = self.sdf_from_observed_space(p_mid, time)[..., 0] - tau
            low = f_mid < 0
            d_low = torch.where(low, d_pred, d_low)
            f_low = torch.where(low, f_mid, f_low)
            d_high = torch.where(low, d_high, d_pred)
            f_high = torch.where(low, f_high, f_mid)
            d_pred = -f_low * (d_high - d_low) / (f_high - f_low) + d_low
        return d_pred

    def ray_marching(self, rays, tau=0.0, n_steps=[128, 129], n_secant_steps=8, max_points=64000):
        """endosurf.py:344-420: first sign change of the SDF along each ray + secant refinement -> d_i [R,1]
        (inf: no surface, 0: the first sample is already inside).  No host synchronisation."""
        n_steps = int(torch.randint(n_steps[0], n_steps[1], (1,)).item())  # always 128, kept for RNG parity (:352)
        rays = rays.float()
        n_rays = rays.shape[0]
        rays_o, rays_d, rays_d_z = self._split_rays(rays)
        time = rays[..., 8:]
        d1 = -torch.sum(rays_d * rays_o, dim=-1) / torch.sum(rays_d * rays_d, dim=-1)
        pmid = rays_o + d1.unsqueeze(-1) * rays_d
        d2 = torch.sqrt(torch.clamp(1.0 - torch.sum(pmid * pmid, dim=-1), min=0.0)) / torch.norm(rays_d, dim=-1)
        near, far = torch.clamp(d1 - d2, min=0.0)[..., None], (d1 + d2)[..., None]  # utils.py:194-210
        t_vals = torch.linspace(0.0, 1.0, steps=n_steps, device=rays.device)
        d_prop = near * (1.0 - t_vals) + far * t_vals
        pts = rays_o[:, None, :] + d_prop[..., None] * rays_d_z[:, None, :]
        with torch.no_grad():
            val = self.sdf_from_observed_space(pts.reshape(-1, 3), time[:, None, :].expand(n_rays, n_steps, 1)
                                               .reshape(-1, 1)).view(n_rays, n_steps) - tau
        val = -val
        mask0 = val[:, 0] < 0
        sign = torch.cat([torch.sign(val[:, :-1] * val[:, 1:]), torch.ones(n_rays, 1, device=rays.device)], -1)
        cost = sign * torch.arange(n_steps, 0, -1, device=rays.device).float()
        values, idx = torch.min(cost, -1)
        ar = torch.arange(n_rays, device=rays.device)
        mask = (values < 0) & (val[ar, idx] < 0) & mask0
        idx_hi = torch.clamp(idx + 1, max=n_steps - 1)
        # rays without a crossing run the secant on a harmless bracket and are overwritten below
        f_low = torch.where(mask, val[ar, idx], torch.full_like(values, -1.0))
        f_high = torch.where(mask, val[ar, idx_hi], torch.full_like(values, 1.0))
        d_pred = self.secant(f_low, f_high, d_prop[ar, idx], d_prop[ar, idx_hi], n_secant_steps, rays, tau)
        out = torch.where(mask, d_pred, torch.full_like(d_pred, float("inf")))
        out = torch.where(mask0, out, torch.zeros_like(out))
        return out.unsqueeze(-1)

    def surface_neighbour_error(self, rays, mask, iter_step=0, neighbour_rad=0.05):
        """endosurf.py:319-342: normal consistency between the ray-marched surface point and a random neighbour.
        Masked arithmetic instead of boolean indexing, so there is no host sync and always a tensor result."""
        rays = rays.float()
        rays_o, rays_d, rays_d_z = self._split_rays(rays)
        time = rays[..., 8]
        with torch.no_grad():
            d_i = self.ray_marching(rays, max_points=self.net_chunk)
        valid = ((d_i.abs() != np.inf) & (d_i != 0) & (mask == 1))[..., 0]
        d_safe = torch.where(valid[:, None], d_i, torch.ones_like(d_i))
        p_surf = rays_o + d_safe * rays_d_z
        p_neig = p_surf + (torch.rand_like(p_surf) - 0.5) * neighbour_rad
        pp = torch.cat([p_surf, p_neig], 0)
        tt = torch.cat([time, time], 0).unsqueeze(-1)
        _, g = self._point_sdf_and_normal(pp, torch.cat([rays_d, rays_d], 0), tt)
        normal = g / (torch.linalg.norm(g, ord=2, dim=-1, keepdim=True) + 1e-10)
        n = rays.shape[0]
        diff = torch.abs(normal[:n] - normal[n:]) * valid[:, None].to(normal.dtype)
        return diff.sum() / (3.0 * valid.sum().clamp_min(1))

    def renderonpts(self, pts, dirs, ts, net_chunk=80000, cpu=True):
        """endosurf.py:502-521: colour and unit normal at explicit points (surface rendering of mesh vertices)."""
        sh = list(pts.shape[:-1])
        pts = pts.reshape(-1, 3).float()
        dirs = dirs.reshape(-1, 3).float()
        if ts.dim() == 1:
            ts = ts[None, :].expand(pts.shape[0], 1)
        with torch.no_grad():
            o = self.point_forward(pts, dirs, ts)
        g = o["g_o"]
        normal = g / (torch.linalg.norm(g, ord=2, dim=-1, keepdim=True) + 1e-10)
        color, normal = o["rgb"].reshape(*sh, 3), normal.reshape(*sh, 3)
        return (color, normal.cpu().numpy()) if cpu else (color, normal)

    def renderondepth(self, rays, depth):
        """endosurf.py:451-488: surface rendering at a given per-ray depth."""
        rays = rays.float()
        n_rays = rays.shape[0]
        rays_o, rays_d, rays_d_z = self._split_rays(rays)
        time = rays[..., 8]
        d1 = -torch.sum(rays_d * rays_o, dim=-1) / torch.sum(rays_d * rays_d, dim=-1)
        pm = rays_o + d1.unsqueeze(-1) * rays_d
        far = (d1 + torch.sqrt(torch.clamp(1.0 - torch.sum(pm * pm, dim=-1), min=0.0)) / torch.norm(rays_d, dim=-1))[..., None]
        valid = (depth[..., 0] > 0) & (depth[..., 0] != np.inf)
        d_out = torch.where(depth == np.inf, far, depth)
        d_safe = torch.where(valid[:, None], depth, torch.ones_like(depth))
        with torch.no_grad():
            o = self.point_forward(rays_o + rays_d_z * d_safe, rays_d, time.unsqueeze(-1))
        vm = valid[:, None].to(o["rgb"].dtype)
        return o["rgb"] * vm, o["g_o"] * vm, d_out

    def extract_fields(self, t, bound_min, bound_max, resolution, net_chunk=None, cpu=True):
        """utils.py:139-157: SDF on a resolution^3 grid.  One library call: the grid points are generated on the device
        slab by slab and each fused query writes straight into the output volume (the reference moves 5000-point
        chunks through the host)."""
        self._sync_weights()
        lib, ctx = _lib.load(), self._context()
        dev = self.model.deviation_network.variance.device
        u = torch.empty(resolution, resolution, resolution, device=dev)
        t = torch.as_tensor(t, device=dev, dtype=torch.float32).reshape(-1)[:1].contiguous()
        lo = (C.c_float * 3)(*[float(bound_min[i]) for i in range(3)])
        hi = (C.c_float * 3)(*[float(bound_max[i]) for i in range(3)])
        _lib.check(ctx, lib.es_sdf_grid(ctx, lo, hi, int(resolution), _ptr(t), _ptr(u), self._stream()), "es_sdf_grid")
        return u.cpu().numpy() if cpu else u

    def extract_observation_geometry(self, t, bound_min, bound_max, resolution, threshold=0.0, net_chunk=80000,
                                     cpu=True):
        """endosurf.py:490-500: marching-cubes mesh of the SDF at time t (the grid query is the CUDA part; the
        marching cubes itself is PyMCubes on the CPU exactly as in the reference, utils.py:130-136)."""
        u = self.extract_fields(t, bound_min, bound_max, resolution)
        try:
            import mcubes
        except ImportError as e:  # same third-party dependency as the reference; not a fallback we can replace
            raise ImportError("extract_observation_geometry needs PyMCubes (`mcubes`), like the reference") from e
        vertices, triangles = mcubes.marching_cubes(u, threshold)
        b_max = torch.as_tensor(bound_max).detach().cpu().numpy()
        b_min = torch.as_tensor(bound_min).detach().cpu().numpy()
        vertices = vertices / (resolution - 1.0) * (b_max - b_min)[None, :] + b_min[None, :]
        return vertices, triangles

    def sync_check(self):
        """Synchronise and raise if any kernel tripped its device-side watchdog (tests / debugging)."""
        lib, ctx = _lib.load(), self._context()
        _lib.check(ctx, lib.es_sync_check(ctx, self._stream()), "es_sync_check")

    def release_workspace(self):
        """Free the library's scratch workspace (plane records of the backward, per-point scratch)."""
        lib, ctx = _lib.load(), self._context()
        _lib.check(ctx, lib.es_release_workspace(ctx), "es_release_workspace")

    def profile(self, on: bool):
        """Bracket every fused MLP-chain launch with CUDA events (bench.py roofline)."""
        lib, ctx = _lib.load(), self._context()
        _lib.check(ctx, lib.es_profile_enable(ctx, int(on)), "es_profile_enable")

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        lib, ctx = _lib.load(), self._context()
        p = _lib.EsProfile()
        _lib.check(ctx, lib.es_profile_read(ctx, C.byref(p), self._stream()), "es_profile_read")
        names = ["geometry_chain", "color_chain", "sdf_query_chain", "rev_deform_chain", "rev_sdf_chain",
                 "rev_color_chain", "input_adjoint", "wgrad", "wgrad_reduce", "composite", "out_layer_grads"]
        return {n: {"ms": p.ms[i], "launches": int(p.launches[i]), "points": int(p.points[i])}
                for i, n in enumerate(names)}

    def launch_count(self) -> int:
        return int(_lib.load().es_launch_count(self._context()))

    # ------------------------------------------------------------------ network queries (EndoSurfNet surface)
    def sdf_from_observed_space(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """EndoSurfNet.get_sdf_from_observed_space (endosurf.py:570-579); no grad."""
        self._sync_weights()
        lib, ctx = _lib.load(), self._context()
        x = x.detach().reshape(-1, 3).contiguous().float()
        n = x.shape[0]
        t = t.detach().reshape(-1).contiguous().float()
        t_div = 1 if t.numel() == n else n
        out = torch.empty(n, 1, device=x.device, dtype=torch.float32)
        _lib.check(ctx, lib.es_sdf_query(ctx, _ptr(x), _ptr(t), t_div, 1, n, _ptr(out), self._stream()),
                   "es_sdf_query")
        return out

    def point_forward(self, x, d, t, want_feat=False) -> Dict[str, torch.Tensor]:
        """EndoSurfNet.forward + the gradient queries (endosurf.py:581-689) on explicit points; no grad.
        Returns x_c, jac, sdf, g_c, g_o, rgb (and feat)."""
        self._sync_weights()
        lib, ctx = _lib.load(), self._context()
        x = x.detach().reshape(-1, 3).contiguous().float()
        d = d.detach().reshape(-1, 3).contiguous().float() if d is not None else None  # None: geometry only, no rgb
        n = x.shape[0]
        t = t.detach().reshape(-1).contiguous().float()
        t_div = 1 if t.numel() == n else n
        dev = x.device
        o = dict(x_c=torch.empty(n, 3, device=dev), jac=torch.empty(n, 3, 3, device=dev),
                 sdf=torch.empty(n, 1, device=dev), g_c=torch.empty(n, 3, device=dev))
        if d is not None:
            o["rgb"] = torch.empty(n, 3, device=dev)
        if want_feat:
            o["feat"] = torch.empty(n, 256, device=dev)
        _lib.check(ctx, lib.es_point_forward(ctx, _ptr(x), _ptr(t), t_div, 1, _ptr(d), 1, 3, n, _ptr(o["x_c"]),
                                             _ptr(o["jac"]), _ptr(o["sdf"]), _ptr(o["g_c"]), _ptr(o.get("feat")),
                                             _ptr(o.get("rgb")), self._stream()), "es_point_forward")
        o["g_o"] = torch.einsum("nij,ni->nj", o["jac"], o["g_c"])
        return o

    def up_sample(self, rays_o, rays_d, z_vals, sdf, n_importance, inv_s):
        """endosurf.py:221-266"""
        lib, ctx = _lib.load(), self._context()
        R, n = z_vals.shape
        rays = torch.cat([rays_o, rays_d, torch.zeros(R, 3, device=z_vals.device)], -1).contiguous().float()
        u = torch.linspace(0.0 + 0.5 / n_importance, 1.0 - 0.5 / n_importance, steps=n_importance,
                           device=z_vals.device)
        out = torch.empty(R, n_importance, device=z_vals.device)
        _lib.check(ctx, lib.es_up_sample(ctx, _ptr(rays), R, _ptr(z_vals.contiguous().float()),
                                         _ptr(sdf.reshape(R, n).contiguous().float()), n, n_importance, _ptr(u),
                                         float(inv_s), _ptr(out), self._stream()), "es_up_sample")
        return out

    def cat_z_vals(self, rays_o, rays_d, time, z_vals, new_z_vals, sdf, last=False):
        """endosurf.py:268-287: merge new samples into the sorted z list; unless `last`, query the SDF at the new
        samples (fused deform + SDF chain) and permute it alongside.  (render_rays itself runs this step as
        merge_z_kernel inside es_render_rays; this method keeps the reference's name and signature for callers.)"""
        R, n = z_vals.shape
        z_all, index = torch.sort(torch.cat([z_vals, new_z_vals], dim=-1), dim=-1)
        if not last:
            rays_d_z = rays_d / (rays_d[..., 2:] + 1e-6)
            pts = rays_o[:, None, :] + rays_d_z[:, None, :] * new_z_vals[..., :, None]
            t = time.reshape(R, 1, 1).expand(R, new_z_vals.shape[1], 1)
            new_sdf = self.sdf_from_observed_space(pts.reshape(-1, 3), t.reshape(-1, 1)).reshape(R, -1)
            sdf = torch.gather(torch.cat([sdf.reshape(R, n), new_sdf], dim=-1), 1, index)
        return z_all, sdf

    def render_core(self, rays_o, rays_d, time, z_vals, sample_dist, cos_anneal_ratio=0.0, eval=False):
        """endosurf.py:134-213 under the reference's name and signature: mid-points, the three networks with normals
        and Jacobian, NeuS compositing on the given z_vals.  sample_dist must be 2 / n for an integer n <= the number
        of samples (the reference always passes 2 / n_samples, endosurf.py:73)."""
        R, M = z_vals.shape
        ns = int(round(2.0 / float(sample_dist)))
        if ns < 2 or abs(2.0 / ns - float(sample_dist)) > 1e-6 * float(sample_dist) or M < ns:
            raise NotImplementedError("render_core: sample_dist must equal 2 / n with 2 <= n <= z_vals.shape[1]")
        dev = z_vals.device
        rays = torch.cat([rays_o.reshape(R, 3), rays_d.reshape(R, 3), torch.zeros(R, 2, device=dev),
                          time.reshape(R, 1)], dim=-1).float()
        saved = (self.n_samples, self.n_importance, self.important_begin_iter)
        try:
            self.n_samples, self.n_importance, self.important_begin_iter = ns, M - ns, 0
            o = self.render_rays(rays, iter_step=0, perturb_overwrite=False, z_vals_override=z_vals,
                                 cos_anneal_ratio=float(cos_anneal_ratio))
        finally:
            self.n_samples, self.n_importance, self.important_begin_iter = saved
        return {"color_map": o["color_map"], "depth_map": o["depth_map"], "gradients_o": o["gradients_o"],
                "gradient_o_error": o["gradient_o_error"], "cdf": o["cdf"], "weights": o["weights"],
                "s_val": o["s_val"].expand(R, M).reshape(-1, 1)}

    # ------------------------------------------------------------------ the hot path
    def render_rays(self, rays, iter_step=0, perturb_overwrite=None, eval=False, z_vals_override=None,
                    return_extras=False, cos_anneal_ratio=None, **kwargs):
        """EndoSurfRenderer.render_rays (endosurf.py:60-132): rays [R,9] -> the reference's 8-key dict.
        cos_anneal_ratio: overrides get_cos_anneal_ratio(iter_step) (render_core passes the caller's value)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()):
            return self._render_rays_train(rays, iter_step, perturb_overwrite, z_vals_override, return_extras,
                                           cos_anneal_ratio)
        self._sync_weights()
        lib, ctx = _lib.load(), self._context()
        rays = rays.detach().contiguous().float()
        dev = rays.device
        R = rays.shape[0]
        ns, ni = int(self.n_samples), int(self.n_importance)
        perturb = self.perturb if perturb_overwrite is None else perturb_overwrite
        do_up = bool(iter_step >= self.important_begin_iter and ni > 0)
        M = ns + (ni if do_up else 0)
        t_vals = self._const(("t", ns, dev), lambda: torch.linspace(0.0, 1.0, ns, device=dev))
        steps = int(self.up_sample_steps)
        u_vals = None
        if do_up:
            k = ni // steps
            u_vals = self._const(("u", k, dev),
                                 lambda: torch.linspace(0.0 + 0.5 / k, 1.0 - 0.5 / k, steps=k, device=dev))
        t_rand = None
        if perturb and z_vals_override is None:
            t_rand = (torch.rand([R, 1], device=dev) - 0.5).reshape(-1).contiguous()  # endosurf.py:81
        zo = None
        if z_vals_override is not None:
            zo = z_vals_override.detach().contiguous().float()
        out = {
            "color_map": torch.empty(R, 3, device=dev), "depth_map": torch.empty(R, 1, device=dev),
            "gradients_o": torch.empty(R, M, 3, device=dev), "gradient_o_error": torch.empty((), device=dev),
            "weights": torch.empty(R, M, device=dev), "weight_max": torch.empty(R, 1, device=dev),
            "cdf": torch.empty(R, M, device=dev), "s_val": torch.empty(R, 1, device=dev),
        }
        extras = {}
        if return_extras:
            extras = {"z_vals": torch.empty(R, M, device=dev), "sdf": torch.empty(R, M, device=dev),
                      "sampled_color": torch.empty(R, M, 3, device=dev)}
        if zo is not None and zo.shape != (R, M):
            raise ValueError(f"z_vals_override must be [{R},{M}] for this render config, got {tuple(zo.shape)}")
        prm = _lib.EsRenderParams(
            n_samples=ns, n_importance=ni, up_sample_steps=steps, do_upsample=int(do_up),
            cos_anneal_ratio=float(self.get_cos_anneal_ratio(iter_step) if cos_anneal_ratio is None
                                   else cos_anneal_ratio),
            variance=self.model.deviation_network.variance.data_ptr(), t_vals=t_vals.data_ptr(),
            u_vals=u_vals.data_ptr() if u_vals is not None else None,
            t_rand=t_rand.data_ptr() if t_rand is not None else None,
            z_override=zo.data_ptr() if zo is not None else None)
        o = _lib.EsRenderOut(**{k: v.data_ptr() for k, v in {**out, **extras}.items()})
        rc = lib.es_render_rays(ctx, _ptr(rays), R, C.byref(prm), C.byref(o), self._stream())
        _lib.check(ctx, rc, "es_render_rays")
        out.update(extras)
        return out
