"""Differentiable (training) path: thin ``autograd.Function`` wrappers over the C ABI.

The reference differentiates ``render_rays`` with ``torch.autograd`` (``loss.backward()`` at
``src/trainer/trainer_endosurf.py:94-104`` through ``render_core``, ``src/renderer/endosurf.py:134-213``, and the
``create_graph=True`` gradient queries at ``endosurf.py:594-658``).  Here both directions are library calls:

* :class:`RenderFn`      forward = ``es_render_train_forward`` (fused geometry + colour chains keeping fp16 plane
  records of every MMA layer's input rows, fused compositing); backward = ``es_render_train_backward`` (compositing
  backward kernel, three reverse tcgen05 chains, input-adjoint launches, split-K tcgen05 weight-gradient kernel,
  weight-norm backward).
* :class:`PointFieldFn`  the same for explicit points (``errorondepth`` / ``surface_neighbour_error`` during training,
  ``endosurf.py:289-342``).

Gradients are returned directly for the reference's parameters (``bias``, ``weight_g``, ``weight_v`` of every layer and
``variance``); nothing on this path is a PyTorch op or a library GEMM - PyTorch only owns the buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def net_ids(renderer) -> List[int]:
    return ([0] if renderer.model.use_deform else []) + [1, 2]


def param_list(renderer) -> List[torch.Tensor]:
    """[bias, weight_g, weight_v] per layer per network, in ``module.parameters()`` order (without ``variance``)."""
    return renderer._fast_params()[:-1]


class _StaticTables:
    """The part of es_train_params that only depends on WHERE the parameters live: the weight_v / weight_g pointer
    tables, the layout of the flat gradient buffer and the ctypes arrays the gradient pointers are written into.
    Built once per renderer (re-built when a parameter moves) instead of once per backward call: 3 backward calls per
    trainer step (render, errorondepth, surface_neighbour_error) used to spend about 1 ms of host time here each."""

    def __init__(self, renderer, params: Tuple[torch.Tensor, ...]):
        L = renderer._cfg_struct.n_layers
        self.sizes = [p.numel() for p in params]
        self.shapes = [p.shape for p in params]
        self.total = sum(self.sizes)
        offs, off = [], 0
        self.view_args = []  # (shape, contiguous strides, element offset) of every gradient inside the flat buffer
        for p, n in zip(params, self.sizes):
            offs.append(4 * off)
            self.view_args.append((tuple(p.shape), tuple(torch.empty(0).new_empty(p.shape).stride()) if p.dim() > 2
                                   else ((p.shape[1], 1) if p.dim() == 2 else ((1,) if p.dim() == 1 else ())), off))
            off += n
        names = ("v", "g", "grad_v", "grad_g", "grad_b")
        fields = {k: [None, None, None] for k in names}
        self.keep = []
        self.grad_tables = []  # (ctypes array of L pointers, int64 byte offsets of its L gradients in the flat buffer)
        k = 0
        for net in net_ids(renderer):
            tabs = {n: (C.c_void_p * L)() for n in names}
            goff = {n: np.zeros(L, dtype=np.int64) for n in ("grad_b", "grad_g", "grad_v")}
            for l in range(L):
                g, v = params[k + 1], params[k + 2]  # (bias, weight_g, weight_v)
                tabs["v"][l], tabs["g"][l] = v.data_ptr(), g.data_ptr()
                goff["grad_b"][l], goff["grad_g"][l], goff["grad_v"][l] = offs[k], offs[k + 1], offs[k + 2]
                k += 3
            for n in names:
                fields[n][net] = C.cast(tabs[n], C.POINTER(C.c_void_p))
                self.keep.append(tabs[n])
            self.grad_tables += [(tabs[n], goff[n]) for n in goff]
        self.struct = _lib.EsTrainParams(**{n: (C.POINTER(C.c_void_p) * 3)(*[q if q is not None else C.POINTER(C.c_void_p)()
                                                                           for q in fields[n]]) for n in names})


def _static_tables(renderer, params) -> _StaticTables:
    key = tuple([p.data_ptr() for p in params])
    cached = renderer.__dict__.get("_ptab_static")
    if cached is None or cached[0] != key:
        cached = (key, _StaticTables(renderer, params))
        renderer.__dict__["_ptab_static"] = cached
    return cached[1]


class ParamHubFn(torch.autograd.Function):
    """(params...) -> handle [total]: a tensor whose only purpose is its autograd edge.

    The library's backward calls produce ALL parameter gradients as one flat buffer.  A trainer step has three such
    calls (render_rays, errorondepth, surface_neighbour_error); handing 81 views per call to autograd made it add them
    pairwise - 164 tiny launches per step.  Instead every call takes this handle as its one differentiable
    "parameter" input and returns the flat buffer as the handle's gradient: autograd sums the flat buffers (2 adds) and
    this node hands the 81 per-parameter views of the sum to the parameters once."""

    @staticmethod
    def forward(ctx, renderer, *params):
        st = _static_tables(renderer, params)
        ctx.view_args = st.view_args
        ctx.set_materialize_grads(False)  # a bound gradient sink leaves nothing to route: no zero-filled stand-in
        return torch.empty(st.total, device=params[0].device, dtype=torch.float32)

    @staticmethod
    def backward(ctx, flat):
        if flat is None:
            return (None,) * (1 + len(ctx.view_args))
        flat = flat.contiguous()
        base = flat.storage_offset()
        return (None, *[flat.as_strided(shp, std, base + off) for shp, std, off in ctx.view_args])


def param_hub(renderer):
    """-> (params, handle): the network parameters ([bias, weight_g, weight_v] per layer per network) and the hub handle
    that stands for them in the autograd graph.  The hub node only routes gradients to parameter OBJECTS, so it is kept
    across calls and iterations until a parameter object, its requires_grad flag or its device changes."""
    params = tuple(renderer._fast_params()[:-1])
    key = (tuple([id(p) for p in params]), tuple([p.requires_grad for p in params]), params[0].device)
    cached = renderer.__dict__.get("_hub")
    if cached is None or cached[0] != key or not cached[2].requires_grad or not torch.is_grad_enabled():
        handle = ParamHubFn.apply(renderer, *params)
        cached = (key, params, handle)  # the parameters are kept alive with the key: their ids stay valid
        if handle.requires_grad:
            renderer.__dict__["_hub"] = cached
    return cached[1], cached[2]


class _ParamTables:
    """ctypes view (es_train_params) of the parameters and freshly allocated gradient tensors."""

    def __init__(self, renderer, params: Tuple[torch.Tensor, ...]):
        st = _static_tables(renderer, params)
        # ONE zero-filled buffer for all gradients in parameter order (one fill launch; the bias gradients are
        # accumulated by the kernels), handed out as views.  The library reads the pointer tables on the host while
        # the call is being enqueued, so the (cached) ctypes arrays can be re-pointed for every backward call.
        self.flat = torch.zeros(st.total, device=params[0].device, dtype=torch.float32)
        self._st = st
        base = self.flat.data_ptr()
        for arr, off in st.grad_tables:
            ptrs = off + base
            C.memmove(arr, ptrs.ctypes.data, ptrs.nbytes)
        self.struct = st.struct

    @property
    def grads(self) -> List[torch.Tensor]:
        """The gradients as per-parameter views of the flat buffer (only built when autograd needs them: a bound
        gradient sink takes the flat buffer)."""
        flat = self.flat
        return [flat.as_strided(shp, std, off) for shp, std, off in self._st.view_args]


def _deliver(renderer, params, tabs, extra):
    """Hand the gradients of a backward call to autograd as ONE flat tensor (the gradient of the hub handle,
    :class:`ParamHubFn`) - or, when a gradient sink is bound to the renderer (``distributed.FlatGradBucket.bind``: every
    ``p.grad`` is a view into one flat buffer laid out in exactly this parameter order), add them there with ONE launch
    and return ``None``.  ``extra`` = [(parameter, gradient)] delivered the same way.  -> (extra gradients, flat)"""
    sink = getattr(renderer, "_grad_sink", None)
    if sink is not None and sink.accumulate(params, tabs.flat, extra):
        return [None] * len(extra), None
    return [g for _, g in extra], tabs.flat


def _stash(renderer, n: int, dev) -> torch.Tensor:
    lib, ectx = _lib.load(), renderer._context()
    nbytes = C.c_int64()
    _lib.check(ectx, lib.es_train_stash_bytes(ectx, n, C.byref(nbytes)), "es_train_stash_bytes")
    return torch.empty(int(nbytes.value), dtype=torch.uint8, device=dev)


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().contiguous().float()


class RenderFn(torch.autograd.Function):
    """(rays [R,9], z_vals [R,M], variance, hub handle, params) -> (color_map, depth_map, gradients_o, gradient_o_error,
    weights, cdf, sdf, sampled_color, weight_max, s_val, eikonal_den); the last three are not differentiable here (the
    reference only logs weight_max / s_val; eikonal_den = sum(relax) + 1e-6 is what data-parallel training needs to
    re-normalise the eikonal mean over all ranks)."""

    @staticmethod
    def forward(ctx, renderer, rays, z, cos_ratio, variance, hub, params):
        """hub: the handle of :func:`param_hub` (its gradient is the flat parameter-gradient buffer); params: the
        parameter tuple it stands for (opaque to autograd)."""
        lib, ectx = _lib.load(), renderer._context()
        renderer._sync_weights()
        dev = rays.device
        R, M = z.shape
        n = R * M
        deform = renderer.model.use_deform
        rays = rays.detach().contiguous().float()
        z = z.detach().contiguous().float()
        x_c = torch.empty(n, 3, device=dev)
        jac = torch.empty(n, 9, device=dev) if deform else None
        sdf = torch.empty(R, M, device=dev)
        g_c = torch.empty(n, 3, device=dev)
        rgb = torch.empty(R, M, 3, device=dev)
        stash = _stash(renderer, n, dev)
        out = {
            "color_map": torch.empty(R, 3, device=dev), "depth_map": torch.empty(R, 1, device=dev),
            "gradients_o": torch.empty(R, M, 3, device=dev), "gradient_o_error": torch.empty((), device=dev),
            "weights": torch.empty(R, M, device=dev), "weight_max": torch.empty(R, 1, device=dev),
            "cdf": torch.empty(R, M, device=dev), "s_val": torch.empty(R, 1, device=dev),
        }
        eik_den = torch.empty(1, device=dev)
        o = _lib.EsRenderOut(**{k: v.data_ptr() for k, v in out.items()})
        rc = lib.es_render_train_forward(ectx, _ptr(rays), R, _ptr(z), M, int(renderer.n_samples), float(cos_ratio),
                                         _ptr(variance), _ptr(x_c), _ptr(jac), _ptr(sdf), _ptr(g_c), _ptr(rgb),
                                         _ptr(stash), C.byref(o), _ptr(eik_den), renderer._stream())
        _lib.check(ectx, rc, "es_render_train_forward")
        renderer._poll_device_error()
        ctx.renderer = renderer
        ctx.meta = (R, M, float(cos_ratio), renderer._packed_version)
        ctx.params = params
        ctx.save_for_backward(rays, z, variance, x_c, jac, sdf, g_c, rgb, stash, eik_den)
        eik_den_out = eik_den.clone()
        ctx.mark_non_differentiable(out["weight_max"], out["s_val"], eik_den_out)
        return (out["color_map"], out["depth_map"], out["gradients_o"], out["gradient_o_error"], out["weights"],
                out["cdf"], sdf, rgb, out["weight_max"], out["s_val"], eik_den_out)

    @staticmethod
    def backward(ctx, color_b, depth_b, go_b, eik_b, w_b, cdf_b, sdf_b, rgb_b, _wm_b, _sv_b, _ed_b):
        renderer = ctx.renderer
        lib, ectx = _lib.load(), renderer._context()
        R, M, cos_ratio, version = ctx.meta
        if renderer._packed_version != version:
            raise RuntimeError("parameters changed between the forward and the backward of a render call: the "
                               "context no longer holds the weights the plane records were produced with")
        rays, z, variance, x_c, jac, sdf, g_c, rgb, stash, eik_den = ctx.saved_tensors
        tabs = _ParamTables(renderer, ctx.params)
        bars = [_c(t) for t in (color_b, depth_b, go_b, eik_b, w_b, cdf_b, sdf_b, rgb_b)]
        bar = _lib.EsRenderGrads(color_map=_ptr(bars[0]), depth_map=_ptr(bars[1]), gradients_o=_ptr(bars[2]),
                                 gradient_o_error=_ptr(bars[3]), weights=_ptr(bars[4]), cdf=_ptr(bars[5]),
                                 sdf=_ptr(bars[6]), sampled_color=_ptr(bars[7]))
        var_grad = torch.empty(1, device=rays.device)
        rc = lib.es_render_train_backward(ectx, _ptr(rays), R, _ptr(z), M, int(renderer.n_samples), cos_ratio,
                                          _ptr(variance), _ptr(x_c), _ptr(jac), _ptr(sdf), _ptr(g_c), _ptr(rgb),
                                          _ptr(stash), _ptr(eik_den), C.byref(bar), C.byref(tabs.struct),
                                          _ptr(var_grad), renderer._stream())
        _lib.check(ectx, rc, "es_render_train_backward")
        renderer._poll_device_error()
        (vg,), flat = _deliver(renderer, ctx.params, tabs, [(variance, var_grad.reshape(variance.shape))])
        return (None, None, None, None, vg, flat, None)


class PointFieldFn(torch.autograd.Function):
    """(x [n,3], d [n,3], t [n,1], hub handle, params) -> (sdf [n,1], g_c [n,3], jac [n,3,3], rgb [n,3])."""

    @staticmethod
    def forward(ctx, renderer, x, d, t, hub, params):
        lib, ectx = _lib.load(), renderer._context()
        renderer._sync_weights()
        dev = x.device
        n = x.shape[0]
        deform = renderer.model.use_deform
        x = x.detach().reshape(n, 3).contiguous().float()
        d = d.detach().reshape(n, 3).contiguous().float()
        t = t.detach().reshape(-1).contiguous().float()
        x_c = torch.empty(n, 3, device=dev)
        jac = torch.empty(n, 3, 3, device=dev) if deform else None
        sdf = torch.empty(n, 1, device=dev)
        g_c = torch.empty(n, 3, device=dev)
        rgb = torch.empty(n, 3, device=dev)
        stash = _stash(renderer, n, dev)
        rc = lib.es_point_train_forward(ectx, _ptr(x), _ptr(t), 1, 1, _ptr(d), 1, 3, n, _ptr(x_c), _ptr(jac),
                                        _ptr(sdf), _ptr(g_c), _ptr(rgb), _ptr(stash), renderer._stream())
        _lib.check(ectx, rc, "es_point_train_forward")
        renderer._poll_device_error()
        ctx.renderer = renderer
        ctx.meta = (n, renderer._packed_version)
        ctx.params = params
        ctx.save_for_backward(d, x_c, jac, g_c, rgb, stash)
        jac_out = jac if deform else torch.eye(3, device=dev).expand(n, 3, 3).contiguous()
        return sdf, g_c, jac_out, rgb

    @staticmethod
    def backward(ctx, sdf_b, gc_b, jac_b, rgb_b):
        renderer = ctx.renderer
        lib, ectx = _lib.load(), renderer._context()
        n, version = ctx.meta
        if renderer._packed_version != version:
            raise RuntimeError("parameters changed between the forward and the backward of a point_field call")
        d, x_c, jac, g_c, rgb, stash = ctx.saved_tensors
        tabs = _ParamTables(renderer, ctx.params)
        sdf_b, gc_b, jac_b, rgb_b = _c(sdf_b), _c(gc_b), _c(jac_b), _c(rgb_b)
        rc = lib.es_point_train_backward(ectx, n, _ptr(d), 1, 3, _ptr(x_c), _ptr(jac), _ptr(g_c), _ptr(rgb),
                                         _ptr(stash), _ptr(sdf_b), _ptr(gc_b), _ptr(jac_b), _ptr(rgb_b),
                                         C.byref(tabs.struct), renderer._stream())
        _lib.check(ectx, rc, "es_point_train_backward")
        renderer._poll_device_error()
        _, flat = _deliver(renderer, ctx.params, tabs, [])
        return (None, None, None, None, flat, None)
