"""Differentiable (training) point pipeline on top of the fused sm_100a chains.

The reference differentiates ``EndoSurfNet.forward`` with ``torch.autograd`` (``create_graph=True`` at
``src/renderer/endosurf.py:594-658``).  Here :class:`PointFieldFn` is a custom ``autograd.Function``:

* forward  = ``es_point_forward_train``: the same fused tcgen05 kernels as inference, additionally keeping every MMA
  layer's input rows (primal + 3 forward-mode tangent rows per point) as fp16 hi/lo planes (the "stash");
* backward = three fused reverse chains (``es_point_backward``: colour, sdf, deform) that push the output adjoints
  through the transposed weights on the tensor cores, gate them with the stashed activations (ReLU mask / softplus'
  and softplus'' cross terms of the tangent rows) and write the adjoint of every forward pre-activation ("zbar").
  Weight gradients are then plain ``zbar^T @ stash`` GEMMs (library GEMMs on the stashed planes), and the adjoints of
  the small per-point quantities (positional encodings, J d normalisation) are closed-form elementwise PyTorch.

Gradients are returned with respect to the *effective* weights ``W = g v/||v||`` and biases; PyTorch's own autograd
carries them on to ``weight_g`` / ``weight_v`` (the weight-norm fold is differentiable plumbing, reference
utils.py:57-58).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List

import torch

from . import _lib

SQRT2 = math.sqrt(2.0)


# ------------------------------------------------------------------------------------------------ small torch helpers
def freq_enc(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """[x, sin(2^k x), cos(2^k x)]_k in the reference's column order (encoder.py:40-54).  All frequencies in one
    sin and one cos launch (x * 2^k is exact in fp32, so the values are those of the per-frequency loop)."""
    if n_freqs == 0:
        return x
    f = torch.exp2(torch.arange(n_freqs, device=x.device, dtype=x.dtype))
    arg = x[:, None, :] * f[None, :, None]                                  # [n, L, dim]
    sc = torch.stack([torch.sin(arg), torch.cos(arg)], 2)                   # [n, L, 2, dim]
    return torch.cat([x, sc.reshape(x.shape[0], -1)], -1)


def freq_enc_tangent(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """d enc(x) / d x_j for j = 0..2  ->  [n, 3, 3(2L+1)] (what the kernels feed to the tangent rows)."""
    n, dim = x.shape
    f = torch.exp2(torch.arange(n_freqs, device=x.device, dtype=x.dtype))
    arg = x[:, None, :] * f[None, :, None]
    d = torch.stack([f[None, :, None] * torch.cos(arg), -f[None, :, None] * torch.sin(arg)], 2)  # [n, L, 2, dim]
    dfull = torch.cat([torch.ones_like(x), d.reshape(n, -1)], -1)  # derivative of every column wrt its own component
    comp = (torch.arange(dfull.shape[1], device=x.device) % dim)
    sel = torch.stack([(comp == j) for j in range(dim)], 0).to(x.dtype)  # [3, width]
    return dfull[:, None, :] * sel[None, :, :]


def rows_from_points(v: torch.Tensor) -> torch.Tensor:
    """[p_g, 4, k] (point, stream) -> [4 p_g, k] in the geometry kernels' tile row order: a 128-row tile holds 32
    points; stream s of point (quadrant Q, p) is tile row 32 Q + 8 s + p (csrc/es_mlp.cu, tangent mode)."""
    pg = v.shape[0]
    return v.reshape(pg // 32, 4, 8, 4, -1).permute(0, 1, 3, 2, 4).reshape(pg * 4, -1)


def points_from_rows(r: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`rows_from_points`: [rows, k] -> [rows / 4, 4, k]."""
    rows = r.shape[0]
    return r.reshape(rows // 128, 4, 4, 8, -1).permute(0, 1, 3, 2, 4).reshape(rows // 4, 4, -1)


def stream_rows(r: torch.Tensor, s: int) -> torch.Tensor:
    """Rows of stream s (0 = primal) of a [rows, k] geometry plane in point order -> [rows / 4, k] (copies a quarter)."""
    rows = r.shape[0]
    return r.view(rows // 128, 4, 4, 8, -1)[:, :, s].reshape(rows // 4, -1)


_MM_OUT_DTYPE_OK = None
# fp16 hi/lo terms of the weight-gradient GEMMs.  1 = zbar_hi^T @ stash_hi with fp32 accumulation: the rounding of the
# individual products (2^-11) averages out over the >= 10^5 rows of a batch - measured on a 1024-ray batch against
# exact fp32 products: whole-gradient deviation 6.7e-5, worst parameter tensor 1.6e-4 (3 terms: 2.5e-5 / 4e-5, which
# is the fp32 summation noise of the comparison itself; tools/wgrad_terms_check.py).  The forward pass and the
# activation-gradient chains keep the full 3-term products.
WGRAD_TERMS = 1
FAST_MIN_ROWS = 32768  # below this the exact fp32 product is cheap (tests); above, fp16 tensor-core library GEMMs


def _mm16(a16: torch.Tensor, b16: torch.Tensor) -> torch.Tensor:
    return torch.mm(a16, b16, out_dtype=torch.float32)


def split16(x: torch.Tensor):
    """fp32 -> (hi, lo) fp16 with x ~= hi + lo (22 bits), the same split the kernels use."""
    hi = x.to(torch.float16)
    return hi, (x - hi.float()).to(torch.float16)


def tn_planes(a_hi, a_lo, b_hi, b_lo) -> torch.Tensor:
    """A^T @ B for A [rows, a], B [rows, b] held as fp16 hi/lo planes -> [a, b] fp32.
    Large row counts: three fp16 tensor-core library GEMMs with fp32 output (hi*hi + hi*lo + lo*hi, the split the
    fused kernels use; these are plain GEMMs, cuBLAS picks a split-K kernel); small ones: the exact fp32 product."""
    global _MM_OUT_DTYPE_OK
    ah, al = a_hi.view(torch.float16), a_lo.view(torch.float16)
    bh, bl = b_hi.view(torch.float16), b_lo.view(torch.float16)
    if ah.shape[0] >= FAST_MIN_ROWS and _MM_OUT_DTYPE_OK is not False:
        try:
            at_h, at_l = ah.t(), al.t()
            out = _mm16(at_h, bh)
            if WGRAD_TERMS >= 2:
                out += _mm16(at_h, bl)
            if WGRAD_TERMS >= 3:
                out += _mm16(at_l, bh)
            _MM_OUT_DTYPE_OK = True
            return out
        except (TypeError, RuntimeError):
            _MM_OUT_DTYPE_OK = False
    if WGRAD_TERMS == 1:  # the lo planes may not have been written (es_set_plane_mode)
        return ah.float().t() @ bh.float()
    return (ah.float() + al.float()).t() @ (bh.float() + bl.float())


def planes_mm(a_hi, a_lo, w: torch.Tensor) -> torch.Tensor:
    """A @ w for A [rows, 256] held as fp16 hi/lo planes and a small fp32 w [256, k] -> [rows, k] fp32 (input-layer
    adjoints).  Large row counts: hi@w_hi + hi@w_lo + lo@w_hi on the tensor cores; small ones: the exact product."""
    ah, al = a_hi.view(torch.float16), a_lo.view(torch.float16)
    k = w.shape[1]
    if ah.shape[0] >= FAST_MIN_ROWS and _MM_OUT_DTYPE_OK is not False:
        kp = (k + 15) // 16 * 16
        wp = torch.zeros(w.shape[0], kp, device=w.device)
        wp[:, :k] = w
        wh, wl = split16(wp)
        try:
            out = _mm16(ah, wh)
            out += _mm16(ah, wl)
            out += _mm16(al, wh)
            return out[:, :k]
        except (TypeError, RuntimeError):
            pass
    return (ah.float() + al.float()) @ w


def rowsum_planes(hi, lo, sel16: torch.Tensor) -> torch.Tensor:
    """sum over the rows selected by the 0/1 fp16 row vector sel16 [1, rows] of a hi/lo plane pair -> [cols] fp32."""
    h, l = hi.view(torch.float16), lo.view(torch.float16)
    if h.shape[0] >= FAST_MIN_ROWS and _MM_OUT_DTYPE_OK is not False:
        try:
            return (_mm16(sel16, h) + _mm16(sel16, l))[0] if WGRAD_TERMS >= 2 else _mm16(sel16, h)[0]
        except (TypeError, RuntimeError):
            pass
    if WGRAD_TERMS == 1:
        return (sel16.float() @ h.float())[0]
    return (sel16.float() @ (h.float() + l.float()))[0]


def _pow2_scale(*tensors) -> torch.Tensor:
    """Power-of-two loss scale so that the largest adjoint entering a reverse chain is ~16 (fp16 hi/lo planes keep
    22 bits relative to that; no host sync)."""
    amax = torch.stack([t.detach().abs().max() for t in tensors if t is not None and t.numel() > 0]).max()
    amax = torch.where(amax > 0, amax, torch.ones_like(amax))
    return torch.exp2(torch.floor(torch.log2(16.0 / amax)))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _upload(renderer, epoch, nets, ws, bs, L):
    """Pack this call's effective weights into the context unless it already holds exactly them (same epoch)."""
    if getattr(renderer, "_loaded_epoch", None) == epoch:
        return
    lib, ectx = _lib.load(), renderer._context()
    for net in nets:
        wp = (C.c_void_p * L)(*[w.data_ptr() for w in ws[net]])
        bp = (C.c_void_p * L)(*[b.data_ptr() for b in bs[net]])
        _lib.check(ectx, lib.es_load_network(ectx, net, wp, bp, renderer._stream()), "es_load_network")
    renderer._loaded_epoch = epoch
    renderer._packed_version = None  # the context now holds these weights, not necessarily the module's


class PointFieldFn(torch.autograd.Function):
    """(x, d, t, effective weights...) -> (sdf [n,1], g_c [n,3], jac [n,3,3], rgb [n,3])."""

    @staticmethod
    def forward(ctx, renderer, epoch, x, d, t, *wb):
        lib, ectx = _lib.load(), renderer._context()
        stream = renderer._stream()
        n = x.shape[0]
        dev = x.device
        use_deform = renderer.model.use_deform
        L = renderer._cfg_struct.n_layers
        nets = ([0] if use_deform else []) + [1, 2]
        # unpack [W_0..W_{L-1}, b_0..b_{L-1}] per network and upload (packs forward + transposed units)
        ws, bs, k = {}, {}, 0
        for net in nets:
            ws[net] = [w.detach().contiguous() for w in wb[k:k + L]]
            bs[net] = [b.detach().contiguous() for b in wb[k + L:k + 2 * L]]
            k += 2 * L
        _upload(renderer, epoch, nets, ws, bs, L)
        _lib.check(ectx, lib.es_set_plane_mode(ectx, int(WGRAD_TERMS != 1)), "es_set_plane_mode")
        ctx.full_planes = WGRAD_TERMS != 1
        lay = (C.c_int64 * 6)()
        _lib.check(ectx, lib.es_train_layout(ectx, n, lay), "es_train_layout")
        g_rows, g_slots, c_rows, c_slots, z_slots, sdf_off = [int(v) for v in lay]
        gs_hi = torch.empty(g_slots, g_rows, 256, dtype=torch.int16, device=dev)
        gs_lo = torch.empty_like(gs_hi)
        cs_hi = torch.empty(c_slots, c_rows, 256, dtype=torch.int16, device=dev)
        cs_lo = torch.empty_like(cs_hi)
        x = x.detach().contiguous().float()
        d = d.detach().contiguous().float()
        t = t.detach().reshape(-1).contiguous().float()
        x_c = torch.empty(n, 3, device=dev)
        jac = torch.empty(n, 3, 3, device=dev)
        sdf = torch.empty(n, 1, device=dev)
        g_c = torch.empty(n, 3, device=dev)
        feat = torch.empty(n, 256, device=dev)
        rgb = torch.empty(n, 3, device=dev)
        rc = lib.es_point_forward_train(ectx, _ptr(x), _ptr(t), 1, 1, _ptr(d), 1, 3, n, _ptr(x_c),
                                        _ptr(jac if use_deform else None), _ptr(sdf),
                                        _ptr(g_c), _ptr(feat), _ptr(rgb), _ptr(gs_hi), _ptr(gs_lo), _ptr(cs_hi),
                                        _ptr(cs_lo), stream)
        _lib.check(ectx, rc, "es_point_forward_train")
        if not use_deform:
            jac = torch.eye(3, device=dev).expand(n, 3, 3).contiguous()
        ctx.renderer = renderer
        ctx.epoch = epoch
        ctx.meta = (n, L, use_deform, nets, g_rows, g_slots, c_rows, c_slots, z_slots, sdf_off)
        ctx.ws, ctx.bs = ws, bs
        ctx.save_for_backward(x, d, t, x_c, jac, g_c, feat, rgb, gs_hi, gs_lo, cs_hi, cs_lo)
        return sdf, g_c, jac, rgb

    @staticmethod
    def backward(ctx, sdf_bar, gc_bar, jac_bar, rgb_bar):
        renderer = ctx.renderer
        lib, ectx = _lib.load(), renderer._context()
        stream = renderer._stream()
        n, L, use_deform, nets, g_rows, g_slots, c_rows, c_slots, z_slots, sdf_off = ctx.meta
        x, d, t, x_c, jac, g_c, feat, rgb, gs_hi, gs_lo, cs_hi, cs_lo = ctx.saved_tensors
        dev = x.device
        ws, bs = ctx.ws, ctx.bs
        cfg = renderer._cfg_struct
        skip = cfg.skip_layer
        # the context may have been re-packed by another forward since; make sure it holds THIS call's weights
        _upload(renderer, ctx.epoch, nets, ws, bs, L)
        if not ctx.full_planes and WGRAD_TERMS != 1:
            raise RuntimeError("WGRAD_TERMS changed between forward and backward: the lo planes were not written")
        _lib.check(ectx, lib.es_set_plane_mode(ectx, int(ctx.full_planes)), "es_set_plane_mode")

        def zeros_like_or(tns, shape):
            return tns if tns is not None else torch.zeros(shape, device=dev)

        sdf_bar = zeros_like_or(sdf_bar, (n, 1)).float()
        gc_bar = zeros_like_or(gc_bar, (n, 3)).float()
        jac_bar = zeros_like_or(jac_bar, (n, 3, 3)).float()
        rgb_bar = zeros_like_or(rgb_bar, (n, 3)).float()
        gw = {net: [None] * L for net in nets}
        gb = {net: [None] * L for net in nets}
        p_g = g_rows // 4  # padded point count of the geometry chains
        f16 = torch.float16
        # 0/1 row selectors for bias gradients (tangent rows carry no bias)
        ones_c = torch.ones(1, c_rows, dtype=f16, device=dev)
        prim_g = torch.zeros(p_g, 4, 1, dtype=f16, device=dev)
        prim_g[:, 0] = 1
        prim_g = rows_from_points(prim_g).reshape(1, g_rows)

        def run_reverse(net, adj, adj_feat, stash_hi, stash_lo, rows):
            zb_hi = torch.empty(z_slots, rows, 256, dtype=torch.int16, device=dev)
            zb_lo = torch.empty_like(zb_hi)
            rc = lib.es_point_backward(ectx, net, n, _ptr(stash_hi), _ptr(stash_lo), _ptr(adj), _ptr(adj_feat),
                                       _ptr(zb_hi), _ptr(zb_lo), stream)
            _lib.check(ectx, rc, "es_point_backward")
            return zb_hi, zb_lo

        def padded_planes(v, rows):
            """[n, k] fp32 (plain chains: row = point) or [n, 4, k] (geometry chains: 4 streams per point, kernel
            tile row order) -> hi/lo planes zero-padded to the stash row count."""
            cols = (v.shape[-1] + 15) // 16 * 16   # odd widths push cuBLAS onto legacy kernels
            if v.dim() == 3:
                buf = torch.zeros(rows // 4, 4, cols, device=dev)
                buf[:v.shape[0], :, :v.shape[-1]] = v
                return split16(rows_from_points(buf))
            buf = torch.zeros(rows, cols, device=dev)
            buf[:v.shape[0], :v.shape[1]] = v
            return split16(buf)

        # ============================================================== colour network
        d_c_u = (jac * d[:, None, :]).sum(-1)   # J d; elementwise: batched 3x3 GEMMs cost 0.6 ms each
        d_c = d_c_u / (torch.linalg.norm(d_c_u, dim=-1, keepdim=True) + 1e-10)
        inp_c = torch.cat([freq_enc(x_c, cfg.multires_color_pos), g_c, freq_enc(d_c, cfg.multires_color_dir), feat], -1)
        inp_hi, inp_lo = padded_planes(inp_c, c_rows)
        o_c = rgb_bar * rgb * (1.0 - rgb)                     # through the output sigmoid (endosurf.py:841)
        s_c = _pow2_scale(o_c)
        adj_c = torch.cat([o_c * s_c, torch.zeros(n, 1, device=dev)], -1).contiguous()
        zc_hi, zc_lo = run_reverse(2, adj_c, None, cs_hi, cs_lo, c_rows)
        Wc = ws[2]
        inp_bar = None
        for m in range(0, L - 1):
            if m == 0 or m == skip:  # layers that read the network input: need zbar itself for the input adjoint
                w_in = Wc[0] if m == 0 else Wc[m][:, 256:] / SQRT2
                ib = planes_mm(zc_hi[m], zc_lo[m], w_in)[:n]
                inp_bar = ib if inp_bar is None else inp_bar + ib
            g_in = tn_planes(zc_hi[m], zc_lo[m], inp_hi, inp_lo)[:, :inp_c.shape[1]] if (m == 0 or m == skip) else None
            if m == 0:
                g = g_in
            else:
                g = tn_planes(zc_hi[m], zc_lo[m], cs_hi[m], cs_lo[m])
                if m == skip:
                    g = torch.cat([g, g_in], 1) / SQRT2
            gw[2][m] = g / s_c
            gb[2][m] = rowsum_planes(zc_hi[m], zc_lo[m], ones_c) / s_c
        inp_bar = inp_bar / s_c
        oc_hi, oc_lo = padded_planes(o_c, c_rows)
        gw[2][L - 1] = tn_planes(oc_hi, oc_lo, cs_hi[L - 1], cs_lo[L - 1])[:3]
        gb[2][L - 1] = o_c.sum(0)
        nx = 3 * (1 + 2 * cfg.multires_color_pos)
        nd = 3 * (1 + 2 * cfg.multires_color_dir)
        ex_bar, gc_col_bar, ed_bar, feat_bar = inp_bar[:, :nx], inp_bar[:, nx:nx + 3], \
            inp_bar[:, nx + 3:nx + 3 + nd], inp_bar[:, nx + 3 + nd:]
        # adjoints of x_c (through enc10) and of J (through d_c = normalize(J d)): tiny elementwise graphs
        with torch.enable_grad():
            xc_r = x_c.detach().requires_grad_(True)
            j_r = jac.detach().requires_grad_(True)
            u = (j_r * d[:, None, :]).sum(-1)
            dcr = u / (torch.linalg.norm(u, dim=-1, keepdim=True) + 1e-10)
            obj = (freq_enc(xc_r, cfg.multires_color_pos) * ex_bar).sum() + \
                (freq_enc(dcr, cfg.multires_color_dir) * ed_bar).sum()
            xc_bar, jbar_color = torch.autograd.grad(obj, [xc_r, j_r])
        del zc_hi, zc_lo, inp_hi, inp_lo

        # ============================================================== sdf network
        gc_tot = gc_bar + gc_col_bar
        adj_s = torch.zeros(p_g, 4, 4, device=dev)
        s_s = _pow2_scale(sdf_bar, gc_tot, feat_bar)
        adj_s[:n, 0, 3] = sdf_bar[:, 0] * s_s
        adj_s[:n, 1:, 3] = gc_tot * s_s
        feat_bar_s = (feat_bar * s_s).contiguous()
        zs_hi, zs_lo = run_reverse(1, adj_s, feat_bar_s, gs_hi, gs_lo, g_rows)
        Ws = ws[1]
        with torch.enable_grad():
            xc_r = x_c.detach().requires_grad_(True)
            e0 = freq_enc(xc_r, cfg.multires_sdf_pos)                     # [n, 39]
            et = freq_enc_tangent(xc_r, cfg.multires_sdf_pos)             # [n, 3, 39]
            a0 = torch.cat([e0[:, None, :], et], 1)                       # rows of the first sdf layer [n,4,39]
            a0_hi, a0_lo = padded_planes(a0.detach(), g_rows)
            E = None
            for m in range(0, L - 1):
                if m == 0 or m == skip:
                    w_in = Ws[0] if m == 0 else Ws[m][:, 256:] / SQRT2
                    Em = points_from_rows(planes_mm(zs_hi[m], zs_lo[m], w_in))[:n]
                    E = Em if E is None else E + Em                          # adjoint of the input rows [n,4,39]
                g_in = tn_planes(zs_hi[m], zs_lo[m], a0_hi, a0_lo)[:, :a0.shape[-1]] if (m == 0 or m == skip) else None
                if m == 0:
                    g = g_in
                else:
                    g = tn_planes(zs_hi[m], zs_lo[m], gs_hi[sdf_off + m], gs_lo[sdf_off + m])
                    if m == skip:
                        g = torch.cat([g, g_in], 1) / SQRT2
                gw[1][m] = g / s_s
                gb[1][m] = rowsum_planes(zs_hi[m], zs_lo[m], prim_g) / s_s
            xc_bar = xc_bar + torch.autograd.grad((a0 * (E / s_s).detach()).sum(), xc_r)[0]
        r_rows = torch.cat([sdf_bar, gc_tot], 1)                          # [n,4]: adjoint of the sdf-row output
        rr_hi, rr_lo = padded_planes(r_rows.reshape(n, 4, 1), g_rows)
        sl = sdf_off + L - 1
        g_row0 = tn_planes(rr_hi, rr_lo, gs_hi[sl], gs_lo[sl])[:1]        # [1,256]
        h8p_hi = stream_rows(gs_hi[sl], 0)[:n]                            # primal rows of the output layer's input
        h8p_lo = stream_rows(gs_lo[sl], 0)[:n]
        fb_hi, fb_lo = split16(feat_bar)
        gw[1][L - 1] = torch.cat([g_row0, tn_planes(fb_hi, fb_lo, h8p_hi, h8p_lo)], 0)
        gb[1][L - 1] = torch.cat([sdf_bar.sum(0), feat_bar.sum(0)], 0)
        del zs_hi, zs_lo, a0_hi, a0_lo

        # ============================================================== deformation network
        if use_deform:
            jbar = jac_bar + jbar_color
            adj_d = torch.zeros(p_g, 4, 4, device=dev)
            s_d = _pow2_scale(xc_bar, jbar)
            adj_d[:n, 0, :3] = xc_bar * s_d
            adj_d[:n, 1:, :3] = jbar.permute(0, 2, 1) * s_d              # tangent row j carries d/d(dx_c_i/dx_j)
            zd_hi, zd_lo = run_reverse(0, adj_d, None, gs_hi, gs_lo, g_rows)
            Wd = ws[0]
            out_dims = [w.shape[0] for w in Wd]
            ex = freq_enc(x, cfg.multires_deform_pos)
            etx = freq_enc_tangent(x, cfg.multires_deform_pos)
            tt = freq_enc(t.reshape(-1, 1), cfg.multires_deform_time)
            a0 = torch.cat([torch.cat([ex, tt], -1)[:, None, :],
                            torch.cat([etx, torch.zeros(n, 3, tt.shape[1], device=dev)], -1)], 1)  # [n,4,52]
            a0_hi, a0_lo = padded_planes(a0, g_rows)
            for m in range(0, L - 1):
                g_in = tn_planes(zd_hi[m], zd_lo[m], a0_hi, a0_lo)[:, :a0.shape[-1]] if (m == 0 or m == skip) else None
                if m == 0:
                    g = g_in
                else:
                    g = tn_planes(zd_hi[m], zd_lo[m], gs_hi[m], gs_lo[m])
                    if m == skip:
                        g = torch.cat([g[:, :out_dims[m - 1]], g_in], 1) / SQRT2
                gw[0][m] = (g / s_d)[:out_dims[m]]
                gb[0][m] = (rowsum_planes(zd_hi[m], zd_lo[m], prim_g) / s_d)[:out_dims[m]]
            o_rows = torch.cat([xc_bar[:, None, :], jbar.permute(0, 2, 1)], 1)   # [n,4,3]
            or_hi, or_lo = padded_planes(o_rows, g_rows)
            gw[0][L - 1] = tn_planes(or_hi, or_lo, gs_hi[L - 1], gs_lo[L - 1])[:3]
            gb[0][L - 1] = xc_bar.sum(0)

        grads: List[torch.Tensor] = []
        for net in nets:
            grads += gw[net] + gb[net]
        return (None, None, None, None, None, *grads)


def effective_weights(model) -> List[torch.Tensor]:
    """[W_0..W_{L-1}, b_0..b_{L-1}] per network, differentiable w.r.t. weight_g / weight_v / bias."""
    out = []
    nets = ([model.deform_network] if model.use_deform else []) + [model.sdf_network, model.color_network]
    for net in nets:
        out += [l.effective_weight() for l in net.net]
        out += [l.bias for l in net.net]
    return out


class _CumprodPos(torch.autograd.Function):
    """torch.cumprod along the last dim for strictly positive inputs (here 1 - alpha + 1e-7 >= 1e-7).  Same forward
    values; the backward is the closed form  d/dx_j = sum_{k>=j} g_k y_k / x_j  without torch's data-dependent
    zero check (a host sync, which also makes the step impossible to capture in a CUDA graph)."""

    @staticmethod
    def forward(ctx, x):
        y = torch.cumprod(x, -1)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        return torch.flip(torch.cumsum(torch.flip(g * y, (-1,)), -1), (-1,)) / x


def composite(sdf, g_o, rgb, rays_d, pts, z_vals, sample_dist, inv_s, cos_ratio):
    """NeuS alpha compositing of render_core (endosurf.py:168-203) in differentiable PyTorch (training path only;
    inference uses the CUDA composite kernel).  sdf [R,M], g_o [R,M,3], rgb [R,M,3]."""
    R, M = z_vals.shape
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full((R, 1), sample_dist, device=z_vals.device)], -1)
    mid_z = z_vals + dists * 0.5
    true_cos = (rays_d[:, None, :] * g_o).sum(-1)
    iter_cos = -(torch.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_ratio) + torch.relu(-true_cos) * cos_ratio)
    prev_cdf = torch.sigmoid((sdf - iter_cos * dists * 0.5) * inv_s)
    next_cdf = torch.sigmoid((sdf + iter_cos * dists * 0.5) * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-6) / (prev_cdf + 1e-6)).clip(0.0, 1.0)
    ones = torch.ones(R, 1, device=z_vals.device)
    weights = alpha * _CumprodPos.apply(torch.cat([ones, 1.0 - alpha + 1e-7], -1))[:, :-1]
    relax = (torch.linalg.norm(pts, dim=-1) < 1.2).to(sdf.dtype).detach()
    g_err = (torch.linalg.norm(g_o, dim=-1) - 1.0) ** 2
    g_err = (relax * g_err).sum() / (relax.sum() + 1e-6)
    return {
        "color_map": (rgb * weights[:, :, None]).sum(1),
        "depth_map": (weights * mid_z).sum(-1, keepdim=True),
        "gradients_o": g_o,
        "gradient_o_error": g_err,
        "weights": weights,
        "cdf": prev_cdf,
    }


class GraphedTrainStep:
    """Forward + loss + backward of one training step captured in a CUDA graph.

    The eager step is ~2500 launches (fused chains, library GEMMs and many small PyTorch ops on [R,M] tensors); replaying
    them as one graph removes the launch gaps and the Python/dispatcher time.  Weight packing runs inside the graph, so
    a replay always uses the current parameters.  ``iter_step`` (the cosine-anneal ratio, a host scalar in the
    reference too, endosurf.py:215-219) is fixed at capture time: re-capture when it changes (it is constant after
    ``anneal_end``).  Ray / target tensors are copied into static buffers; gradients land in the parameters' ``.grad``.

        step = GraphedTrainStep(renderer, loss_fn, rays, (color_gt, depth_gt), iter_step)
        out, loss = step(rays, color_gt, depth_gt); optimizer.step()
    """

    def __init__(self, renderer, loss_fn, rays, targets, iter_step, warmup=2):
        self.renderer, self.loss_fn, self.iter_step = renderer, loss_fn, iter_step
        self.rays = rays.detach().clone()
        self.targets = [t.detach().clone() for t in targets]
        self.params = [p for p in renderer.parameters() if p.requires_grad]
        renderer.profile(False)  # event records are not part of the graph
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream: workspace growth, cuBLAS handles, autotuning
            for _ in range(warmup):
                self._zero()
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._zero()
        renderer._packed_version = None  # the sampling chains' weight packing must be part of the graph
        self.graph = torch.cuda.CUDAGraph()
        # "thread_local": the backward runs on autograd's device thread, whose bookkeeping calls (allocator growth,
        # cuBLAS workspace queries) must not invalidate the capture; everything it enqueues on the capture stream is
        # still recorded (tests/test_gpu_training.py::test_graphed_train_step checks replay == eager).
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            out, loss = self._run()
        self.out = {k: v.detach() for k, v in out.items()}
        self.loss = loss.detach()

    def _zero(self):
        for p in self.params:
            p.grad = None

    def _run(self):
        out = self.renderer(self.rays, iter_step=self.iter_step)
        loss = self.loss_fn(out, *self.targets)
        loss.backward()
        return out, loss

    def __call__(self, rays, *targets):
        self.rays.copy_(rays, non_blocking=True)
        for s, t in zip(self.targets, targets):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.out, self.loss
