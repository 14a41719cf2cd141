"""Host-side numerical model of what the fused sm_100a kernels compute (test helper).

The CUDA kernels do NOT run autograd: normals and the deformation Jacobian come from
forward-mode tangent rows carried through the same GEMMs as the primal row, and every
GEMM is a 3-term fp16 hi/lo split (A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, fp32 accumulate).  A bf16 pair was the
first design; it is kept here (mm_bf16x3) because the tests document why it was dropped (10x larger error).
This file restates exactly that arithmetic in PyTorch so that CPU tests can check the
*algorithm* (tangent formulation, split precision, skip folding, K padding) against the
golden fixtures without a GPU.  It is not used by the product path.
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


def split_bf16(x):
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def split_f16(x):
    hi = x.to(torch.float16).to(torch.float32)
    lo = (x - hi).to(torch.float16).to(torch.float32)
    return hi, lo


def mm_f16x3(a, w):
    ah, al = split_f16(a)
    wh, wl = split_f16(w)
    return ah @ wh.t() + (ah @ wl.t() + al @ wh.t())


def mm_f16x1(a, w):
    return split_f16(a)[0] @ split_f16(w)[0].t()


def mm_exact(a, w):
    return a @ w.t()


def mm_bf16x3(a, w):
    ah, al = split_bf16(a)
    wh, wl = split_bf16(w)
    return ah @ wh.t() + (ah @ wl.t() + al @ wh.t())


def mm_bf16x1(a, w):
    return split_bf16(a)[0] @ split_bf16(w)[0].t()


def enc_with_tangents(x, n_freqs):
    """Returns enc(x) [n, D*(2L+1)] and d enc / d x_j as [3?, n, D*(2L+1)] (one per input dim)."""
    n, dim = x.shape
    outs = [x]
    douts = [torch.ones_like(x)]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        s, c = torch.sin(x * f), torch.cos(x * f)
        outs += [s, c]
        douts += [f * c, -f * s]
    enc = torch.cat(outs, -1)
    dfull = torch.cat(douts, -1)  # derivative of each output wrt its own input component
    tang = []
    width = dfull.shape[1]
    comp = torch.arange(width) % dim  # blocks of `dim` -> component index
    for j in range(dim):
        tang.append(dfull * (comp == j).to(x.dtype)[None, :])
    return enc, torch.stack(tang, 0)


def mlp_tangent(layers, inp, dinp, skips, act, dact, mm):
    """layers: list of (W, b).  inp [n,K0]; dinp [T,n,K0].  Returns primal out, tangent out."""
    h, dh = inp, dinp
    n_l = len(layers)
    T = dinp.shape[0]
    for l, (w, b) in enumerate(layers):
        if l in skips:
            h = torch.cat([h, inp], -1) / SQRT2
            dh = torch.cat([dh, dinp], -1) / SQRT2
        z = mm(h, w) + b
        dz = mm(dh.reshape(-1, dh.shape[-1]), w).reshape(T, -1, w.shape[0])
        if l != n_l - 1:
            h = act(z)
            dh = dact(z)[None] * dz
        else:
            h, dh = z, dz
    return h, dh


def fold(sd, l):
    g, v = sd[f"net.{l}.weight_g"], sd[f"net.{l}.weight_v"]
    return v * (g / torch.linalg.norm(v, dim=1, keepdim=True)), sd[f"net.{l}.bias"]


def n_layers(sd):
    n = 0
    while f"net.{n}.bias" in sd:
        n += 1
    return n


def softplus100(z):
    return F.softplus(z, beta=100)


def dsoftplus100(z):
    return torch.sigmoid(100.0 * z)


def point_pipeline(ckpt, net_cfg, x, d, t, mm=mm_f16x3):
    """x [n,3], d [n,3], t [n,1] -> dict(x_c, jac, sdf, feat, g_c, g_o, d_c, rgb)."""
    use_deform = net_cfg["use_deform"]
    n = x.shape[0]
    eye = torch.eye(3, dtype=x.dtype)
    if use_deform:
        sd = ckpt["deform_network"]
        cfg = net_cfg["deform_network"]
        layers = [fold(sd, l) for l in range(n_layers(sd))]
        ex, dex = enc_with_tangents(x, cfg["enc_pos_cfg"]["multires"])
        et, _ = enc_with_tangents(t, cfg["enc_time_cfg"]["multires"])
        inp = torch.cat([ex, et], -1)
        dinp = torch.cat([dex, torch.zeros(3, n, et.shape[1])], -1)
        delta, ddelta = mlp_tangent(layers, inp, dinp, cfg["skips"], F.relu, lambda z: (z > 0).to(z.dtype), mm)
        x_c = x + delta
        # jac[n, i, j] = d x_c_i / d x_j ; ddelta[j, n, i]
        jac = eye[None] + ddelta.permute(1, 2, 0)
    else:
        x_c = x
        jac = eye[None].expand(n, 3, 3)
    sd = ckpt["sdf_network"]
    cfg = net_cfg["sdf_network"]
    layers = [fold(sd, l) for l in range(n_layers(sd))]
    ex, dex = enc_with_tangents(x_c, cfg["enc_pos_cfg"]["multires"])
    h, dh = mlp_tangent(layers, ex, dex, cfg["skips"], softplus100, dsoftplus100, mm)
    sdf, feat = h[:, :1], h[:, 1:]
    g_c = dh[:, :, 0].t()  # [n, 3]
    g_o = torch.einsum("nij,ni->nj", jac, g_c)  # J^T g_c
    d_c = torch.einsum("nij,nj->ni", jac, d)
    d_c = d_c / (torch.linalg.norm(d_c, dim=-1, keepdim=True) + 1e-10)
    sd = ckpt["color_network"]
    cfg = net_cfg["color_network"]
    layers = [fold(sd, l) for l in range(n_layers(sd))]
    ex, _ = enc_with_tangents(x_c, cfg["enc_pos_cfg"]["multires"])
    ed, _ = enc_with_tangents(d_c, cfg["enc_dir_cfg"]["multires"])
    inp = torch.cat([ex, g_c, ed, feat], -1)
    hcol = inp
    for l, (w, b) in enumerate(layers):
        if l in cfg["skips"]:
            hcol = torch.cat([hcol, inp], -1) / SQRT2
        hcol = mm(hcol, w) + b
        if l != len(layers) - 1:
            hcol = F.relu(hcol)
    rgb = torch.sigmoid(hcol)
    return dict(x_c=x_c, jac=jac, sdf=sdf, feat=feat, g_c=g_c, g_o=g_o, d_c=d_c, rgb=rgb)


def composite(sdf, g_o, rgb, rays_d, pts, z_vals, sample_dist, inv_s, cos_ratio):
    """NeuS compositing as the kernel does it; all inputs [R, M, ...]."""
    R, M = z_vals.shape
    dists = torch.cat([z_vals[:, 1:] - z_vals[:, :-1], torch.full((R, 1), sample_dist)], -1)
    mid_z = z_vals + dists * 0.5
    true_cos = (rays_d[:, None, :] * g_o).sum(-1)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_ratio) + F.relu(-true_cos) * cos_ratio)
    prev_cdf = torch.sigmoid((sdf - iter_cos * dists * 0.5) * inv_s)
    next_cdf = torch.sigmoid((sdf + iter_cos * dists * 0.5) * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-6) / (prev_cdf + 1e-6)).clip(0, 1)
    T = torch.cumprod(torch.cat([torch.ones(R, 1), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    w = alpha * T
    relax = (pts.norm(dim=-1) < 1.2).float()
    gerr = ((g_o.norm(dim=-1) - 1.0) ** 2 * relax).sum() / (relax.sum() + 1e-6)
    return dict(color_map=(rgb * w[..., None]).sum(1), depth_map=(w * mid_z).sum(-1, keepdim=True), weights=w,
                cdf=prev_cdf, gradient_o_error=gerr, mid_z=mid_z)
