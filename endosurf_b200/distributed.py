"""Data-parallel training over ray shards (BASELINE configs[3]: 65,536 rays/step on 8 GPUs).

Rays are independent in ``render_rays``; one process per GPU renders its own shard of the batch.  The only cross-ray
reductions of a training step are the scalar sums of the masked-mean losses (``color_mask.sum()``,
``(valid_depth_region * mask).sum()`` at reference ``trainer_endosurf.py:135,149`` and the eikonal normaliser
``relax_inside_sphere.sum()`` at ``endosurf.py:202-203``).  To obtain EXACTLY the gradient of the single-GPU batch that
is the union of the shards,

    loss = sum_k w_k * (sum_ranks num_k) / (sum_ranks den_k),

every rank needs the global denominators before it starts its backward, and the global numerators only for logging:

1. forward on the local shard -> per-term ``(weight, numerator, denominator)``;
2. all-reduce of the K denominators (K <= 8 floats; they depend on the batch and on the no-grad sampling only);
3. backward of ``sum_k w_k num_k / den_k^global`` - gradients accumulate IN PLACE into one flat fp32 bucket, because
   every ``p.grad`` is a view into it (no ``torch.cat``, no copy-back);
4. ONE all-reduce (NCCL over NVLink / NVSwitch) of the bucket, whose tail carries the K local numerators.

After step 4 every rank holds the identical global gradient and applies the identical optimizer step.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Term = Tuple[float, torch.Tensor, torch.Tensor]  # (weight, numerator (differentiable), denominator (no grad))


class FlatGradBucket:
    """One flat fp32 buffer holding every parameter gradient plus ``n_extra`` trailing scalars."""

    def __init__(self, params: Sequence[torch.nn.Parameter], n_extra: int = 8):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n + n_extra, device=dev, dtype=torch.float32)
        self.n_grad, self.n_extra = n, n_extra
        self.views: List[torch.Tensor] = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.attach()

    def attach(self):
        """(Re-)install the views as ``p.grad`` (``optimizer.zero_grad(set_to_none=True)`` drops them)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()
        self.attach()

    def bind(self, renderer):
        """Let ``renderer``'s backward calls add their gradients straight into this buffer (one launch per call) instead
        of returning 82 tensors for autograd to accumulate one by one.  Only for steps that call ``loss.backward()``
        (as ``dp_backward`` does) - ``torch.autograd.grad`` would not see these gradients."""
        renderer._grad_sink = self
        return self

    def accumulate(self, params, flat_grads: torch.Tensor, extra) -> bool:
        """flat_grads: the gradients of ``params`` (in that order) as one flat tensor; extra: [(parameter, gradient)].
        Returns False (nothing done) unless the buffer is laid out in exactly this order and attached."""
        n = len(params)
        mine = self.params
        if len(mine) < n or any(a is not b for a, b in zip(mine[:n], params)):
            return False
        if any(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(mine, self.views)):
            return False
        index = {id(p): i for i, p in enumerate(mine)}
        if any(id(p) not in index for p, _ in extra):
            return False
        with torch.no_grad():
            self.flat[:flat_grads.numel()].add_(flat_grads)
            for p, g in extra:
                self.views[index[id(p)]].add_(g)
        return True

    @property
    def extra(self) -> torch.Tensor:
        return self.flat[self.n_grad:]


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_rays(n_rays: int, world: int, rank: int) -> slice:
    """Rank r renders rays [r R/G, (r+1) R/G)."""
    per = (n_rays + world - 1) // world
    return slice(min(rank * per, n_rays), min((rank + 1) * per, n_rays))


def dp_backward(bucket: FlatGradBucket, terms: Dict[str, Term], eps: Dict[str, float], group=None,
                extra_loss: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Steps 2-4 of the module docstring.  ``terms[name] = (weight, num_local, den_local)``, ``eps[name]`` is the
    constant the reference adds to that denominator.  ``extra_loss``: local loss terms that are plain per-rank means
    of equal-sized shards (averaged over ranks).  Returns the GLOBAL value of every term and of the total loss
    (detached, for logging); gradients are left, already all-reduced, in ``p.grad``."""
    world, _ = world_info(group)
    names = list(terms)
    if len(names) > bucket.n_extra:
        raise ValueError("more loss terms than extra bucket slots")
    dens = torch.stack([terms[k][2].detach().float().reshape(()) for k in names])
    if world > 1:
        dist.all_reduce(dens, op=dist.ReduceOp.SUM, group=group)
    dens = dens + torch.tensor([eps[k] for k in names], device=dens.device)
    local = sum(terms[k][0] * terms[k][1] / dens[i] for i, k in enumerate(names))
    if extra_loss is not None:
        local = local + extra_loss / world
    bucket.zero()
    local.backward()
    with torch.no_grad():
        bucket.extra.zero_()
        for i, k in enumerate(names):
            bucket.extra[i] = terms[k][1].detach()
        if world > 1:
            dist.all_reduce(bucket.flat, op=dist.ReduceOp.SUM, group=group)
        out = {k: bucket.extra[i] / dens[i] for i, k in enumerate(names)}
        out["loss"] = sum(terms[k][0] * out[k] for k in names)
    return out


def render_loss_terms(renderer, out: Dict[str, torch.Tensor], color_gt, depth_gt, color_mask, depth_mask,
                      color_w=1.0, depth_w=1.0, eik_w=0.1):
    """The render_rays part of the reference loss (trainer_endosurf.py:131-152) as (weight, numerator, denominator)
    terms.  The renderer's eikonal term arrives normalised by the LOCAL ``sum(relax) + 1e-6``; its numerator is
    recovered with the denominator the forward kept."""
    ce = (out["color_map"] - color_gt) * color_mask
    de = (out["depth_map"] - depth_gt) * depth_mask
    eik_den = renderer.last_eikonal_den.detach().reshape(())
    terms = {
        "color": (color_w, ce.abs().sum(), color_mask.sum().detach()),
        "depth": (depth_w, de.abs().sum(), depth_mask.sum().detach()),
        "eikonal": (eik_w, out["gradient_o_error"] * eik_den, eik_den - 1e-6),
    }
    eps = {"color": 1e-10, "depth": 1e-10, "eikonal": 1e-6}
    return terms, eps
