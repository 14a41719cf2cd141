"""N>1 host logic on CPU (gloo, world size 2): endosurf_b200.distributed gives EXACTLY the single-process gradient of
the union batch - masked-mean numerators / denominators are summed over ranks, gradients land in one flat bucket and
are all-reduced once - and the bench shards rays per rank."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(9, 16), torch.nn.Tanh(), torch.nn.Linear(16, 5))


def _toy_terms(model, rays, color_gt, depth_gt, mask):
    """A stand-in 'renderer' with the loss structure of the trainer: two masked means and one mean over a
    data-dependent subset (the eikonal term's relax mask)."""
    o = model(rays)
    ce = (o[:, :3] - color_gt) * mask
    de = (o[:, 3:4] - depth_gt) * mask
    relax = (rays[:, :1].abs() < 0.7).float()
    terms = {"color": (1.0, ce.abs().sum(), mask.sum()), "depth": (1.0, de.abs().sum(), mask.sum()),
             "eikonal": (0.1, (relax * (o[:, 4:5] - 1.0) ** 2).sum(), relax.sum())}
    eps = {"color": 1e-10, "depth": 1e-10, "eikonal": 1e-6}
    return terms, eps


def _batch(n):
    g = torch.Generator().manual_seed(5)
    return (torch.randn(n, 9, generator=g), torch.rand(n, 3, generator=g), torch.rand(n, 1, generator=g),
            (torch.rand(n, 1, generator=g) < 0.6).float())


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from endosurf_b200 import distributed as dp
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _toy_model()
    params = list(model.parameters())
    bucket = dp.FlatGradBucket(params)
    data = _batch(24)
    sl = dp.shard_rays(24, world, rank)
    terms, eps = _toy_terms(model, *[x[sl] for x in data])
    logs = dp.dp_backward(bucket, terms, eps)
    # p.grad are views of the bucket: nothing was concatenated or copied back
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views))
    out[rank] = (torch.cat([p.grad.reshape(-1) for p in params]).clone(), logs["loss"].item(), (sl.start, sl.stop))
    dist.destroy_process_group()


def test_data_parallel_step_equals_single_process_on_the_union_batch():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # single process, whole batch, the trainer's own normalisation
    model = _toy_model()
    terms, eps = _toy_terms(model, *_batch(24))
    loss = sum(w * num / (den + eps[k]) for k, (w, num, den) in terms.items())
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    for r in range(world):
        g, l, _ = out[r]
        assert torch.allclose(g, ref, rtol=1e-5, atol=1e-7), f"rank {r}: DP gradient != union-batch gradient"
        assert abs(l - loss.item()) <= 1e-5 * abs(loss.item())
    assert out[0][2] == (0, 12) and out[1][2] == (12, 24)
    # averaging per-rank means instead (what plain DDP would do) is NOT the same gradient: the masks differ per shard
    g_avg = []
    for r in range(world):
        model = _toy_model()
        sl = slice(*out[r][2])
        t, e = _toy_terms(model, *[x[sl] for x in _batch(24)])
        sum(w * num / (den + e[k]) for k, (w, num, den) in t.items()).backward()
        g_avg.append(torch.cat([p.grad.reshape(-1) for p in model.parameters()]))
    assert not torch.allclose(sum(g_avg) / world, ref, rtol=1e-3, atol=1e-6)


def _renderer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import copy
    import fake_lib
    from conftest import load_cfg
    from endosurf_b200 import EndoSurfRenderer, distributed as dp
    fake_lib.install()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = load_cfg()
    rc = copy.deepcopy(cfg["render"])
    rc.update(n_samples=8, n_importance=8)
    torch.manual_seed(0)
    r = EndoSurfRenderer(rc, cfg["net"], device="cpu")
    r.train()
    params = [p for v in r.get_train_params().values() for p in v]
    bucket = dp.FlatGradBucket(params).bind(r)
    g = torch.Generator().manual_seed(3)
    rays, cgt, dgt = torch.rand(16, 9, generator=g), torch.rand(16, 3, generator=g), torch.rand(16, 1, generator=g)
    msk = (torch.rand(16, 1, generator=g) < 0.6).float()
    sl = dp.shard_rays(16, world, rank)
    o = r(rays[sl], iter_step=1000)
    terms, eps = dp.render_loss_terms(r, o, cgt[sl], dgt[sl], msk[sl], msk[sl])
    # (the stand-in library leaves the outputs uninitialised: the loss VALUE is garbage, its graph is what is tested)
    dp.dp_backward(bucket, terms, eps)
    n_net = sum(p.numel() for p in params[:-1])
    out[rank] = (bucket.flat[:n_net].clone(), all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket.views)))
    dist.destroy_process_group()


def test_renderer_gradient_sink_under_data_parallel():
    """The renderer's own host path at world size 2 (stand-in library, tests/fake_lib.py): each rank's library backward
    adds its flat gradient (a known pattern) into the bound bucket, the parameter hub stays out of the way, ONE
    all-reduce leaves world x pattern on every rank, and p.grad are still the bucket's views."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fake_lib
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_renderer_worker, args=(world, port, out), nprocs=world, join=True)
    for rk in range(world):
        flat, views_ok = out[rk]
        assert views_ok
        assert torch.equal(flat, world * fake_lib.pattern(flat.numel())), f"rank {rk}"


def test_bench_shards_rays_per_rank():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.make_rays(8, frame=(0 * 17) % 60, seed=0)
    b = bench.make_rays(8, frame=(1 * 17) % 60, seed=1)
    assert a[0, 8].item() != b[0, 8].item() and not torch.equal(a[:, 3:6], b[:, 3:6])


def test_reference_arm_is_rank0_only(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("RANK", "1")
    class A:  # noqa
        steps, warmup, ref_rays, rays, gpus, mode, precision_terms = 1, 1, 2, 4096, 2, "forward", 3
    bench.run_reference(A)
    assert capsys.readouterr().out == ""


def test_shard_rays_partitions_any_batch():
    """Ragged and empty cases of the ray partition: for every batch size and world size the shards are
    contiguous, disjoint, in rank order and cover the batch exactly; surplus ranks get empty shards."""
    sys.path.insert(0, ROOT)
    from endosurf_b200 import distributed as dp
    for world in range(1, 9):
        for n in list(range(0, 40)) + [4096, 65536, 65537]:
            shards = [dp.shard_rays(n, world, r) for r in range(world)]
            covered = []
            for sl in shards:
                assert 0 <= sl.start <= sl.stop <= n
                covered += list(range(sl.start, sl.stop)) if n < 100 else []
            assert shards[0].start == 0 and shards[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(shards, shards[1:]))
            if n < 100:
                assert covered == list(range(n))
            sizes = [sl.stop - sl.start for sl in shards]
            assert max(sizes) == (n + world - 1) // world
