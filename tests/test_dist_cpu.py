"""N>1 host logic on CPU (gloo, world size 2): ray sharding and the single gradient all-reduce of bench.py."""
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


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    # identical replicas, different ray shards -> different local gradients
    params = [torch.nn.Parameter(torch.ones(5, 3)), torch.nn.Parameter(torch.ones(7)), torch.nn.Parameter(torch.ones(()))]
    rays = bench.make_rays(8, frame=(rank * 17) % 60, seed=rank)
    loss = sum((p * (rays[:, 3:6].sum() + rank + 1)).sum() for p in params)
    loss.backward()
    local = [p.grad.clone() for p in params]
    bench.allreduce_gradients(params, world)
    gathered = [torch.zeros_like(torch.cat([g.reshape(-1) for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
    mean = torch.stack(gathered).mean(0)
    got = torch.cat([p.grad.reshape(-1) for p in params])
    out[rank] = (torch.allclose(got, mean, atol=1e-6), rays[0, 8].item(), rays[:, 3:6].sum().item())
    dist.destroy_process_group()


def test_gradient_allreduce_and_ray_sharding_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r][0] for r in range(world)), "all-reduced gradient != mean of the local gradients"
    assert out[0][1] != out[1][1] or out[0][2] != out[1][2], "ranks must render different ray shards"


def test_reference_arm_is_rank0_only(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("RANK", "1")
    class A:  # noqa
        steps, warmup, ref_rays, rays, gpus, mode = 1, 1, 2, 4096, 2, "forward"
    bench.run_reference(A)
    assert capsys.readouterr().out == ""
