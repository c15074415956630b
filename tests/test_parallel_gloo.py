"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path (pcnerf_b200/parallel.py) -- gradient
bucket all-reduce, the 1/world scaling rule of the child depth loss, ray / candidate-group sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pcnerf_oracle as orc
from pcnerf_b200 import parallel, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # ---- 1. GradBucket: flat buffer aliases .grad, all-reduce gives the rank mean
        lin = torch.nn.Linear(7, 5)
        bucket = parallel.GradBucket(list(lin.parameters()))
        x = torch.full((3, 7), float(rank + 1))
        bucket.zero()
        lin(x).sum().backward()
        local = [p.grad.clone() for p in lin.parameters()]
        assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in lin.parameters())
        bucket.allreduce_mean()
        gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
        for g, out in zip(local, gathered):
            dist.all_gather(out, g)
        for p, out in zip(lin.parameters(), gathered):
            assert torch.allclose(p.grad, sum(out) / world, atol=1e-6)
        # ---- 2. loss scaling: head-only losses of one batch of 2N rays == rank mean of the shard losses, with the
        #         child depth loss additionally divided by the world size (nof/render.py:155 carries 1/N)
        N, S = 64, 32
        rays = torch.from_numpy(synth.synth_train_rays(3, N * world, K=8))
        gen = torch.Generator().manual_seed(5)
        z = orc.sample_z(rays, S, 1, 0.1, 0, None)
        p = torch.sigmoid(torch.randn(N * world, S, generator=gen))
        fl_g, dl_g, depth_g, _ = orc.train_head(p, z, rays, None, 0.0, 1e-10, 1)
        range_g = orc.smooth_l1_mean(10 * depth_g, 10 * rays[:, 14])
        a, b = parallel.shard_rows(N * world, world, rank)
        assert (a, b) == (rank * N, (rank + 1) * N)
        fl, dl, depth, _ = orc.train_head(p[a:b], z[a:b], rays[a:b], None, 0.0, 1e-10, 1)
        terms = torch.stack([fl, dl * parallel.depth_loss_scale(), orc.smooth_l1_mean(10 * depth, 10 * rays[a:b, 14])])
        dist.all_reduce(terms)
        terms /= world
        np.testing.assert_allclose(terms.numpy(), torch.stack([fl_g, dl_g, range_g]).numpy(), rtol=1e-5)
        # ---- 3. candidate groups are never split
        rows, other, _ = synth.synth_infer_rows(9, 37)
        s0, s1 = parallel.shard_groups(torch.from_numpy(other), world, rank)
        bounds = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(bounds, torch.tensor([s0, s1]))
        assert bounds[0][0] == 0 and bounds[-1][1] == rows.shape[0]
        for i in range(world - 1):
            assert bounds[i][1] == bounds[i + 1][0]
        if s0 < rows.shape[0]:
            assert rows[s0, 12] >= 0           # a shard starts at a group head (followers carry -1)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_rows_covers_everything():
    for n in (0, 1, 7, 131072):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard_rows(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1
