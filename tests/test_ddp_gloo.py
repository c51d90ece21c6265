"""Host-side logic of the N>1 path on CPU: two `gloo` processes (world_size 2, 127.0.0.1).

The data-parallel step is: every rank computes the gradient of ITS batch's mean loss into the flat
buffer, ONE all_reduce(SUM), Adam with grad_mul = 1/world. Checked here with the torch-CPU oracle as
the stand-in for the (GPU-only) backward kernel: the result must equal the gradient of the mean loss
over the union batch, which is what the reference's DDP computes. Also: view sharding for inference and
the rank-0 occupancy broadcast.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _loss_grad(oracle, meta, enc, rgbp, x, d, tgt):
    enc = enc.clone().requires_grad_(True)
    rgbp = rgbp.clone().requires_grad_(True)
    sig, rgb, _ = oracle.torch_ngp_forward(meta, enc, rgbp, torch.full((1, 3), -0.5), torch.full((1, 3), 0.5), x, d)
    loss = ((rgb - tgt) ** 2).mean() + 1e-3 * sig.mean()
    loss.backward()
    return torch.cat([enc.grad, rgbp.grad])


def _problem():
    from oracle import oracle
    rng = np.random.RandomState(0)
    L, log2_T = 2, 8
    b = float(np.float32(np.exp(np.log(2048 * 0.5 / 16) / 15)))
    meta, entries = oracle.grid_meta(L, log2_T, 16, b)
    enc = torch.as_tensor(np.concatenate([rng.uniform(-0.3, 0.3, 3072), rng.uniform(-0.5, 0.5, 2 * entries)]).astype(np.float32))
    rgbp = torch.as_tensor(rng.uniform(-0.3, 0.3, 7168).astype(np.float32))
    x = torch.as_tensor(rng.uniform(-0.5, 0.5, (64, 3)).astype(np.float32))
    d = torch.as_tensor(rng.normal(size=(64, 3)).astype(np.float32))
    tgt = torch.as_tensor(rng.rand(64, 3).astype(np.float32))
    return oracle, meta, enc, rgbp, x, d, tgt


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngp_pl_b200.trainer import allreduce_gradients, broadcast_occupancy, shard_range
    oracle, meta, enc, rgbp, x, d, tgt = _problem()
    lo, hi = shard_range(x.shape[0], world, rank)
    g = _loss_grad(oracle, meta, enc, rgbp, x[lo:hi], d[lo:hi], tgt[lo:hi])
    allreduce_gradients(g, world)
    g = g * (1.0 / world)  # what ngp_adam_step's grad_mul does
    bits = torch.full((16,), rank + 1, dtype=torch.uint8)
    broadcast_occupancy(bits, world)
    if rank == 0:
        torch.save({"g": g, "bits": bits}, out)
    else:
        assert int(bits[0]) == 1
    dist.destroy_process_group()


def test_two_rank_gradient_average_matches_global_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    oracle, meta, enc, rgbp, x, d, tgt = _problem()
    g_full = _loss_grad(oracle, meta, enc, rgbp, x, d, tgt)
    assert torch.allclose(res["g"], g_full, rtol=1e-4, atol=1e-7)
    assert int(res["bits"][0]) == 1


def test_shard_range_partitions_exactly():
    from ngp_pl_b200.trainer import shard_range
    for n in (0, 1, 7, 200, 201):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_zero_shards_partition_the_parameters():
    """"zero" mode (reduce_scatter + sharded Adam + all_gather): equal 16-byte aligned shards that cover every parameter once"""
    from ngp_pl_b200.trainer import zero_shard
    for n in (12206080 + 3072 + 7168, 1001, 4, 7):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(world):
                lo, hi, n_pad = zero_shard(n, world, r)
                assert n_pad % (4 * world) == 0 and n <= n_pad < n + 4 * world
                assert (lo % 4 == 0 or hi == lo) and 0 <= lo <= hi <= n and hi - lo <= n_pad // world
                assert lo == min(r * (n_pad // world), n)
                covered += hi - lo
            assert covered == n
