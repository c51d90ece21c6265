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


def _zero_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ngp_pl_b200.trainer import zero_exchange, zero_shard
    n = 1003  # not a multiple of 4 * world: exercises the padding
    g = torch.Generator().manual_seed(7)
    P0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) for _ in range(world)]  # every rank's local gradient (known to all, for the check)
    lo, hi, n_pad = zero_shard(n, world, rank)
    P, M, V = P0.clone(), torch.zeros(n), torch.zeros(n)
    G_full = torch.zeros(n_pad)
    G_full[:n] = grads[rank]
    G_shard = torch.zeros(n_pad // world)
    Ph_full = torch.zeros(n_pad, dtype=torch.float16)
    Ph_full[:n] = P0.half()
    lr, b1, b2, eps = 1e-2, 0.9, 0.999, 1e-15

    def adam_on_shard(lo_, hi_, g_shard):  # torch stand-in for ngp_adam_step(grad_mul = 1/world) on the owned slice
        gm = g_shard[:hi_ - lo_] / world
        M[lo_:hi_] = b1 * M[lo_:hi_] + (1 - b1) * gm
        V[lo_:hi_] = b2 * V[lo_:hi_] + (1 - b2) * gm * gm
        P[lo_:hi_] -= lr / (1 - b1) * M[lo_:hi_] / (V[lo_:hi_].sqrt() / (1 - b2) ** 0.5 + eps)
        Ph_full[lo_:hi_] = P[lo_:hi_].half()
        g_shard.zero_()
    zero_exchange(G_full, G_shard, Ph_full, (lo, hi, n_pad), rank, world, None, adam_on_shard)
    assert float(G_full.abs().max()) == 0.0  # local gradient cleared for the next step
    torch.save({"Ph": Ph_full[:n].clone(), "P_shard": P[lo:hi].clone(), "lo": lo, "hi": hi}, out % rank)
    dist.destroy_process_group()


def test_zero_mode_exchange_equals_allreduce_plus_full_adam(tmp_path):
    """reduce_scatter -> Adam on the owned shard -> all_gather(fp16 copy) gives every rank the parameters that an
    all_reduce + full Adam step gives (first step, mean gradient)"""
    world = 2
    out = str(tmp_path / "zero_r%d.pt")
    mp.spawn(_zero_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    n = 1003
    g = torch.Generator().manual_seed(7)
    P0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) for _ in range(world)]
    ref = P0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2, betas=(0.9, 0.999), eps=1e-15)
    ref.grad = sum(grads) / world
    opt.step()
    res = [torch.load(out % r) for r in range(world)]
    assert torch.equal(res[0]["Ph"], res[1]["Ph"])  # every rank holds the complete fp16 working copy
    assert torch.allclose(res[0]["Ph"].float(), ref.detach().half().float(), atol=2e-3, rtol=2e-3)
    for r in res:  # fp32 master values of the owned shard
        assert torch.allclose(r["P_shard"], ref.detach()[r["lo"]:r["hi"]], rtol=1e-5, atol=1e-6)
    assert res[0]["lo"] == 0 and res[0]["hi"] == res[1]["lo"] and res[1]["hi"] == n
