"""Per-phase timing of the multi-GPU step (run under torchrun on one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_probe.py [p2p|nvls|p2p_host|zero|nccl] [steps]

Every rank trains the synthetic Lego scene like bench.py, then times -- with CUDA events on its own stream -- the pieces of
the step in isolation (compute graph, each collective / barrier / kernel of the optimiser exchange) and the pipelined
train_step(); rank 0 prints min / mean / max over ranks. Answers: how long is the exchange really, how much of it the
pipelined step hides, how much rank skew there is.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import synth, _lib  # noqa: E402
from ngp_pl_b200.models.networks import NGP  # noqa: E402
from ngp_pl_b200.trainer import Trainer  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    scene = synth.lego_scene(0)
    bank = synth.RayBank(scene, n_images=100, device=dev, seed=rank)
    model = NGP(scene.scale).to(dev)
    try:
        tr = Trainer(model, n_rays=8192, lr=1e-2, process_group=None, world_size=world, rank=rank, seed=rank, ddp=mode)
    except RuntimeError as e:
        if rank == 0:
            print(json.dumps({"mode": mode, "world": world, "unavailable": str(e)}))
        dist.barrier()
        dist.destroy_process_group()
        return
    tr.attach_bank(bank)
    tr.capture(sample=True)
    for _ in range(steps):
        tr.train_step()
    torch.cuda.synchronize()
    dist.barrier()

    def timed(fn, n=50, sync_ranks=True):
        out = []
        for i in range(n + 5):
            if sync_ranks:
                dist.barrier()
                torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if i >= 5:
                out.append(a.elapsed_time(b) * 1e3)
        return sum(out) / len(out)

    res = {}
    cur = tr._cur
    res["compute_graph_us"] = timed(lambda: tr.g_compute[cur][tr._gcur].replay())
    res["prepare_graph_us"] = timed(tr.g_prepare[cur].replay)
    if mode in ("p2p", "nvls"):
        # the self-synchronising exchange kernel (start barrier + reduce-scatter/Adam/all-gather + clear + end barrier);
        # ranks enter together (timed() barriers first), so this is the kernel's own time
        res["fused_exchange_us"] = timed(tr._graph_update)
        res["fused_exchange_skewed_us"] = timed(tr._graph_update, sync_ranks=False)
    elif mode == "p2p_host":
        res["barrier_us"] = timed(lambda: tr.hG.barrier(channel=0))
        L = _lib.lib()

        def kernel_only():
            rc = L.ngp_adam_step_p2p(tr.world_size, tr.rank, tr.peer_G, tr.P.data_ptr(), tr.M.data_ptr(), tr.V.data_ptr(),
                                     tr.peer_Ph, tr.n_params, tr.lr_dev.data_ptr(), tr.step_dev.data_ptr(), tr.betas[0],
                                     tr.betas[1], tr.eps, 0, tr._st())
            _lib.check(rc, "adam_step_p2p")
        res["p2p_kernel_us"] = timed(kernel_only)
        res["zero_grad_us"] = timed(tr.G.zero_)
    elif mode == "zero":
        lo, hi, n_pad = tr._zero
        shard = n_pad // world
        res["reduce_scatter_us"] = timed(lambda: dist.reduce_scatter_tensor(tr.G_shard, tr.G_full))
        res["zero_grad_us"] = timed(tr.G_full.zero_)
        res["all_gather_us"] = timed(lambda: dist.all_gather_into_tensor(tr.Ph_full, tr.Ph_full[lo:lo + shard]))
    else:
        res["all_reduce_us"] = timed(lambda: dist.all_reduce(tr.G))
        res["adam_graph_us"] = timed(tr.g_update[0].replay)
    res["optimizer_step_us"] = timed(tr._graph_update)
    # the pipelined step, ranks free-running (what bench.py measures)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(300):
        tr.train_step()
    b.record()
    torch.cuda.synchronize()
    res["train_step_us"] = a.elapsed_time(b) / 300 * 1e3
    st = tr.stats()
    res["rm_samples"], res["bw_samples"] = st["rm_samples"], st["bw_samples"]

    keys = sorted(res)
    t = torch.tensor([float(res[k]) for k in keys], device=dev, dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        m = torch.stack(allt).cpu()
        out = {k: {"min": float(m[:, i].min()), "mean": float(m[:, i].mean()), "max": float(m[:, i].max())} for i, k in enumerate(keys)}
        print(json.dumps({"mode": mode, "world": world, "phases": out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
