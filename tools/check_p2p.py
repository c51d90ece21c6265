"""torchrun --nproc-per-node N tools/check_p2p.py [p2p|nvls|p2p_host] : the fused NVLink optimiser step
(reduce-scatter + sharded Adam + all-gather in one kernel; `p2p` / `nvls` = ngp_adam_step_fused with its in-kernel
barriers and the alternating gradient buffers, `p2p_host` = round 1's ngp_adam_step_p2p between host-launched barriers)
against NCCL all_reduce + the full-size ngp_adam_step, from identical parameters / moments / per-rank gradients.
Deterministic (no atomics involved), so at N=2 the results must be bitwise equal (a+b == b+a); for N>2 only the fp32
summation order differs. Part 1: single steps, checked one by one. Part 2: 40 steps back to back with NO host
synchronisation in between (stresses the in-kernel barriers and the buffer alternation), checked at the end."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import _lib  # noqa: E402
from ngp_pl_b200.models.networks import NGP  # noqa: E402
from ngp_pl_b200.trainer import Trainer  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    torch.manual_seed(0)
    model = NGP(0.5).cuda()
    try:
        tr = Trainer(model, n_rays=256, process_group=dist.group.WORLD, world_size=world, rank=rank, seed=rank, ddp=mode)
    except RuntimeError as e:
        if mode == "nvls" and "multicast" in str(e):
            print("rank %d: SKIP nvls (%s)" % (rank, e), flush=True)
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(0)
        raise
    n = tr.n_params
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    # N = 2: a + b == b + a, bitwise. N > 2: the fp32 summation order differs from NCCL's (a few 1e-7 on parameters of
    # magnitude <= 1), and a parameter that lands on the other side of an fp16 rounding boundary moves the fp16 working
    # copy by one fp16 ulp (<= 2^-10 relative)
    tol = 0.0 if world == 2 else 2e-6
    tol_half = 0.0 if world == 2 else 2.0 ** -10
    ok = True

    def nccl_adam(Pa, Ga, Ma, Va, Pha, step_a):
        dist.all_reduce(Ga)
        _lib.check(L.ngp_adam_step(Pa.data_ptr(), Ga.data_ptr(), Ma.data_ptr(), Va.data_ptr(), Pha.data_ptr(), n,
                                   tr.lr_dev.data_ptr(), step_a.data_ptr(), 0.9, 0.999, 1e-15, 1.0 / world, 1, st), "adam")

    def compare(tag, Pa, Ma, Pha, step_a):
        torch.cuda.synchronize()
        tr.gather_master_params()
        torch.cuda.synchronize()
        lo, hi = tr.shard_bounds()
        # this rank holds valid Adam moments only for its own shard, so path A is meaningful on [lo, hi)
        dP = (Pa[lo:hi] - tr.P[lo:hi]).abs().max().item()
        dH = (Pha[lo:hi].float() - tr.Ph[lo:hi].float()).abs().max().item()
        dM = (Ma[lo:hi] - tr.M[lo:hi]).abs().max().item()
        gz = tr.G.abs().max().item()  # the buffer the next step accumulates into must be clear
        # every rank must end up with the same complete fp16 working copy and (after the gather) fp32 master copy
        ref_h, ref_p = tr.Ph.clone(), tr.P.clone()
        dist.broadcast(ref_h, src=0)
        dist.broadcast(ref_p, src=0)
        same = bool((ref_h == tr.Ph).all()) and bool((ref_p == tr.P).all())
        full_h = bool((tr.Ph.float() - tr.P).abs().max().item() < 1e-2 * max(1.0, tr.P.abs().max().item()))
        scale = max(1.0, Pa.abs().max().item())
        good = dP <= tol * scale and dH <= tol_half * scale and dM <= tol * scale and gz == 0.0 and int(tr.step_dev) == int(step_a) \
            and same and full_h
        print("rank %d %s [%s]: shard max|dP| %.3e |dPh| %.3e |dM| %.3e  grad cleared %s  ranks identical %s  Ph==half(P) %s  step %d -> %s"
              % (rank, tag, mode, dP, dH, dM, gz == 0.0, same, full_h, int(tr.step_dev), "OK" if good else "MISMATCH"), flush=True)
        return good

    # ---- part 1: single steps ----------------------------------------------------------------------------------
    for it in range(4):
        g = torch.Generator("cuda").manual_seed(100 * it + rank)
        grad = torch.randn(n, device="cuda", generator=g) * 10.0 ** (-2 * (it % 3))
        grad[::5] = 0
        Pa, Ma, Va = tr.P.clone(), tr.M.clone(), tr.V.clone()
        Ga = grad.clone()
        Pha = torch.empty(n, device="cuda", dtype=torch.float16)
        step_a = tr.step_dev.clone()
        nccl_adam(Pa, Ga, Ma, Va, Pha, step_a)
        tr.G.copy_(grad)
        if mode == "p2p_host":
            torch.cuda.synchronize()
            dist.barrier()
        tr.optimizer_step()  # the fused modes synchronise the ranks inside the kernel
        ok = compare("it %d" % it, Pa, Ma, Pha, step_a) and ok

    # ---- part 2: back-to-back steps, no host synchronisation ---------------------------------------------------
    # the non-owned shards of P/M/V are stale on this rank (sharded optimiser): path A therefore runs on clones taken now
    # and is only compared on the owned shard
    Pa, Ma, Va = tr.P.clone(), tr.M.clone(), tr.V.clone()
    Pha = torch.empty(n, device="cuda", dtype=torch.float16)
    step_a = tr.step_dev.clone()
    gen = torch.Generator("cuda").manual_seed(777 + rank)
    grads = [torch.randn(n, device="cuda", generator=gen) * 1e-3 for _ in range(4)]
    n_steps = 40
    for k in range(n_steps):
        Ga = grads[k % 4].clone()
        nccl_adam(Pa, Ga, Ma, Va, Pha, step_a)
    torch.cuda.synchronize()
    dist.barrier()
    for k in range(n_steps):
        tr.G.copy_(grads[k % 4])
        tr.optimizer_step()
    ok = compare("%d steps back to back" % n_steps, Pa, Ma, Pha, step_a) and ok
    if mode in ("p2p", "nvls"):
        tr.check_exchange()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
