#!/usr/bin/env python
"""bench.py -- headline benchmark of the ngp_pl hot path on B200 (BASELINE.json metric: training rays/s,
plus 800x800 render FPS), one JSON line on rank 0.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference ...                     # the reference's own path, same config
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N ranks, NCCL

Workload (config.workload): BASELINE config 2 -- Lego-shaped synthetic scene (no dataset on the box),
8192 rays/step per GPU, L=16 T=2^19 hash grid, 800x800 training images, Adam lr 1e-2 eps 1e-15,
occupancy refresh every 16 steps; weak scaling (every rank draws its own 8192 rays, one gradient
all-reduce per step). A "step" = batch assembly + march + network forward + compositing + NeRFLoss +
backward + (all-reduce) + Adam, plus the occupancy refresh on its cadence. The timed steps run after
`--pretrain` untimed steps so the occupancy grid is in its steady state (the reference's 30k-step
headline is >99% steady-state steps); both arms do the same.

  value : whole-job rays/s with the image bank resident in HBM, CUDA-graph replay, CUDA events, max over ranks
  e2e   : the same step through Trainer.set_batch()/train_step() with HOST (pinned) ray batches copied
          H2D every step and the loss scalars read back D2H (and waited for) every step
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 8192
N_TRAIN_IMAGES = 100


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pretrain", type=int, default=None, help="untimed steps before the measurement")
    ap.add_argument("--fps-views", type=int, default=5)
    ap.add_argument("--no-fps", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ddp", default="auto", choices=["auto", "p2p", "nccl", "zero"],
                    help="N>1 gradient exchange. p2p: fused NVLink reduce-scatter+Adam+all-gather kernel; zero: the same "
                         "algorithm with NCCL reduce_scatter/all_gather; nccl: all-reduce + full Adam (the reference's DDP). "
                         "auto = zero for 2 GPUs, p2p beyond (measured best, profiles/r01_bench_n*.log)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe). Started
    early (nvidia-smi takes ~1 s to produce its first row); rows are time-stamped and only those inside
    [mark_begin(), mark_end()] are summarised."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=20):
        super().__init__(daemon=True)
        self.index, self.period = index, period_ms
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def wait_first_row(self, timeout=5.0):
        t = time.time()
        while not self.rows and time.time() - t < timeout:
            time.sleep(0.05)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) + 0.05]
        in_region = len(rows)
        if not rows:  # region shorter than one sampling period: take the rows closest to it
            rows = [r for (_, r) in self.rows[-3:]]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[1]) for r in rows if len(r) > 2 and num(r[1]) is not None]
        mx = [num(r[2]) for r in rows if len(r) > 2 and num(r[2]) is not None]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": in_region}


def dist_setup(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pg = dist.group.WORLD
    return world, rank, local, pg


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
def cpu_baseline_port(n_rays=16384):
    """The C oracle (a port: the reference has no CPU path) rendering a bounded sample of the same
    workload on one host core: march + network forward + compositing. Forward only -- the C oracle has no
    backward -- so this is an UPPER bound on what a CPU training step could reach."""
    import torch
    from oracle import oracle as O
    from ngp_pl_b200 import synth
    O.build()
    scene = synth.lego_scene(0)
    bits = synth.pack_bits(synth.occupancy_grid(scene))
    dirs = synth.ray_directions(synth.intrinsics())
    poses = torch.as_tensor(synth.camera_poses(N_TRAIN_IMAGES))
    rng = np.random.RandomState(0)
    img = torch.as_tensor(rng.randint(0, N_TRAIN_IMAGES, n_rays))
    pix = torch.as_tensor(rng.randint(0, dirs.shape[0], n_rays))
    o, d = synth.get_rays(dirs[pix], poses[img])
    o, d = o.numpy(), d.numpy()
    b = float(np.float32(np.exp(np.log(2048 * 0.5 / 16) / 15)))
    meta, entries = O.grid_meta(16, 19, 16, b)
    enc = rng.uniform(-0.1, 0.1, 3072 + 2 * entries).astype(np.float32)
    rgbp = rng.uniform(-0.2, 0.2, 7168).astype(np.float32)
    mn, mx = np.full((1, 3), -0.5, np.float32), np.full((1, 3), 0.5, np.float32)
    t0 = time.perf_counter()
    hits = O.ray_aabb(o, d, np.zeros(3, np.float32), np.full(3, 0.5, np.float32), 0.01)
    ra, xyzs, dd, deltas, ts = O.march_train(o, d, hits, bits, 1, 0.5, 0.0, rng.rand(n_rays).astype(np.float32), 128, 1024)
    sig, rgbs, _ = O.ngp_forward_c(meta, enc, rgbp, mn, mx, xyzs, dd)
    O.composite_train_fw(sig, rgbs, deltas, ts, ra, 1e-4)
    dt = time.perf_counter() - t0
    return {"value": n_rays / dt, "unit": "rays/s", "cores": 1, "kind": "port",
            "sample": "%d rays (%d samples) of the same workload, FORWARD render only (march + hash/MLP network + "
                      "compositing) by the C oracle on 1 host core, %.1f s; the reference has no CPU path and the oracle "
                      "has no backward, so this over-states CPU training throughput" % (n_rays, len(ts), dt)}


# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    from ngp_pl_b200 import _lib, synth
    from ngp_pl_b200.models.networks import NGP
    from ngp_pl_b200.models.rendering import render
    from ngp_pl_b200.trainer import Trainer
    world, rank, local, pg = dist_setup(args)
    dev = torch.device("cuda", local)
    scene = synth.lego_scene(0)
    bank = synth.RayBank(scene, n_images=N_TRAIN_IMAGES, device=dev, seed=rank)  # every rank: own images order/sampling
    model = NGP(scene.scale).to(dev)
    ddp_mode = args.ddp if args.ddp != "auto" else ("zero" if world <= 2 else "p2p")
    try:
        tr = Trainer(model, n_rays=N_RAYS, lr=1e-2, process_group=pg, world_size=world, rank=rank, seed=rank, ddp=ddp_mode)
    except Exception as e:  # symmetric memory unavailable: fall back to NCCL and say so in the line
        if world == 1 or ddp_mode != "p2p":
            raise
        ddp_mode = "zero (p2p unavailable: %s)" % type(e).__name__
        model = NGP(scene.scale).to(dev)
        tr = Trainer(model, n_rays=N_RAYS, lr=1e-2, process_group=pg, world_size=world, rank=rank, seed=rank, ddp="zero")
    tr.attach_bank(bank)
    pretrain = args.pretrain if args.pretrain is not None else 1000
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local)
    sampler.start()

    # CUDA graphs: [batch + march] / [network fwd + loss + bwd] / [Adam]; the next step's [batch + march] replays on
    # a side stream while this step's optimiser (Adam, or all-reduce / the fused NVLink kernel for N>1) runs
    tr.capture(sample=True)
    step = tr.train_step
    for _ in range(pretrain):
        step()
    for _ in range(W):
        step()
    barrier(world)
    sampler.wait_first_row()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    sampler.mark_begin()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier(world)
    sampler.mark_end()
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    clocks = sampler.stop()
    stats = tr.stats()
    value = world * N_RAYS * K / (ms * 1e-3)

    # ---- e2e: host batches, H2D every step, loss read back (and waited for) every step ------------------
    n_host = 32
    host = [tuple(t.cpu().pin_memory() for t in bank.sample(N_RAYS)) for _ in range(n_host)]
    out_host = torch.zeros(8, dtype=torch.float32).pin_memory()
    tr.capture(sample=False)
    step_nosample = lambda: tr.train_step(sample=False)

    def e2e_step(i):
        # step i consumes the batch staged before; the NEXT batch's host->device copy and march are enqueued (side stream)
        # before this step's result is waited for, as any prefetching loader does -- every step still pays its own H2D
        # copy and its own loss read-back inside the timed region
        step_nosample()
        tr.stage_batch(*host[(i + 1) % n_host])
        out_host.copy_(tr.scalars, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(out_host[2])
    tr.stage_batch(*host[0])
    for i in range(W):
        e2e_step(i)
    barrier(world)
    Ke = min(K, 300)
    e0.record()
    for i in range(Ke):
        e2e_step(i)
    e1.record()
    barrier(world)
    ms_e2e = max_over_ranks(e0.elapsed_time(e1), world)
    e2e = {"value": world * N_RAYS * Ke / (ms_e2e * 1e-3), "unit": "rays/s", "steps": Ke,
           "h2d_bytes_per_step": N_RAYS * 9 * 4, "d2h_bytes_per_step": 32}

    # ---- roofline of the dominant kernel (network backward), timed alone with CUDA events ------------------
    roof = None
    fps = None
    if rank == 0:
        hbm, tf, which = peaks()
        n_samples = stats["rm_samples"]
        smp = _lib.NgpSamples()
        smp.rays_o, smp.rays_d = tr.rays_o.data_ptr(), tr.rays_d.data_ptr()
        smp.ray_idx, smp.ts = tr.ray_idx.data_ptr(), tr.ts.data_ptr()
        smp.n, smp.n_dev = tr.capacity, tr.counters[2:].data_ptr()
        # the backward visits only the composited samples (live list of the last step; counters[5] = its length)
        smp_b = _lib.NgpSamples.from_buffer_copy(smp)
        n_bwd = stats["bw_samples"]
        if tr.live_idx is not None:
            smp_b.live_idx, smp_b.n_live_dev = tr.live_idx.data_ptr(), tr.counters[5:].data_ptr()
        L = _lib.lib()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def timed(fn, iters=10):
            ts_ = []
            for it in range(iters + 3):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                torch.cuda.synchronize()
                if it >= 3:
                    ts_.append(a.elapsed_time(b))
            return float(np.mean(ts_))
        st = torch.cuda.current_stream().cuda_stream
        ws = (tr.bwd_ws.data_ptr(), tr.bwd_ws.numel())
        t_mlp = timed(lambda: L.ngp_net_backward_mlp(C.byref(tr.net), C.byref(smp_b), tr.dsigmas.data_ptr(), tr.drgbs.data_ptr(),
                                                     tr.feat_save.data_ptr(), tr.scalars[1:].data_ptr(), tr.G.data_ptr(),
                                                     tr.G[tr.n_enc:].data_ptr(), ws[0], ws[1], st))
        t_sc = timed(lambda: L.ngp_net_backward_scatter(C.byref(tr.net), C.byref(smp_b), tr.scalars[1:].data_ptr(), tr.G.data_ptr(),
                                                        ws[0], ws[1], st))
        t_fwd = timed(lambda: L.ngp_net_forward(C.byref(tr.net), C.byref(smp), 1, tr.sigmas.data_ptr(), tr.rgbs.data_ptr(),
                                                None, tr.feat_save.data_ptr(), st))
        tr.G.zero_()
        traffic = {}
        prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof))
            except Exception:
                traffic = {}

        def entry(kernel, bound, ms_, n_, per_sample, peak, unit, note):
            alg = n_ * per_sample
            ach = alg / (ms_ * 1e-3) / (1e9 if unit == "GB/s" else 1e12)
            return {"kernel": kernel, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                    "traffic": traffic.get(kernel + "_dram_bytes_per_launch"), "ms_per_launch": ms_,
                    "samples_per_launch": n_, "algorithmic": note}
        # algorithmic work per sample (SURVEY.md section 8d / DESIGN.md): forward 512 B of table reads, scatter 1,024 B of
        # table-gradient read-modify-write, MLP backward 40,960 FLOP (dgrad + wgrad; the forward recompute is not counted)
        ks = [entry("k_ngp_fwd", "hbm", t_fwd, n_samples, 512.0, hbm, "GB/s", "512 B/sample table gathers, every marched sample"),
              entry("k_grid_scatter_merged", "hbm", t_sc, n_bwd, 1024.0, hbm, "GB/s",
                    "1,024 B/sample gradient RMW, composited samples only"),
              entry("k_ngp_bwd", "tensor", t_mlp, n_bwd, 40960.0, tf, "TFLOP/s",
                    "40,960 FLOP/sample dgrad+wgrad, composited samples only")]
        roof = dict(max(ks, key=lambda e: e["ms_per_launch"]))  # the dominant kernel of the step
        roof["peak_source"] = which
        roof["kernels"] = ks
        roof["note"] = ("hash table (22.9 MB fp16) and its fp32 gradient (45.8 MB) are L2-resident on B200, so DRAM traffic stays far "
                        "below the algorithmic bytes; the physical limiters are L1 wavefronts of divergent 4-B gathers / 8-B reductions "
                        "and, for the MLP backward, latency at 12% occupancy (profiles/)")
        # ---- 800x800 render FPS with the trained model (BASELINE config 3) --------------------------------
        if not args.no_fps:
            try:
                fps = render_fps(lambda o, d: render(model, o, d, test_time=True), scene, dev, args.fps_views)
            except Exception as e:  # a secondary number: never let it sink the bench line
                fps = {"unavailable": repr(e)}

    if rank != 0:
        return
    per_update = 9  # kernels of ngp_update_density_grid for one cascade
    # per step: sample_rays, march, scan, compact | fwd, composite fw + loss + composite bw, loss scale, MLP bwd, scatter |
    # adam, step_inc
    launches = K * 11 + (K // tr.update_interval + 1) * per_update
    line = {
        "metric": "train_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic (seeded Lego-shaped box scene, 100 ray-traced 800x800 training views; random-init weights "
                "pre-trained %d untimed steps)" % pretrain,
        "config": {"workload": "BASELINE config 2: Lego 800x800, 8192 rays/step/GPU, L=16 T=2^19 F=2, Adam lr 1e-2, "
                               "occupancy refresh every 16 steps", "rays_per_step_per_gpu": N_RAYS, "global_rays_per_step": world * N_RAYS,
                   "parallelism": "dp%d" % world + ("" if world == 1 else " [%s]" % (
                       {"p2p": "fused NVLink reduce-scatter+Adam+all-gather kernel",
                                                 "zero": "NCCL reduce_scatter + sharded Adam + all_gather(fp16 params)",
                                                 "nccl": "NCCL all_reduce + full Adam"}.get(ddp_mode, ddp_mode))),
                   "pretrain_steps": pretrain,
                   "l2": "no explicit flush: each step streams params+grads+Adam moments (~230 MB) > 126 MB L2",
                   "samples_per_ray_marched": stats["rm_samples"] / N_RAYS, "samples_per_ray_composited": stats["vr_samples"] / N_RAYS,
                   "samples_per_ray_in_backward": stats["bw_samples"] / N_RAYS,
                   "train_psnr_last_batch": stats["psnr"]},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof,
    }
    if fps is not None:
        line["render_fps"] = fps
    if not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline_port()
        except Exception as e:  # the oracle is only a reported baseline; never let it sink the bench line
            line["cpu_baseline"] = {"unavailable": repr(e)}
    print(json.dumps(line))


def render_fps(render_fn, scene, dev, n_views):
    """mean wall time per 800x800 image over test views, torch.cuda.synchronize() bracketed as in the
    reference's test.ipynb cell 2"""
    import torch
    from ngp_pl_b200 import synth
    K = synth.intrinsics()
    dirs = synth.ray_directions(K, dev)
    poses = torch.as_tensor(synth.camera_poses(n_views + 1, seed=1234)).to(dev)
    times, samples = [], []
    for i in range(n_views + 1):
        o, d = synth.get_rays(dirs, poses[i])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = render_fn(o, d)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i > 0:  # first view warms up allocations
            times.append(dt)
            samples.append(float(res["total_samples"]) / o.shape[0])
    gt = synth.trace(scene, o, d)
    mse = ((res["rgb"].float() - gt) ** 2).mean().item()
    return {"value": 1.0 / float(np.mean(times)), "unit": "frames/s", "resolution": "800x800", "views": n_views,
            "ms_per_frame": 1e3 * float(np.mean(times)), "samples_per_ray": float(np.mean(samples)),
            "psnr_last_view": -10 * float(np.log10(max(mse, 1e-12)))}


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own path: its compiled vren kernels + its unmodified models/{rendering,networks,
    custom_functions}.py and losses.py (oracle/_ref), tinycudann replaced by the PyTorch stand-in
    (tinycudann is unavailable), driven by a loop that mirrors NeRFSystem.training_step (train.py:159-185)
    with torch.optim.Adam(eps=1e-15) in place of apex FusedAdam. Same scene, config, pretrain and timing."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the reference arm is a 1-GPU baseline ("not required" for N>1)
    from oracle import ref_env
    if not ref_env.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference vren build) is not present on this box"}))
        return
    import torch
    torch.cuda.set_device(0)
    from ngp_pl_b200 import synth
    ref = ref_env.load_reference()
    dev = torch.device("cuda", 0)
    scene = synth.lego_scene(0)
    bank = synth.RayBank(scene, n_images=N_TRAIN_IMAGES, device=dev, seed=0)
    model = ref.NGP(scale=scene.scale).to(dev)
    G = model.grid_size
    model.register_buffer("density_grid", torch.zeros(model.cascades, G ** 3, device=dev))  # train.py:73-76
    gx = torch.stack(torch.meshgrid(*[torch.arange(G, dtype=torch.int32, device=dev)] * 3, indexing="ij"), -1).reshape(-1, 3)
    model.register_buffer("grid_coords", gx)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    loss_fn = ref.losses.NeRFLoss(lambda_distortion=0)
    state = {"step": 0, "res": None}

    def step():
        # PL runs training_step under fp16 autocast with its GradScaler (Trainer(precision=16), train.py:274): the
        # network outputs are fp16, so without loss scaling the per-sample gradients underflow
        o, d, rgb = bank.sample(N_RAYS)  # fp32 rays (the reference builds them under autocast(dtype=float32), ray_utils.py:46)
        with torch.autocast("cuda", dtype=torch.float16):
            if state["step"] % 16 == 0:
                model.update_density_grid(0.01 * 1024 / 3 ** 0.5, warmup=state["step"] < 256)
            res = ref.render(model, o, d)
            loss = sum(v.mean() for v in loss_fn(res, {"rgb": rgb}).values())
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        state["step"] += 1
        state["res"] = (res, rgb)
    pretrain = args.pretrain if args.pretrain is not None else 1000
    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(0)
    sampler.start()
    for _ in range(pretrain + W):
        step()
    torch.cuda.synchronize()
    sampler.wait_first_row()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler.mark_begin()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    res, rgb = state["res"]
    mse = ((res["rgb"].float() - rgb) ** 2).mean().item()
    value = N_RAYS * K / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "train_rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": 1, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic (same scene/bank as the b200 arm), pre-trained %d untimed steps" % pretrain,
        "config": {"workload": "BASELINE config 2 (same as the b200 arm)", "rays_per_step_per_gpu": N_RAYS,
                   "parallelism": "dp1", "pretrain_steps": pretrain,
                   "stack": "reference vren CUDA kernels (compiled from /root/reference/models/csrc) + unmodified reference "
                            "render()/NGP/custom_functions/NeRFLoss + tinycudann STAND-IN (PyTorch ops; tinycudann unavailable) "
                            "+ torch.optim.Adam (apex unavailable)",
                   "samples_per_ray_marched": float(res["rm_samples"]) / N_RAYS,
                   "samples_per_ray_composited": float(res["vr_samples"]) / N_RAYS,
                   "train_psnr_last_batch": -10 * float(np.log10(max(mse, 1e-12)))},
        "clocks": clocks,
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": 0, "kind": "reference",
                         "sample": "the reference has NO CPU path (every op TORCH_CHECKs is_cuda); this is its own GPU path on the "
                                   "same B200, all %d steps of the workload" % K},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_fps:
        line["render_fps"] = render_fps(lambda o, d: ref.render(model, o, d, test_time=True), scene, dev, args.fps_views)
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
    try:
        import torch.distributed as _d
        if _d.is_initialized():
            _d.destroy_process_group()
    except Exception:
        pass
