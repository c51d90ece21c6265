"""Loss curves of this repo's training step and of the reference arm on IDENTICAL inputs: the same initial weights
(both stacks initialise from the same seeds), the same host-built ray batches, the same march jitter, step by step.

    python tools/curve_parity.py [steps] [out.json]

Arms
  b200_eager   : Trainer pieces called one by one (set_batch, march, fused compute, Adam), no CUDA graph, host-fed batches
  b200_graph   : the production path (captured graphs, device-side batch sampling, pipelined march) -- different batches,
                 same distribution; answers "does the pipelined path train like the eager one"
  ref_standin  : reference vren kernels + unmodified reference Python + the CHECKER tinycudann stand-in (fp32 autograd)
  ref_fast     : the same with the performance-grade stand-in (fp16 activations / gradients, GradScaler)
Grid modes
  fixed : every arm marches the scene's analytic occupancy bitfield, no refresh  -> isolates network + optimiser numerics
  own   : every arm refreshes its own occupancy grid on the reference's cadence (every 16 steps, warm-up 256)
Output: per-arm train-batch PSNR at checkpoints + the mean over the last 50 steps, samples per ray, and at step 0 the
per-segment cosine / norm ratio of this repo's gradient against the reference arm's (same weights, same batch).
"""
import json
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import synth  # noqa: E402
from ngp_pl_b200.models.networks import NGP  # noqa: E402
from ngp_pl_b200.trainer import Trainer  # noqa: E402

N_RAYS = 8192
DEV = "cuda"


def make_batches(scene, steps, seed=0):
    bank = synth.RayBank(scene, n_images=100, device=DEV, seed=0)
    g = torch.Generator(device=DEV).manual_seed(1234 + seed)
    out = []
    for _ in range(steps):
        o, d, rgb = bank.sample(N_RAYS)
        out.append((o, d, rgb, torch.rand(N_RAYS, device=DEV, generator=g)))
    return bank, out


def psnr(mse):
    return -10 * math.log10(max(mse, 1e-12))


def analytic_bits(scene):
    return torch.as_tensor(synth.pack_bits(synth.occupancy_grid(scene))).to(DEV)


def run_b200_eager(scene, batches, grid, grads_at=()):
    model = NGP(scene.scale).to(DEV)
    tr = Trainer(model, n_rays=N_RAYS, lr=1e-2)
    if grid == "fixed":
        model.density_bitfield.copy_(analytic_bits(scene))
    curve, spr, grads = [], [], {}
    for i, (o, d, rgb, noise) in enumerate(batches):
        if grid == "own" and i % 16 == 0:
            tr.host_step = i
            tr.update_density_grid(warmup=i < 256)
        tr.set_batch(o, d, rgb)
        tr.noise.copy_(noise)
        tr.march(jitter=False)
        tr._compute()
        if i in grads_at:
            grads[i] = (tr.G.clone(), tr.P.clone())
        tr._update()
        s = tr.scalars.tolist()
        curve.append(s[2] / (3 * N_RAYS))
        spr.append(tr.counters.tolist()[2] / N_RAYS)
    return curve, spr, grads, tr


def run_b200_graph(scene, steps, grid):
    model = NGP(scene.scale).to(DEV)
    bank = synth.RayBank(scene, n_images=100, device=DEV, seed=0)
    tr = Trainer(model, n_rays=N_RAYS, lr=1e-2, update_interval=16 if grid == "own" else 1 << 30)
    tr.attach_bank(bank)
    if grid == "fixed":
        tr.host_step = 1  # never hits the refresh cadence
        model.density_bitfield.copy_(analytic_bits(scene))
    tr.capture(sample=True)
    curve, spr = [], []
    for i in range(steps):
        tr.train_step()
        s = tr.scalars.tolist()
        curve.append(s[2] / (3 * N_RAYS))
        spr.append(tr.counters.tolist()[2] / N_RAYS)
    return curve, spr


def run_ref(scene, batches, grid, tcnn, load_from=None, grads_at=()):
    from oracle import ref_env
    ref = ref_env.load_reference(tcnn=tcnn)
    model = ref.NGP(scale=scene.scale).to(DEV)
    G = model.grid_size
    model.register_buffer("density_grid", torch.zeros(model.cascades, G ** 3, device=DEV))
    gx = torch.stack(torch.meshgrid(*[torch.arange(G, dtype=torch.int32, device=DEV)] * 3, indexing="ij"), -1).reshape(-1, 3)
    model.register_buffer("grid_coords", gx)
    if grid == "fixed":
        model.density_bitfield.copy_(analytic_bits(scene))
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    loss_fn = ref.losses.NeRFLoss(lambda_distortion=0)
    curve, spr, grads = [], [], {}
    orig_rand_like = torch.rand_like
    for i, (o, d, rgb, noise) in enumerate(batches):
        torch.rand_like = lambda x, **k: noise.clone() if x.shape == noise.shape else orig_rand_like(x, **k)
        try:
            with torch.autocast("cuda", dtype=torch.float16):
                if grid == "own" and i % 16 == 0:
                    torch.rand_like = orig_rand_like
                    model.update_density_grid(0.01 * 1024 / 3 ** 0.5, warmup=i < 256)
                    torch.rand_like = lambda x, **k: noise.clone() if x.shape == noise.shape else orig_rand_like(x, **k)
                res = ref.render(model, o, d)
                loss = sum(v.mean() for v in loss_fn(res, {"rgb": rgb}).values())
        finally:
            torch.rand_like = orig_rand_like
        opt.zero_grad(set_to_none=True)
        scaler.scale(loss).backward()
        if i in grads_at:
            inv = 1.0 / scaler.get_scale()
            grads[i] = torch.cat([model.xyz_encoder.params.grad.float() * inv, model.rgb_net.params.grad.float() * inv])
        scaler.step(opt)
        scaler.update()
        curve.append(((res["rgb"].float() - rgb) ** 2).mean().item())
        spr.append(float(res["rm_samples"]) / N_RAYS)
    return curve, spr, grads


def summarise(curve, spr):
    pts = [0, 9, 49, 99, 199, 299, 499, 999, 1999]
    out = {"psnr_at": {str(p + 1): psnr(curve[p]) for p in pts if p < len(curve)},
           "psnr_mean_last50": psnr(float(np.mean(curve[-50:]))),
           "samples_per_ray_last": spr[-1], "samples_per_ray_mean": float(np.mean(spr))}
    return out


def grad_compare(g_ours, g_ref, n_enc):
    segs = {"W1d": (0, 2048), "W2d": (2048, 3072), "table": (3072, n_enc), "W1r": (n_enc, n_enc + 2048),
            "W2r": (n_enc + 2048, n_enc + 6144), "W3r": (n_enc + 6144, n_enc + 7168)}
    out = {}
    for k, (a, b) in segs.items():
        x, y = g_ours[a:b].double(), g_ref[a:b].double()
        out[k] = {"cos": float((x * y).sum() / (x.norm() * y.norm() + 1e-300)), "norm_ratio": float(x.norm() / (y.norm() + 1e-300)),
                  "max_abs_err_over_max": float((x - y).abs().max() / (y.abs().max() + 1e-300)),
                  "nonzero_ours": int((x != 0).sum()), "nonzero_ref": int((y != 0).sum())}
    return out


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    scene = synth.lego_scene(0)
    _, batches = make_batches(scene, steps)
    res = {"steps": steps, "rays_per_step": N_RAYS}
    for grid in ("fixed", "own"):
        r = {}
        c, s, g_ours, tr = run_b200_eager(scene, batches, grid, grads_at=(0,))
        r["b200_eager"] = summarise(c, s)
        n_enc = tr.n_enc
        del tr
        c, s = run_b200_graph(scene, steps, grid)
        r["b200_graph"] = summarise(c, s)
        c, s, g_ref = run_ref(scene, batches, grid, "standin", grads_at=(0,))
        r["ref_standin"] = summarise(c, s)
        r["grad_step0_vs_standin"] = grad_compare(g_ours[0][0], g_ref[0], n_enc)
        c, s, g_fast = run_ref(scene, batches, grid, "fast", grads_at=(0,))
        r["ref_fast"] = summarise(c, s)
        r["grad_step0_fast_vs_standin"] = grad_compare(g_fast[0], g_ref[0], n_enc)
        res[grid] = r
        print(grid, json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
