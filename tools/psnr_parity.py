"""PSNR parity at matched steps (BASELINE.json: "PSNR parity +-0.1 dB"): train this repo's path and the reference arm
(reference vren kernels + unmodified reference Python + tinycudann stand-in) on the same synthetic Lego
scene with the same hyper-parameters for the same number of steps, then render the same held-out views.

    python tools/psnr_parity.py [steps] [out.json]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import synth  # noqa: E402
from ngp_pl_b200.models.networks import NGP  # noqa: E402
from ngp_pl_b200.models.rendering import render  # noqa: E402
from ngp_pl_b200.trainer import Trainer  # noqa: E402

N_RAYS = 8192


def eval_psnr(render_fn, scene, n_views=4, res=400):
    K = synth.intrinsics(W=res, H=res, fx=1111.11 * res / 800)
    dirs = synth.ray_directions(K, "cuda")
    poses = torch.as_tensor(synth.camera_poses(n_views, seed=4321)).cuda()
    out = []
    for i in range(n_views):
        o, d = synth.get_rays(dirs, poses[i])
        rgb = render_fn(o, d)["rgb"].float()
        gt = synth.trace(scene, o, d)
        out.append(-10 * float(torch.log10(((rgb - gt) ** 2).mean())))
    return float(np.mean(out)), out


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    scene = synth.lego_scene(0)
    bank = synth.RayBank(scene, n_images=100, device="cuda", seed=0)
    res = {"steps": steps, "rays_per_step": N_RAYS}

    model = NGP(scene.scale).cuda()
    tr = Trainer(model, n_rays=N_RAYS, lr=1e-2)
    tr.attach_bank(bank)
    tr.capture(sample=True)
    for _ in range(steps):
        tr.train_step()
    torch.cuda.synchronize()
    res["b200_psnr"], res["b200_psnr_views"] = eval_psnr(lambda o, d: render(model, o, d, test_time=True), scene)
    res["b200_train_psnr_last_batch"] = tr.stats()["psnr"]

    from oracle import ref_env
    if ref_env.available():
        ref = ref_env.load_reference()
        bank2 = synth.RayBank(scene, n_images=100, device="cuda", seed=0)
        m2 = ref.NGP(scale=scene.scale).cuda()
        G = m2.grid_size
        m2.register_buffer("density_grid", torch.zeros(m2.cascades, G ** 3, device="cuda"))
        gx = torch.stack(torch.meshgrid(*[torch.arange(G, dtype=torch.int32, device="cuda")] * 3, indexing="ij"), -1).reshape(-1, 3)
        m2.register_buffer("grid_coords", gx)
        opt = torch.optim.Adam(m2.parameters(), lr=1e-2, eps=1e-15)
        scaler = torch.amp.GradScaler("cuda")  # PL precision=16 (reference train.py:274)
        loss_fn = ref.losses.NeRFLoss(lambda_distortion=0)
        for step in range(steps):
            o, d, rgb = bank2.sample(N_RAYS)
            with torch.autocast("cuda", dtype=torch.float16):
                if step % 16 == 0:
                    m2.update_density_grid(0.01 * 1024 / 3 ** 0.5, warmup=step < 256)
                r = ref.render(m2, o, d)
                loss = sum(v.mean() for v in loss_fn(r, {"rgb": rgb}).values())
            opt.zero_grad(set_to_none=True)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
        torch.cuda.synchronize()
        def ref_render(o, d):
            with torch.autocast("cuda", dtype=torch.float16):
                return ref.render(m2, o, d, test_time=True)
        res["reference_psnr"], res["reference_psnr_views"] = eval_psnr(ref_render, scene)
        res["delta_db"] = res["b200_psnr"] - res["reference_psnr"]
    print(json.dumps(res))
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
