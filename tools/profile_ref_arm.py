"""Where does the reference arm's step time go? torch.profiler over a few steady-state steps of `bench.py --impl reference`'s
loop (reference vren kernels + unmodified reference Python + a tinycudann stand-in). Prints the top ops by CUDA time and
the wall time per step with and without the profiler.   python tools/profile_ref_arm.py [fast|standin] [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import synth  # noqa: E402
from oracle import ref_env  # noqa: E402


def main():
    tcnn = sys.argv[1] if len(sys.argv) > 1 else "fast"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    ref = ref_env.load_reference(tcnn=tcnn)
    dev = torch.device("cuda", 0)
    scene = synth.lego_scene(0)
    bank = synth.RayBank(scene, n_images=100, device=dev, seed=0)
    model = ref.NGP(scale=scene.scale).to(dev)
    G = model.grid_size
    model.register_buffer("density_grid", torch.zeros(model.cascades, G ** 3, device=dev))
    gx = torch.stack(torch.meshgrid(*[torch.arange(G, dtype=torch.int32, device=dev)] * 3, indexing="ij"), -1).reshape(-1, 3)
    model.register_buffer("grid_coords", gx)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, eps=1e-15, fused=True)
    scaler = torch.amp.GradScaler("cuda")
    loss_fn = ref.losses.NeRFLoss(lambda_distortion=0)
    state = {"step": 0}

    def step():
        o, d, rgb = bank.sample(8192)
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
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(32):
        step()
    torch.cuda.synchronize()
    print("steady state: %.2f ms/step (32 steps incl. 2 occupancy refreshes)" % ((time.perf_counter() - t0) / 32 * 1e3))
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(16):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))


if __name__ == "__main__":
    main()
