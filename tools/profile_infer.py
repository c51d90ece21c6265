"""One 800x800 inference frame for a profiler: trains the synthetic Lego scene for a few hundred steps (so the occupancy
grid and the weights are realistic), renders two warm-up frames, then brackets ONE frame with cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/infer_launches.csv \
        python tools/profile_infer.py [train_steps] [--unfused]
Without a profiler it prints the frame time (CUDA events) and the sample count."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import synth  # noqa: E402
from ngp_pl_b200.models.networks import NGP  # noqa: E402
from ngp_pl_b200.models.rendering import render  # noqa: E402
from ngp_pl_b200.trainer import Trainer  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 600
    fused = "--unfused" not in sys.argv
    graph = "--no-graph" not in sys.argv
    scene = synth.lego_scene(0)
    bank = synth.RayBank(scene, n_images=100, device="cuda", seed=0)
    model = NGP(scene.scale).cuda()
    tr = Trainer(model, n_rays=8192, lr=1e-2)
    tr.attach_bank(bank)
    tr.capture(sample=True)
    for _ in range(steps):
        tr.train_step()
    torch.cuda.synchronize()
    dirs = synth.ray_directions(synth.intrinsics(), "cuda")
    poses = torch.as_tensor(synth.camera_poses(4, seed=1234)).cuda()
    for i in range(2):
        o, d = synth.get_rays(dirs, poses[i])
        render(model, o, d, test_time=True, fused=fused, graph=graph)
    o, d = synth.get_rays(dirs, poses[2])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    a.record()
    res = render(model, o, d, test_time=True, fused=fused, graph=graph)
    b.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("frame %.3f ms, %d samples (%.2f per ray), fused=%s graph=%s graph_ok=%s" % (
        a.elapsed_time(b), int(res["total_samples"]), int(res["total_samples"]) / o.shape[0], fused, graph,
        getattr(model, "_infer_graph_ok", True)))
    # a few more frames, wall clock like bench.py's render_fps
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(12):
        o, d = synth.get_rays(dirs, poses[i % 4])
        render(model, o, d, test_time=True, fused=fused, graph=graph)
    torch.cuda.synchronize()
    print("12 frames: %.3f ms/frame wall" % ((time.perf_counter() - t0) / 12 * 1e3))


if __name__ == "__main__":
    main()
