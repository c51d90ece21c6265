"""Times the step graphs of the Trainer and their pieces alone and overlapped (CUDA events, graph replays)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import synth, _lib  # noqa: E402
from ngp_pl_b200.models.networks import NGP  # noqa: E402
from ngp_pl_b200.trainer import Trainer  # noqa: E402

scene = synth.lego_scene(0)
bank = synth.RayBank(scene, n_images=100, device="cuda")
model = NGP(0.5).cuda()
tr = Trainer(model, n_rays=8192)
tr.attach_bank(bank)
tr.capture(sample=True)
for _ in range(1000):
    tr.train_step()
torch.cuda.synchronize()
print("stats", tr.stats())
main = torch.cuda.current_stream()
side = torch.cuda.Stream()


def t(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def piece(fn):
    s = torch.cuda.Stream()
    s.wait_stream(main)
    with torch.cuda.stream(s):
        fn()
    main.wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    if fn in (tr.sample_batch, noise):
        g.register_generator_state(tr.gen)
    with torch.cuda.graph(g):
        fn()
    return g.replay


def noise():
    tr.noise.uniform_(0, 1, generator=tr.gen)


def march_only():
    _lib.check(_lib.lib().ngp_render_train_march(C.byref(tr.cfg), C.byref(tr.buf), tr._st()), "march")


def loss_only():
    tr.scalars[2:4].zero_()
    _lib.check(_lib.lib().ngp_nerf_loss_grad(C.byref(tr.cfg), C.byref(tr.buf), tr.rgb_gt.data_ptr(), tr.dL_drgb.data_ptr(),
                                             tr.dL_dopacity.data_ptr(), tr._st()), "loss")


def bwd_only():
    _lib.check(_lib.lib().ngp_render_train_bwd(C.byref(tr.net), C.byref(tr.cfg), C.byref(tr.buf), tr.dL_drgb.data_ptr(),
                                               tr.dL_dopacity.data_ptr(), None, None, tr.G.data_ptr(),
                                               tr.G[tr.n_enc:].data_ptr(), tr._st()), "bwd")


def in_context(fns, n=40):
    """per-piece time when the pieces run back to back in a loop (events around each piece)"""
    for _ in range(5):
        for f in fns:
            f()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(fns) + 1)] for _ in range(n)]
    for i in range(n):
        ev[i][0].record()
        for j, f in enumerate(fns):
            f()
            ev[i][j + 1].record()
    torch.cuda.synchronize()
    return [sum(ev[i][j].elapsed_time(ev[i][j + 1]) for i in range(n)) / n * 1e3 for j in range(len(fns))]


cur = tr._cur
fwd_p, loss_p, bwd_p = piece(tr.network), piece(loss_only), piece(bwd_only)
C_ = tr.g_compute[cur][0].replay


def seq3():
    fwd_p(); loss_p(); bwd_p()


print("V2 t(C)                     ", t(C_, 200))
print("V1 t(fwd;loss;bwd pieces)   ", t(seq3, 200))
print("V5 t(fwd) t(loss) t(bwd)    ", t(fwd_p, 200), t(loss_p, 200), t(bwd_p, 200))
print("V3 in_context([C])          ", in_context([C_], 100))
print("V4 in_context([fwd,loss,bwd])", in_context([fwd_p, loss_p, bwd_p], 100))
print("V6 in_context([fwd,bwd])    ", in_context([fwd_p, bwd_p], 100))
print("V7 in_context([C,U])        ", in_context([C_, tr.g_update[0].replay], 100))
print("V8 t(C;U)                   ", t(lambda: (C_(), tr.g_update[0].replay()), 200))
print("stats", tr.stats())
