"""Can the optimiser run UNDER the hash-gradient scatter? Times, on the buffers of a trained step (CUDA graphs, events):
the scatter alone, Adam alone, Adam on a quarter of the table alone, and scatter || Adam on two streams.
The scatter is bound by L2 atomics, Adam by HBM streaming; if they overlap well a level-chunked scatter -> Adam pipeline
would take ~max instead of the sum off the step's critical path."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import _lib, synth  # noqa: E402
from ngp_pl_b200.models.networks import NGP  # noqa: E402
from ngp_pl_b200.trainer import Trainer  # noqa: E402

scene = synth.lego_scene(0)
bank = synth.RayBank(scene, n_images=100, device="cuda")
model = NGP(scene.scale).cuda()
tr = Trainer(model, n_rays=8192)
tr.attach_bank(bank)
tr.capture(sample=True)
for _ in range(1000):
    tr.train_step()
torch.cuda.synchronize()
L = _lib.lib()
smp = _lib.NgpSamples()
smp.rays_o, smp.rays_d = tr.rays_o.data_ptr(), tr.rays_d.data_ptr()
smp.ray_idx, smp.ts = tr.ray_idx.data_ptr(), tr.ts.data_ptr()
smp.n = tr.capacity
smp.n_dev = tr.counters.data_ptr()
smp.live_idx = tr.live_idx.data_ptr()
smp.n_live_dev = tr.counters[5:].data_ptr()
G2 = torch.zeros_like(tr.G)  # a second gradient buffer: Adam reads/clears G2 while the scatter reduces into G
P2, M2, V2, Ph2 = tr.P.clone(), tr.M.clone(), tr.V.clone(), tr.Ph.clone()
G2.normal_(0, 1e-4)
main = torch.cuda.current_stream()
side = torch.cuda.Stream()


def scatter(stream):
    _lib.check(L.ngp_net_backward_scatter(C.byref(tr.net), C.byref(smp), tr.scalars[1:].data_ptr(), tr.G.data_ptr(),
                                          tr.bwd_ws.data_ptr(), tr.bwd_ws.numel(), stream.cuda_stream), "scatter")


def adam(stream, lo=0, n=None):
    n = tr.n_params - lo if n is None else n
    _lib.check(L.ngp_adam_step(P2[lo:].data_ptr(), G2[lo:].data_ptr(), M2[lo:].data_ptr(), V2[lo:].data_ptr(), Ph2[lo:].data_ptr(), n,
                               tr.lr_dev.data_ptr(), tr.step_dev.data_ptr(), 0.9, 0.999, 1e-15, 1.0, 0, stream.cuda_stream), "adam")


def graph(fn):
    s = torch.cuda.Stream()
    s.wait_stream(main)
    with torch.cuda.stream(s):
        fn(s, torch.cuda.Stream())
    main.wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        fn(cur, side)
    return g


def t(g, n=40):
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        G2.normal_(0, 1e-4)  # keep Adam's work realistic: a non-zero gradient everywhere (outside the timed graphs' kernels but inside the window: subtract)
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def both(cur, sd):
    sd.wait_stream(cur)
    with torch.cuda.stream(sd):
        adam(sd)
    scatter(cur)
    cur.wait_stream(sd)


def both_quarters(cur, sd):
    # scatter as it is, Adam as four quarter launches on the side stream (what a level-chunked pipeline would issue)
    sd.wait_stream(cur)
    q = (tr.n_params // 4) & ~3
    with torch.cuda.stream(sd):
        for k in range(4):
            adam(sd, k * q, q if k < 3 else tr.n_params - 3 * q)
    scatter(cur)
    cur.wait_stream(sd)


g_fill = graph(lambda c, s: None)
base = t(g_fill)
res = {}
for name, fn in (("scatter alone", lambda c, s: scatter(c)), ("adam alone", lambda c, s: adam(c)),
                 ("adam quarter alone", lambda c, s: adam(c, 0, (tr.n_params // 4) & ~3)),
                 ("scatter then adam (one stream)", lambda c, s: (scatter(c), adam(c))),
                 ("scatter || adam (two streams)", both), ("scatter || 4 adam quarters", both_quarters)):
    res[name] = t(graph(fn)) - base
    print("%-34s %7.1f us" % (name, res[name]), flush=True)
print("live samples %d of %d" % (int(tr.counters[5]), int(tr.counters[2])))
