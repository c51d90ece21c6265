"""Burst vs sustained: times the trainer's graphs in loops of increasing length while sampling SM clock and power (NVML)."""
import ctypes as C
import os
import sys
import threading
import time

import pynvml
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import synth, _lib  # noqa: E402
from ngp_pl_b200.models.networks import NGP  # noqa: E402
from ngp_pl_b200.trainer import Trainer  # noqa: E402

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False


def sampler():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.002)


th = threading.Thread(target=sampler, daemon=True)
th.start()

scene = synth.lego_scene(0)
bank = synth.RayBank(scene, n_images=100, device="cuda")
model = NGP(0.5).cuda()
tr = Trainer(model, n_rays=8192)
tr.attach_bank(bank)
tr.capture(sample=True)
for _ in range(1000):
    tr.train_step()
torch.cuda.synchronize()
cur = tr._cur


def run(name, fn, n):
    torch.cuda.synchronize()
    time.sleep(0.3)  # let the chip idle back to boost
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    s = [x for x in samples if t0 <= x[0] <= t1]
    clk = sorted(x[1] for x in s) or [0]
    pw = sorted(x[2] for x in s) or [0]
    print("%-22s n=%5d  %.1f us/iter   sm MHz min/med/max %d/%d/%d   W med/max %.0f/%.0f  (%d samples)" % (
        name, n, a.elapsed_time(b) / n * 1e3, clk[0], clk[len(clk) // 2], clk[-1], pw[len(pw) // 2], pw[-1], len(s)))


def cu():
    tr.g_compute[cur][0].replay()
    tr.g_update[0].replay()


for n in (50, 300, 2000):
    run("C", tr.g_compute[cur][0].replay, n)
for n in (50, 300, 2000):
    run("U", tr.g_update[0].replay, n)
for n in (50, 300, 2000):
    run("C;U", cu, n)
for n in (50, 300, 2000):
    run("train_step", tr.train_step, n)
stop = True
