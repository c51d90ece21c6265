"""Times the three step graphs of the Trainer alone and overlapped (CUDA events)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ngp_pl_b200 import synth
from ngp_pl_b200.models.networks import NGP
from ngp_pl_b200.trainer import Trainer

scene = synth.lego_scene(0)
bank = synth.RayBank(scene, n_images=100, device="cuda")
model = NGP(0.5).cuda()
tr = Trainer(model, n_rays=8192)
tr.attach_bank(bank)
tr.capture(sample=True)
for _ in range(600):
    tr.train_step()
torch.cuda.synchronize()
main = torch.cuda.current_stream()
side = torch.cuda.Stream()

def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

def both():
    side.wait_stream(main)
    with torch.cuda.stream(side):
        tr.g_prepare.replay()
    tr.g_update.replay()
    main.wait_stream(side)

print("prepare us", t(tr.g_prepare.replay))
print("compute us", t(tr.g_compute.replay))
print("update  us", t(tr.g_update.replay))
print("prepare||update us", t(both))
def seq():
    tr.g_prepare.replay(); tr.g_compute.replay(); tr.g_update.replay()
print("sequential all three us", t(seq))
def full():
    tr.train_step()
print("train_step us", t(full, 200))
