"""CosineAnnealingLR of ngp_pl_b200.trainer against torch's scheduler configured as reference train.py:135-137
(T_max = num_epochs, eta_min = lr/30) and stepped once per epoch as pytorch-lightning does."""
import torch

from ngp_pl_b200.trainer import CosineAnnealingLR


def test_cosine_schedule_matches_torch_per_epoch():
    lr0, epochs = 1e-2, 30
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=lr0, eps=1e-15)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, epochs, lr0 / 30)
    ours = CosineAnnealingLR(lr0, T_max=epochs, steps_per_epoch=1000)
    for e in range(epochs):
        want = opt.param_groups[0]["lr"]
        for s in (0, 1, 999):
            assert abs(ours.lr_at_step(1000 * e + s) - want) < 1e-12 * max(1.0, want), (e, s)
        opt.step()
        sch.step()
    assert abs(ours.lr_at_step(0) - lr0) < 1e-15
    assert abs(ours.lr_at_epoch(epochs) - lr0 / 30) < 1e-15
