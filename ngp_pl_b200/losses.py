"""Losses of the reference (reference losses.py:6-60) on the B200 operators.

    DistortionLoss  torch.autograd.Function over vren.distortion_loss_fw / _bw   (losses.py:6-37)
    NeRFLoss        rgb MSE + opacity entropy + optional distortion               (losses.py:40-60)
"""
import torch
from torch import nn

from . import vren


class DistortionLoss(torch.autograd.Function):
    """Mip-NeRF 360 distortion loss, DVGO-v2 prefix-sum form.
    Inputs: ws (S), deltas (S), ts (S), rays_a (N,3) [ray_idx, start_idx, N_samples] -> loss (N)"""

    @staticmethod
    def forward(ctx, ws, deltas, ts, rays_a):
        loss, ws_inc, wts_inc = vren.distortion_loss_fw(ws.contiguous(), deltas.contiguous(), ts.contiguous(), rays_a)
        ctx.save_for_backward(ws_inc, wts_inc, ws, deltas, ts, rays_a)
        return loss

    @staticmethod
    def backward(ctx, dL_dloss):
        ws_inc, wts_inc, ws, deltas, ts, rays_a = ctx.saved_tensors
        dL_dws = vren.distortion_loss_bw(dL_dloss.contiguous(), ws_inc, wts_inc, ws.contiguous(), deltas.contiguous(),
                                         ts.contiguous(), rays_a)
        return dL_dws, None, None, None


class NeRFLoss(nn.Module):
    def __init__(self, lambda_opacity=1e-3, lambda_distortion=1e-3):
        super().__init__()
        self.lambda_opacity = lambda_opacity
        self.lambda_distortion = lambda_distortion

    def forward(self, results, target, **kwargs):
        d = {'rgb': (results['rgb'] - target['rgb']) ** 2}
        o = results['opacity'] + 1e-10
        d['opacity'] = self.lambda_opacity * (-o * torch.log(o))  # pushes opacity towards 0 or 1
        if self.lambda_distortion > 0:
            d['distortion'] = self.lambda_distortion * DistortionLoss.apply(results['ws'], results['deltas'], results['ts'],
                                                                            results['rays_a'])
        return d
