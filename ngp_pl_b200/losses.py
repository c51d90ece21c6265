"""Host mirror of the reference's loss module, running on this repo's operators.

What a reference user imports stays importable under the same names:

* ``DistortionLoss.apply(ws, deltas, ts, rays_a) -> per-ray loss``  -- reference losses.py:6-37
* ``NeRFLoss(lambda_opacity, lambda_distortion)(results, target) -> dict of un-reduced terms`` -- reference losses.py:40-60

The arithmetic lives in libngp_b200.so (``ngp_distortion_loss_fw/_bw`` through ``ngp_pl_b200.vren``); nothing here
computes on the CPU.
"""
import torch

from . import vren

_EPS_OPACITY = 1e-10  # keeps log() finite for fully transparent rays


def _contig(*tensors):
    return tuple(t.contiguous() for t in tensors)


class DistortionLoss(torch.autograd.Function):
    """Distortion regulariser of Mip-NeRF 360 evaluated per ray with inclusive prefix sums (the DVGO-v2 formulation).

    ws, deltas, ts: one value per sample, samples of a ray contiguous; rays_a: (n_rays, 3) rows of
    [ray index, first sample, sample count]. Returns one loss value per ray; only ``ws`` receives a gradient.
    """

    @staticmethod
    def forward(ctx, ws, deltas, ts, rays_a):
        ws_c, deltas_c, ts_c = _contig(ws, deltas, ts)
        per_ray, scan_w, scan_wt = vren.distortion_loss_fw(ws_c, deltas_c, ts_c, rays_a)
        ctx.rays_a = rays_a
        ctx.save_for_backward(scan_w, scan_wt, ws_c, deltas_c, ts_c)
        return per_ray

    @staticmethod
    def backward(ctx, grad_per_ray):
        scan_w, scan_wt, ws_c, deltas_c, ts_c = ctx.saved_tensors
        grad_ws = vren.distortion_loss_bw(grad_per_ray.contiguous(), scan_w, scan_wt, ws_c, deltas_c, ts_c, ctx.rays_a)
        return grad_ws, None, None, None


class NeRFLoss(torch.nn.Module):
    """Photometric term + opacity entropy (+ distortion when its weight is positive). Returns the per-element terms in a
    dict -- the caller reduces them (the reference's training loop sums their means)."""

    def __init__(self, lambda_opacity=1e-3, lambda_distortion=1e-3):
        super().__init__()
        self.lambda_opacity, self.lambda_distortion = lambda_opacity, lambda_distortion

    @staticmethod
    def _photometric(pred, gt):
        diff = pred - gt
        return diff * diff

    def _opacity_entropy(self, opacity):
        o = opacity + _EPS_OPACITY
        return self.lambda_opacity * (o * torch.log(o)).neg()  # minimal at o -> 0 and o -> 1

    def forward(self, results, target, **kwargs):
        terms = {"rgb": self._photometric(results["rgb"], target["rgb"]),
                 "opacity": self._opacity_entropy(results["opacity"])}
        if self.lambda_distortion > 0:
            per_ray = DistortionLoss.apply(results["ws"], results["deltas"], results["ts"], results["rays_a"])
            terms["distortion"] = self.lambda_distortion * per_ray
        return terms
