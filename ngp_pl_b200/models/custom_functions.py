"""Autograd entry points of the operator-level API.

A reference user calls five ``torch.autograd.Function``s from ``models/custom_functions.py``; the same five names with the
same ``.apply`` argument order are defined here on top of ``ngp_pl_b200.vren`` (libngp_b200.so):

    ================================  =========================================
    RayAABBIntersector.apply(o, d, center, half_size, max_hits)   reference custom_functions.py:8-29
    RaySphereIntersector.apply(o, d, center, radii, max_hits)     reference custom_functions.py:32-52
    RayMarcher.apply(o, d, hits_t, bitfield, cascades, scale, exp_step_factor, grid_size, max_samples)   :55-112
    VolumeRenderer.apply(sigmas, rgbs, deltas, ts, rays_a, T_threshold)                                  :115-159
    TruncExp.apply(x)                                                                                    :162-173
    ================================  =========================================

Like the reference's wrappers they run the native operators in fp32 whatever the autocast state is.
"""
import torch
from torch.amp import custom_bwd, custom_fwd

from .. import vren

_fp32_forward = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_backward = custom_bwd(device_type="cuda")


def _intersector(operator_name, summary):
    """Function class for a ray / primitive intersection operator: no gradient, three outputs
    (hit count (N), hit intervals (N, max_hits, 2) sorted near to far with -1 = no hit, primitive index (N, max_hits))."""

    class _Intersect(torch.autograd.Function):
        __doc__ = summary

        @staticmethod
        @_fp32_forward
        def forward(ctx, rays_o, rays_d, center, extent, max_hits):
            hit_cnt, hits_t, hits_idx = getattr(vren, operator_name)(rays_o, rays_d, center, extent, max_hits)
            return hit_cnt, hits_t, hits_idx

    return _Intersect


RayAABBIntersector = _intersector("ray_aabb_intersect", "Rays (N,3) against axis-aligned boxes given by centre (V,3) and half size (V,3).")
RayAABBIntersector.__name__ = RayAABBIntersector.__qualname__ = "RayAABBIntersector"
RaySphereIntersector = _intersector("ray_sphere_intersect", "Rays (N,3) against spheres given by centre (V,3) and radius (V).")
RaySphereIntersector.__name__ = RaySphereIntersector.__qualname__ = "RaySphereIntersector"


def _sum_per_ray(per_sample, rays_a):
    """(S, C) per-sample values -> (N, C) per-ray sums, row r = the ray rays_a[r, 0]. rays_a rows are [ray, first sample, count]
    in ANY row order (the reference's marcher emits them in atomic order); a ray's samples are contiguous from its first
    sample. The reference uses torch_scatter.segment_csr on the same offsets (custom_functions.py:107-110), which
    additionally assumes rows sorted by first sample; this version does not."""
    n_rays, S = rays_a.shape[0], per_sample.shape[0]
    # owner row of every sample: +1 at each non-empty row's first sample (in sample order), running count - 1
    order = torch.argsort(rays_a[:, 1], stable=True)
    counts = rays_a[order, 2]
    owner_sorted = torch.repeat_interleave(torch.arange(n_rays, device=per_sample.device), counts, output_size=S)
    owner = rays_a[order, 0][owner_sorted]
    return torch.zeros(n_rays, per_sample.shape[1], device=per_sample.device, dtype=per_sample.dtype).index_add_(0, owner, per_sample)


class RayMarcher(torch.autograd.Function):
    """Occupancy-grid march of a batch of rays (training): every ray contributes the sample positions that fall into
    occupied cells between its entry and exit of the scene box, with a random sub-step start offset.

    Returns ``rays_a`` (N,3) int64 rows [ray index, first sample, sample count], ``xyzs``/``dirs`` (S,3), ``deltas``/``ts`` (S)
    and the total S (a 0-d tensor). Gradients flow to ``rays_o`` / ``rays_d`` (x = o + t d), as in the reference.
    ``RayMarcher.noise_override`` (an (N,) tensor, class attribute) replaces the internally drawn start offsets -- the parity
    tests use it to march identical rays through the reference's kernels.
    """
    noise_override = None

    @staticmethod
    @_fp32_forward
    def forward(ctx, rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples):
        jitter = RayMarcher.noise_override
        if jitter is None:
            jitter = torch.rand(rays_o.shape[0], device=rays_o.device, dtype=rays_o.dtype)
        rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(
            rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, jitter, grid_size, max_samples)
        total = counter[0]
        used = int(total)  # the operator-level API sizes its outputs on the host, like the reference (:91-96): one sync
        ts_used = ts[:used]
        ctx.save_for_backward(rays_a, ts_used)
        return rays_a, xyzs[:used], dirs[:used], deltas[:used], ts_used, total

    @staticmethod
    @_amp_backward
    def backward(ctx, g_rays_a, g_xyzs, g_dirs, g_deltas, g_ts, g_total):
        rays_a, ts = ctx.saved_tensors
        # x = o + t d and dirs = d per sample  =>  dL/do = sum g_x ,  dL/dd = sum (t g_x + g_dirs)   over the ray's samples
        grad_o = _sum_per_ray(g_xyzs, rays_a)
        grad_d = _sum_per_ray(g_xyzs * ts.unsqueeze(1) + g_dirs, rays_a)
        return (grad_o, grad_d) + (None,) * 7


class VolumeRenderer(torch.autograd.Function):
    """Front-to-back alpha compositing of ragged per-ray sample lists (training).

    sigmas (S), rgbs (S,3), deltas (S), ts (S), rays_a (N,3), T_threshold -> total composited samples (0-d), opacity (N),
    depth (N), rgb (N,3), per-sample weights ws (S). Gradients flow to sigmas and rgbs.
    """

    @staticmethod
    @_fp32_forward
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        args = [t.contiguous() for t in (sigmas, rgbs, deltas, ts)]
        per_ray_count, opacity, depth, rgb, ws = vren.composite_train_fw(*args, rays_a, T_threshold)
        ctx.T_threshold = T_threshold
        ctx.save_for_backward(*args, rays_a, opacity, depth, rgb, ws)
        return per_ray_count.sum(), opacity, depth, rgb, ws

    @staticmethod
    @_amp_backward
    def backward(ctx, g_count, g_opacity, g_depth, g_rgb, g_ws):
        sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws = ctx.saved_tensors
        g_sigmas, g_rgbs = vren.composite_train_bw(
            g_opacity.contiguous(), g_depth.contiguous(), g_rgb.contiguous(), g_ws.contiguous(),
            sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb, ctx.T_threshold)
        return g_sigmas, g_rgbs, None, None, None, None


class TruncExp(torch.autograd.Function):
    """exp(x) whose derivative is evaluated at clamp(x, -15, 15): the density activation (a runaway pre-activation cannot
    blow the gradient up)."""

    @staticmethod
    @_fp32_forward
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return x.exp()

    @staticmethod
    @_amp_backward
    def backward(ctx, g_out):
        (x,) = ctx.saved_tensors
        return g_out * torch.clamp(x, min=-15.0, max=15.0).exp()
