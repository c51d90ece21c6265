"""The five torch.autograd.Function wrappers of reference models/custom_functions.py, same class names
and `.apply` argument order, running on libngp_b200.so.

    RayAABBIntersector   reference custom_functions.py:8-29
    RaySphereIntersector reference custom_functions.py:32-52
    RayMarcher           reference custom_functions.py:55-112
    VolumeRenderer       reference custom_functions.py:115-159
    TruncExp             reference custom_functions.py:162-173
"""
import torch
from torch.amp import custom_bwd, custom_fwd

from .. import vren


class RayAABBIntersector(torch.autograd.Function):
    """rays (N,3) x boxes (V,3) -> hits_cnt (N), hits_t (N,max_hits,2) near-to-far (-1: no hit),
    hits_voxel_idx (N,max_hits)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, center, half_size, max_hits):
        return tuple(vren.ray_aabb_intersect(rays_o, rays_d, center, half_size, max_hits))


class RaySphereIntersector(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, center, radii, max_hits):
        return tuple(vren.ray_sphere_intersect(rays_o, rays_d, center, radii, max_hits))


class RayMarcher(torch.autograd.Function):
    """March rays through the occupancy bitfield.

    Inputs : rays_o, rays_d (N,3); hits_t (N,2); density_bitfield (C*G^3/8) uint8; cascades; scale;
             exp_step_factor; grid_size; max_samples
    Outputs: rays_a (N,3) [ray_idx,start_idx,N_samples]; xyzs, dirs (S,3); deltas, ts (S); total_samples
    An optional `noise` attribute (class-level, (N,) tensor) replaces the internally drawn start
    jitter -- used by parity tests to march identical rays with the reference kernels.
    """
    noise_override = None

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, grid_size, max_samples):
        if RayMarcher.noise_override is not None:
            noise = RayMarcher.noise_override
        else:
            noise = torch.rand_like(rays_o[:, 0])
        rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(
            rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise, grid_size, max_samples)
        total_samples = counter[0]
        n = int(total_samples)  # the one host sync of this (unfused) API, as in the reference (:91-96)
        xyzs, dirs, deltas, ts = xyzs[:n], dirs[:n], deltas[:n], ts[:n]
        ctx.save_for_backward(rays_a, ts)
        return rays_a, xyzs, dirs, deltas, ts, total_samples

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dL_drays_a, dL_dxyzs, dL_ddirs, dL_ddeltas, dL_dts, dL_dtotal_samples):
        rays_a, ts = ctx.saved_tensors
        # per-ray sums of the per-sample gradients (the reference uses torch_scatter.segment_csr)
        n_rays = rays_a.shape[0]
        seg = torch.repeat_interleave(torch.arange(n_rays, device=ts.device), rays_a[:, 2])
        dL_drays_o = torch.zeros(n_rays, 3, device=ts.device, dtype=dL_dxyzs.dtype).index_add_(0, seg, dL_dxyzs)
        dL_drays_d = torch.zeros(n_rays, 3, device=ts.device, dtype=dL_dxyzs.dtype).index_add_(
            0, seg, dL_dxyzs * ts[:, None] + dL_ddirs)
        # rows of rays_a are ordered by ray index here, so row order == ray order
        return dL_drays_o, dL_drays_d, None, None, None, None, None, None, None


class VolumeRenderer(torch.autograd.Function):
    """Ragged front-to-back compositing (training).

    Inputs : sigmas (S); rgbs (S,3); deltas (S); ts (S); rays_a (N,3); T_threshold
    Outputs: total_samples (scalar); opacity (N); depth (N); rgb (N,3); ws (S)
    """

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, ts, rays_a, T_threshold):
        total_samples, opacity, depth, rgb, ws = vren.composite_train_fw(
            sigmas.contiguous(), rgbs.contiguous(), deltas.contiguous(), ts.contiguous(), rays_a, T_threshold)
        ctx.save_for_backward(sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws)
        ctx.T_threshold = T_threshold
        return total_samples.sum(), opacity, depth, rgb, ws

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dL_dtotal_samples, dL_dopacity, dL_ddepth, dL_drgb, dL_dws):
        sigmas, rgbs, deltas, ts, rays_a, opacity, depth, rgb, ws = ctx.saved_tensors
        dL_dsigmas, dL_drgbs = vren.composite_train_bw(
            dL_dopacity.contiguous(), dL_ddepth.contiguous(), dL_drgb.contiguous(), dL_dws.contiguous(),
            sigmas.contiguous(), rgbs.contiguous(), ws, deltas.contiguous(), ts.contiguous(), rays_a,
            opacity, depth, rgb, ctx.T_threshold)
        return dL_dsigmas, dL_drgbs, None, None, None, None


class TruncExp(torch.autograd.Function):
    """exp with a clamped backward (reference custom_functions.py:162-173)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, dL_dout):
        x = ctx.saved_tensors[0]
        return dL_dout * torch.exp(x.clamp(-15, 15))
