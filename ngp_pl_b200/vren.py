"""Drop-in for the reference's pybind11 module `vren` (reference models/csrc/binding.cpp:234-250).

Same twelve names, same positional arguments, same return lists, same in-place mutations. Tensors are
allocated with torch (plumbing); all arithmetic happens in libngp_b200.so through its C ABI. Like the
reference (models/csrc/include/utils.h:4-6) each tensor argument must be a contiguous CUDA tensor,
otherwise RuntimeError.

    import ngp_pl_b200.vren as vren        # instead of `import vren`
"""
import torch

from . import _lib


def _chk(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("ngp_pl_b200.vren: argument must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError("ngp_pl_b200.vren: argument must be contiguous")


def _st():
    return torch.cuda.current_stream().cuda_stream


class _Same:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_SAME = _Same()


def _on(dev):
    """device guard that costs nothing when `dev` already is the current device (the reference has none at all)"""
    if dev.index is None or dev.index == torch.cuda.current_device():
        return _SAME
    return torch.cuda.device(dev)


def _p(t):
    return t.data_ptr()


def _f32(t):
    if t.dtype != torch.float32:
        raise RuntimeError("ngp_pl_b200.vren: expected float32, got %s" % t.dtype)
    return t


def _intersect(fn_name, rays_o, rays_d, centers, extents, max_hits):
    _chk(rays_o, rays_d, centers, extents)
    n_rays, n_obj = rays_o.shape[0], centers.shape[0]
    dev = rays_o.device
    with _on(dev):
        hits_t = torch.empty(n_rays, max_hits, 2, device=dev, dtype=torch.float32)
        hits_idx = torch.empty(n_rays, max_hits, device=dev, dtype=torch.int64)
        hit_cnt = torch.empty(n_rays, device=dev, dtype=torch.int32)
        rc = getattr(_lib.lib(), fn_name)(_p(_f32(rays_o)), _p(_f32(rays_d)), _p(_f32(centers)), _p(_f32(extents)),
                                          n_rays, n_obj, int(max_hits), _p(hit_cnt), _p(hits_t), _p(hits_idx), _st())
        _lib.check(rc, fn_name)
        if max_hits > 1 or n_obj > 1:
            # near-to-far ordering (reference intersection.cu:95-97, :192-194); identity for the hot path
            order = torch.sort(hits_t[..., 0])[1]
            hits_idx = torch.gather(hits_idx, 1, order)
            hits_t = torch.gather(hits_t, 1, order.unsqueeze(-1).tile((1, 1, 2)))
    return [hit_cnt, hits_t, hits_idx]


def ray_aabb_intersect(rays_o, rays_d, centers, half_sizes, max_hits):
    """reference binding.cpp:4-16"""
    return _intersect("ngp_ray_aabb_intersect", rays_o, rays_d, centers, half_sizes, max_hits)


def ray_sphere_intersect(rays_o, rays_d, centers, radii, max_hits):
    """reference binding.cpp:19-31"""
    return _intersect("ngp_ray_sphere_intersect", rays_o, rays_d, centers, radii, max_hits)


_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.float64: 2}


def packbits(density_grid, density_threshold, density_bitfield):
    """reference binding.cpp:34-43 (in place on density_bitfield)"""
    _chk(density_grid, density_bitfield)
    if density_grid.dtype not in _DTYPE_CODE:
        raise RuntimeError("packbits: unsupported dtype %s" % density_grid.dtype)
    with _on(density_grid.device):
        rc = _lib.lib().ngp_packbits(_p(density_grid), _DTYPE_CODE[density_grid.dtype], density_bitfield.shape[0],
                                     float(density_threshold), None, _p(density_bitfield), _st())
    _lib.check(rc, "packbits")


def morton3D(coords):
    """reference binding.cpp:46-50"""
    _chk(coords)
    if coords.dtype != torch.int32:
        raise RuntimeError("morton3D: expected int32 coords")
    out = torch.empty(coords.shape[0], device=coords.device, dtype=torch.int32)
    with _on(coords.device):
        _lib.check(_lib.lib().ngp_morton3D(_p(coords), coords.shape[0], _p(out), _st()), "morton3D")
    return out


def morton3D_invert(indices):
    """reference binding.cpp:53-57"""
    _chk(indices)
    if indices.dtype != torch.int32:
        raise RuntimeError("morton3D_invert: expected int32 indices")
    out = torch.empty(indices.shape[0], 3, device=indices.device, dtype=torch.int32)
    with _on(indices.device):
        _lib.check(_lib.lib().ngp_morton3D_invert(_p(indices), indices.shape[0], _p(out), _st()), "morton3D_invert")
    return out


def raymarching_train(rays_o, rays_d, hits_t, density_bitfield, cascades, scale, exp_step_factor, noise,
                      grid_size, max_samples):
    """reference binding.cpp:60-81. Returns [rays_a, xyzs, dirs, deltas, ts, counter]; the per-sample
    tensors have N_rays*max_samples rows of which only the first counter[0] are defined."""
    _chk(rays_o, rays_d, hits_t, density_bitfield, noise)
    n_rays = rays_o.shape[0]
    dev = rays_o.device
    cap = n_rays * int(max_samples)
    with _on(dev):
        rays_a = torch.empty(n_rays, 3, device=dev, dtype=torch.int64)
        xyzs = torch.empty(cap, 3, device=dev, dtype=torch.float32)
        dirs = torch.empty(cap, 3, device=dev, dtype=torch.float32)
        deltas = torch.empty(cap, device=dev, dtype=torch.float32)
        ts = torch.empty(cap, device=dev, dtype=torch.float32)
        counter = torch.empty(2, device=dev, dtype=torch.int32)
        L = _lib.lib()
        ws_bytes = L.ngp_raymarching_train_workspace2(n_rays, int(max_samples))
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        rc = L.ngp_raymarching_train(_p(_f32(rays_o)), _p(_f32(rays_d)), _p(_f32(hits_t)), _p(density_bitfield),
                                     int(cascades), float(scale), float(exp_step_factor), _p(_f32(noise)),
                                     int(grid_size), int(max_samples), n_rays,
                                     _p(rays_a), _p(xyzs), _p(dirs), _p(deltas), _p(ts), _p(counter),
                                     _p(ws), ws_bytes, _st())
        _lib.check(rc, "raymarching_train")
    return [rays_a, xyzs, dirs, deltas, ts, counter]


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, density_bitfield, cascades, scale, exp_step_factor,
                     grid_size, max_samples, N_samples):
    """reference binding.cpp:84-106 (hits_t[:,0] is advanced in place)"""
    _chk(rays_o, rays_d, hits_t, alive_indices, density_bitfield)
    if alive_indices.dtype != torch.int64:
        raise RuntimeError("raymarching_test: alive_indices must be int64")
    n_alive = alive_indices.shape[0]
    dev = rays_o.device
    with _on(dev):
        xyzs = torch.empty(n_alive, N_samples, 3, device=dev, dtype=torch.float32)
        dirs = torch.empty(n_alive, N_samples, 3, device=dev, dtype=torch.float32)
        deltas = torch.empty(n_alive, N_samples, device=dev, dtype=torch.float32)
        ts = torch.empty(n_alive, N_samples, device=dev, dtype=torch.float32)
        n_eff = torch.empty(n_alive, device=dev, dtype=torch.int32)
        rc = _lib.lib().ngp_raymarching_test(_p(_f32(rays_o)), _p(_f32(rays_d)), _p(_f32(hits_t)), _p(alive_indices),
                                             _p(density_bitfield), int(cascades), float(scale), float(exp_step_factor),
                                             int(grid_size), int(max_samples), int(N_samples), n_alive,
                                             _p(xyzs), _p(dirs), _p(deltas), _p(ts), _p(n_eff), _st())
        _lib.check(rc, "raymarching_test")
    return [xyzs, dirs, deltas, ts, n_eff]


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    """reference binding.cpp:109-126. Returns [total_samples, opacity, depth, rgb, ws]."""
    _chk(sigmas, rgbs, deltas, ts, rays_a)
    n_rays, n = rays_a.shape[0], sigmas.shape[0]
    dev = sigmas.device
    with _on(dev):
        total = torch.empty(n_rays, device=dev, dtype=torch.int64)
        opacity = torch.empty(n_rays, device=dev, dtype=torch.float32)
        depth = torch.empty(n_rays, device=dev, dtype=torch.float32)
        rgb = torch.empty(n_rays, 3, device=dev, dtype=torch.float32)
        ws = torch.empty(n, device=dev, dtype=torch.float32)
        rc = _lib.lib().ngp_composite_train_fw(_p(_f32(sigmas)), _p(_f32(rgbs)), _p(_f32(deltas)), _p(_f32(ts)),
                                               _p(rays_a), float(T_threshold), n_rays, n,
                                               _p(total), _p(opacity), _p(depth), _p(rgb), _p(ws), _st())
        _lib.check(rc, "composite_train_fw")
    return [total, opacity, depth, rgb, ws]


def composite_train_bw(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a,
                       opacity, depth, rgb, T_threshold):
    """reference binding.cpp:129-163. Returns [dL_dsigmas, dL_drgbs]."""
    _chk(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb)
    n_rays, n = rays_a.shape[0], sigmas.shape[0]
    dev = sigmas.device
    with _on(dev):
        dsig = torch.empty(n, device=dev, dtype=torch.float32)
        drgbs = torch.empty(n, 3, device=dev, dtype=torch.float32)
        rc = _lib.lib().ngp_composite_train_bw(
            _p(_f32(dL_dopacity)), _p(_f32(dL_ddepth)), _p(_f32(dL_drgb)), _p(_f32(dL_dws)), _p(_f32(sigmas)),
            _p(_f32(rgbs)), _p(_f32(ws)), _p(_f32(deltas)), _p(_f32(ts)), _p(rays_a), _p(_f32(opacity)),
            _p(_f32(depth)), _p(_f32(rgb)), float(T_threshold), n_rays, n, _p(dsig), _p(drgbs), _st())
        _lib.check(rc, "composite_train_bw")
    return [dsig, drgbs]


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive_indices, T_threshold, N_eff_samples,
                      opacity, depth, rgb):
    """reference binding.cpp:166-194 (alive_indices, opacity, depth, rgb updated in place)"""
    _chk(sigmas, rgbs, deltas, ts, hits_t, alive_indices, N_eff_samples, opacity, depth, rgb)
    n_alive = alive_indices.shape[0]
    n_samples = sigmas.shape[1] if sigmas.dim() == 2 else 1
    with _on(sigmas.device):
        rc = _lib.lib().ngp_composite_test_fw(_p(_f32(sigmas)), _p(_f32(rgbs)), _p(_f32(deltas)), _p(_f32(ts)),
                                              _p(hits_t), _p(alive_indices), float(T_threshold), _p(N_eff_samples),
                                              n_alive, n_samples, _p(_f32(opacity)), _p(_f32(depth)), _p(_f32(rgb)),
                                              _st())
        _lib.check(rc, "composite_test_fw")


def distortion_loss_fw(ws, deltas, ts, rays_a):
    """reference binding.cpp:197-209. Returns [loss, ws_inclusive_scan, wts_inclusive_scan]."""
    _chk(ws, deltas, ts, rays_a)
    n_rays, n = rays_a.shape[0], ws.shape[0]
    dev = ws.device
    with _on(dev):
        loss = torch.zeros(n_rays, device=dev, dtype=torch.float32)
        ws_inc = torch.empty(n, device=dev, dtype=torch.float32)
        wts_inc = torch.empty(n, device=dev, dtype=torch.float32)
        rc = _lib.lib().ngp_distortion_loss_fw(_p(_f32(ws)), _p(_f32(deltas)), _p(_f32(ts)), _p(rays_a), n_rays, n,
                                               _p(loss), _p(ws_inc), _p(wts_inc), _st())
        _lib.check(rc, "distortion_loss_fw")
    return [loss, ws_inc, wts_inc]


def distortion_loss_bw(dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a):
    """reference binding.cpp:212-231. Returns dL_dws."""
    _chk(dL_dloss, ws_inclusive_scan, wts_inclusive_scan, ws, deltas, ts, rays_a)
    n_rays, n = rays_a.shape[0], ws.shape[0]
    dev = ws.device
    with _on(dev):
        dws = torch.zeros(n, device=dev, dtype=torch.float32)
        rc = _lib.lib().ngp_distortion_loss_bw(_p(_f32(dL_dloss)), _p(_f32(ws_inclusive_scan)),
                                               _p(_f32(wts_inclusive_scan)), _p(_f32(ws)), _p(_f32(deltas)),
                                               _p(_f32(ts)), _p(rays_a), n_rays, n, _p(dws), _st())
        _lib.check(rc, "distortion_loss_bw")
    return dws
