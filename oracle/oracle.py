"""CPU ORACLE front-end -- TEST INFRASTRUCTURE ONLY (see the header of ngp_oracle.c).

Nothing under ngp_pl_b200/ imports this. Allowed callers: tests/, __graft_entry__.smoke(),
bench.py (cpu_baseline leg / --impl reference).

  * C restatement (ngp_oracle.c, gcc -O2 -ffp-contract=off) of the reference kernels
    models/csrc/{intersection,raymarching,volumerendering,losses}.cu  -> numpy in / numpy out
  * torch-CPU fp32 restatement of the tinycudann modules (hash grid, SH-4, the two MLPs) with the
    same fp16 rounding points as the CUDA kernels; its autograd is the gradient oracle.
    PARITY UNPINNED for this part: tinycudann is absent from /root/reference and not installable.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ngp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libngp_oracle.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout.decode())
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_morton3D.restype = C.c_uint32
        _lib.oracle_morton3D.argtypes = [C.c_uint32] * 3
        _lib.oracle_grid_meta_make.restype = C.c_uint32
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class GridMeta(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("hashed_mask", C.c_uint32), ("offset", C.c_uint32 * 17),
                ("res", C.c_uint32 * 16), ("scale", C.c_float * 16)]


def grid_meta(n_levels, log2_T, base_res, per_level_scale):
    m = GridMeta()
    total = lib().oracle_grid_meta_make(C.c_int(n_levels), C.c_int(log2_T), C.c_int(base_res),
                                        C.c_float(per_level_scale), C.byref(m))
    return m, int(total)


def morton3D(coords):
    coords = np.asarray(coords, dtype=np.uint32)
    return np.array([lib().oracle_morton3D(int(x), int(y), int(z)) for x, y, z in coords], dtype=np.int32)


def morton3D_invert(indices):
    out = np.zeros((len(indices), 3), dtype=np.int32)
    x, y, z = C.c_uint32(), C.c_uint32(), C.c_uint32()
    for i, m in enumerate(np.asarray(indices, dtype=np.uint32)):
        lib().oracle_morton3D_invert(C.c_uint32(int(m)), C.byref(x), C.byref(y), C.byref(z))
        out[i] = (x.value, y.value, z.value)
    return out


def packbits(grid, thr):
    grid = _f(grid).reshape(-1)
    bits = np.zeros(grid.size // 8, dtype=np.uint8)
    lib().oracle_packbits(_ptr(grid), C.c_int64(bits.size), C.c_float(thr), _ptr(bits))
    return bits


def ray_aabb(rays_o, rays_d, center, half_size, near_distance=0.0):
    o, d = _f(rays_o), _f(rays_d)
    hits = np.zeros((o.shape[0], 2), dtype=np.float32)
    lib().oracle_ray_aabb(_ptr(o), _ptr(d), _ptr(_f(center).reshape(-1)), _ptr(_f(half_size).reshape(-1)),
                          C.c_int(o.shape[0]), C.c_float(near_distance), _ptr(hits))
    return hits


def march_train(rays_o, rays_d, hits_t, bits, cascades, scale, esf, noise, grid_size, max_samples):
    """-> rays_a (n,3) int64 ordered by ray, xyzs, dirs, deltas, ts"""
    o, d, h, nz = _f(rays_o), _f(rays_d), _f(hits_t), _f(noise)
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    n = o.shape[0]
    cnt = np.zeros(n, dtype=np.int32)
    args = [_ptr(o), _ptr(d), _ptr(h), _ptr(bits), C.c_int(cascades), C.c_float(scale), C.c_float(esf), _ptr(nz),
            C.c_int(grid_size), C.c_int(max_samples), C.c_int(n)]
    lib().oracle_march_train(*args, _ptr(cnt), None, None, None, None, None)
    start = np.zeros(n, dtype=np.int32)
    start[1:] = np.cumsum(cnt)[:-1]
    total = int(cnt.sum())
    xyzs = np.zeros((total, 3), np.float32)
    dirs = np.zeros((total, 3), np.float32)
    deltas = np.zeros(total, np.float32)
    ts = np.zeros(total, np.float32)
    lib().oracle_march_train(*args, _ptr(cnt), _ptr(start), _ptr(xyzs), _ptr(dirs), _ptr(deltas), _ptr(ts))
    rays_a = np.stack([np.arange(n), start, cnt], 1).astype(np.int64)
    return rays_a, xyzs, dirs, deltas, ts


def march_probe(ray_o, ray_d, bits, cascades, scale, esf, dt_scale, grid_size, max_samples, t):
    """per-point part of one marcher visit for parameters t (n,) of ONE ray -> occ (n) int32, dt (n), t_target (n)"""
    o, d, tt = _f(ray_o).reshape(3), _f(ray_d).reshape(3), _f(t).reshape(-1)
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    n = tt.shape[0]
    occ = np.zeros(n, np.int32)
    dt = np.zeros(n, np.float32)
    tgt = np.zeros(n, np.float32)
    lib().oracle_march_probe(_ptr(o), _ptr(d), _ptr(bits), C.c_int(cascades), C.c_float(scale), C.c_float(esf), C.c_float(dt_scale),
                             C.c_int(grid_size), C.c_int(max_samples), _ptr(tt), C.c_int(n), _ptr(occ), _ptr(dt), _ptr(tgt))
    return occ, dt, tgt


def march_test(rays_o, rays_d, hits_t, alive, bits, cascades, scale, esf, grid_size, max_samples, N_samples):
    """hits_t (n_rays,2) float32 array is modified in place. -> xyzs, dirs, deltas, ts, n_eff"""
    o, d = _f(rays_o), _f(rays_d)
    assert hits_t.dtype == np.float32 and hits_t.flags["C_CONTIGUOUS"]
    alive = np.ascontiguousarray(alive, dtype=np.int64)
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    na = alive.shape[0]
    xyzs = np.zeros((na, N_samples, 3), np.float32)
    dirs = np.zeros((na, N_samples, 3), np.float32)
    deltas = np.zeros((na, N_samples), np.float32)
    ts = np.zeros((na, N_samples), np.float32)
    n_eff = np.zeros(na, np.int32)
    lib().oracle_march_test(_ptr(o), _ptr(d), _ptr(hits_t), _ptr(alive), _ptr(bits), C.c_int(cascades), C.c_float(scale),
                            C.c_float(esf), C.c_int(grid_size), C.c_int(max_samples), C.c_int(N_samples), C.c_int(na),
                            _ptr(xyzs), _ptr(dirs), _ptr(deltas), _ptr(ts), _ptr(n_eff))
    return xyzs, dirs, deltas, ts, n_eff


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_thr):
    s, c, dl, t = _f(sigmas), _f(rgbs), _f(deltas), _f(ts)
    ra = np.ascontiguousarray(rays_a, dtype=np.int64)
    nr, n = ra.shape[0], s.shape[0]
    total = np.zeros(nr, np.int64)
    opacity = np.zeros(nr, np.float32)
    depth = np.zeros(nr, np.float32)
    rgb = np.zeros((nr, 3), np.float32)
    ws = np.zeros(n, np.float32)
    lib().oracle_composite_train_fw(_ptr(s), _ptr(c), _ptr(dl), _ptr(t), _ptr(ra), C.c_float(T_thr), C.c_int(nr),
                                    C.c_int64(n), _ptr(total), _ptr(opacity), _ptr(depth), _ptr(rgb), _ptr(ws))
    return total, opacity, depth, rgb, ws


def composite_train_bw(dO, dD, dC, dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb, T_thr):
    ra = np.ascontiguousarray(rays_a, dtype=np.int64)
    nr, n = ra.shape[0], np.asarray(sigmas).shape[0]
    dsig = np.zeros(n, np.float32)
    drgbs = np.zeros((n, 3), np.float32)
    a = [_f(x) for x in (dO, dD, dC, dws, sigmas, rgbs, ws, deltas, ts)]
    b = [_f(x) for x in (opacity, depth, rgb)]
    lib().oracle_composite_train_bw(*[_ptr(x) for x in a], _ptr(ra), *[_ptr(x) for x in b], C.c_float(T_thr),
                                    C.c_int(nr), C.c_int64(n), _ptr(dsig), _ptr(drgbs))
    return dsig, drgbs


def composite_test_fw(sigmas, rgbs, deltas, ts, alive, T_thr, n_eff, opacity, depth, rgb):
    """alive (int64), opacity, depth, rgb (float32, C-contiguous) are modified in place."""
    s, c, dl, t = _f(sigmas), _f(rgbs), _f(deltas), _f(ts)
    ne = np.ascontiguousarray(n_eff, dtype=np.int32)
    for a in (opacity, depth, rgb):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    assert alive.dtype == np.int64
    lib().oracle_composite_test_fw(_ptr(s), _ptr(c), _ptr(dl), _ptr(t), _ptr(alive), C.c_float(T_thr), _ptr(ne),
                                   C.c_int(alive.shape[0]), C.c_int(s.shape[1]), _ptr(opacity), _ptr(depth), _ptr(rgb))


def distortion_fw(ws, deltas, ts, rays_a):
    w, dl, t = _f(ws), _f(deltas), _f(ts)
    ra = np.ascontiguousarray(rays_a, dtype=np.int64)
    loss = np.zeros(ra.shape[0], np.float32)
    wi = np.zeros_like(w)
    wti = np.zeros_like(w)
    lib().oracle_distortion_fw(_ptr(w), _ptr(dl), _ptr(t), _ptr(ra), C.c_int(ra.shape[0]), _ptr(loss), _ptr(wi), _ptr(wti))
    return loss, wi, wti


def distortion_bw(dL, ws_inc, wts_inc, ws, deltas, ts, rays_a):
    ra = np.ascontiguousarray(rays_a, dtype=np.int64)
    dws = np.zeros_like(_f(ws))
    a = [_f(x) for x in (dL, ws_inc, wts_inc, ws, deltas, ts)]
    lib().oracle_distortion_bw(*[_ptr(x) for x in a], _ptr(ra), C.c_int(ra.shape[0]), _ptr(dws))
    return dws


def ngp_forward_c(meta, enc_params, rgb_params, xyz_min, xyz_max, xyzs, dirs, want_rgb=True, rgb_act=1):
    """C restatement of NGP.forward. enc_params / rgb_params: fp32 arrays; rounded to fp16 here, as
    tinycudann casts its fp32 master parameters every forward."""
    enc = _f(np.asarray(enc_params, np.float32).astype(np.float16))
    rgbp = _f(np.asarray(rgb_params, np.float32).astype(np.float16))
    x, d = _f(xyzs), _f(dirs)
    n = x.shape[0]
    sig = np.zeros(n, np.float32)
    rgbs = np.zeros((n, 3), np.float32)
    h = np.zeros((n, 16), np.float32)
    lib().oracle_ngp_forward(C.byref(meta), _ptr(enc), _ptr(rgbp), _ptr(_f(xyz_min).reshape(-1)),
                             _ptr(_f(xyz_max).reshape(-1)), _ptr(x), _ptr(d), C.c_int64(n), C.c_int(int(want_rgb)),
                             C.c_int(rgb_act), _ptr(sig), _ptr(rgbs), _ptr(h))
    return sig, rgbs, h


# ---------------------------------------------------------------------------------------------------
# torch-CPU restatement of the tinycudann modules (gradient oracle)
# ---------------------------------------------------------------------------------------------------
def _rt(x):
    """fp16 rounding point with a straight-through fp32 gradient (x.half().float() alone would also
    round the GRADIENT to fp16 and flush small gradients to zero)"""
    return x + (x.half().float() - x).detach()


def torch_grid_encode(meta, table, x01):
    """table: (entries, 2) fp32 tensor holding fp16-representable values; x01 (n,3) -> (n, 32) fp32"""
    import torch
    n = x01.shape[0]
    feats = []
    for l in range(meta.n_levels):
        res, off = int(meta.res[l]), int(meta.offset[l])
        entries = int(meta.offset[l + 1]) - off
        hashed = (meta.hashed_mask >> l) & 1
        # fmaf(scale, x, 0.5): the double product of two floats is exact, so one rounding to float
        pos = (x01.double() * float(meta.scale[l]) + 0.5).float()
        g = torch.floor(pos)
        w = pos - g
        gi = g.to(torch.int64)
        acc = torch.zeros(n, 2, dtype=torch.float32, device=x01.device)
        for c in range(8):
            px = gi[:, 0] + (c & 1)
            py = gi[:, 1] + ((c >> 1) & 1)
            pz = gi[:, 2] + ((c >> 2) & 1)
            if hashed:
                idx = ((px & 0xFFFFFFFF) ^ ((py * 2654435761) & 0xFFFFFFFF) ^ ((pz * 805459861) & 0xFFFFFFFF)) % entries
            else:
                idx = ((px + py * res + pz * res * res) & 0xFFFFFFFF) % entries
            wt = (w[:, 0] if c & 1 else 1 - w[:, 0]) * (w[:, 1] if c & 2 else 1 - w[:, 1]) * (w[:, 2] if c & 4 else 1 - w[:, 2])
            acc = acc + wt[:, None] * table[off + idx]
        feats.append(_rt(acc))
    for l in range(meta.n_levels, 16):
        feats.append(torch.zeros(n, 2, device=x01.device))
    return torch.cat(feats, 1)


SH_C = [0.28209479177387814, 0.48860251190291987, 1.0925484305920792, 0.94617469575755997, 0.31539156525251999,
        0.54627421529603959, 0.59004358992664352, 2.8906114426405538, 0.45704579946446572, 0.3731763325901154,
        1.4453057213202769]


def torch_sh4(d):
    import torch
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z,
         -0.48860251190291987 * x, 1.0925484305920792 * xy, -1.0925484305920792 * yz,
         0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
         0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2),
         2.8906114426405538 * xy * z, 0.45704579946446572 * y * (1.0 - 5.0 * z2),
         0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
         1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(o, 1)


def torch_ngp_forward(meta, enc_params, rgb_params, xyz_min, xyz_max, xyzs, dirs, rgb_act=1):
    """fp32 torch restatement with fp16 rounding at: parameters, encoded features, every layer output.
    enc_params / rgb_params may require grad (fp32 master parameters). Returns sigmas, rgbs, h."""
    import torch
    encp = _rt(enc_params)
    rgbp = _rt(rgb_params)
    W1d = encp[:2048].view(64, 32)
    W2d = encp[2048:3072].view(16, 64)
    table = encp[3072:].view(-1, 2)
    x01 = (xyzs - xyz_min) / (xyz_max - xyz_min)
    feat = torch_grid_encode(meta, table, x01)
    hid = _rt(torch.relu(feat @ W1d.t()))
    h = _rt(hid @ W2d.t())
    # TruncExp (reference custom_functions.py:162-173): forward exp, backward exp(clamp(x,-15,15))
    h0 = h[:, 0]
    sig = torch.exp(h0.detach()) + (h0 - h0.detach()) * torch.exp(h0.detach().clamp(-15, 15))
    d = dirs / torch.norm(dirs, dim=1, keepdim=True)
    sh = _rt(torch_sh4(d))
    W1r = rgbp[:2048].view(64, 32)
    W2r = rgbp[2048:6144].view(64, 64)
    W3r = rgbp[6144:].view(16, 64)
    r1 = _rt(torch.relu(torch.cat([sh, h], 1) @ W1r.t()))
    r2 = _rt(torch.relu(r1 @ W2r.t()))
    out = (r2 @ W3r.t())[:, :3]
    if rgb_act == 1:
        out = torch.sigmoid(out)
    return sig, _rt(out), h
