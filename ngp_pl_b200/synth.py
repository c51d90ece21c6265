"""Seeded synthetic scenes shaped like the reference's datasets (no dataset ships with the box).

SynthLego: a union of axis-aligned coloured boxes inside [-0.5, 0.5]^3, about
8-9 % of the 128^3 occupancy cells solid, ray-traced analytically against a white background, seen by
cameras on a sphere of radius 1.5 looking at the origin with the Synthetic-NeRF intrinsics
(800x800, fx = fy = 1111.11; reference datasets/nerf.py:27,71-72). Ray convention is the reference's
(datasets/ray_utils.py:33-35,60-68): directions ((u-cx+0.5)/fx, (v-cy+0.5)/fy, 1) UNNORMALISED,
rays_d = dirs @ R^T, rays_o = c2w[:, 3].

SynthMip360: scale 16 (6 cascades), content in the central unit cube plus a sparse far shell, black
background, exp_step_factor = 1/256.
"""
import numpy as np
import torch


class Scene:
    def __init__(self, box_min, box_max, colors, scale, bg):
        self.box_min = box_min  # (B,3) float32
        self.box_max = box_max
        self.colors = colors    # (B,3)
        self.scale = float(scale)
        self.bg = float(bg)

    @property
    def cascades(self):
        return max(1 + int(np.ceil(np.log2(2 * self.scale))), 1)

    @property
    def exp_step_factor(self):
        return 1.0 / 256 if self.scale > 0.5 else 0.0


def lego_scene(seed=0, n_boxes=20):
    """Lego-shaped: a plate, a chassis and studs/bricks; deterministic for a given seed."""
    rng = np.random.RandomState(seed)
    mins, maxs = [], []
    # base plate and a chassis block
    mins.append([-0.34, -0.34, -0.30]); maxs.append([0.34, 0.34, -0.26])
    mins.append([-0.22, -0.10, -0.26]); maxs.append([0.22, 0.10, -0.10])
    while len(mins) < n_boxes:
        c = rng.uniform(-0.33, 0.33, 3)
        c[2] = rng.uniform(-0.24, 0.30)
        h = rng.uniform(0.02, 0.085, 3)
        lo, hi = c - h, c + h
        if np.all(lo > -0.46) and np.all(hi < 0.46):
            mins.append(lo.tolist()); maxs.append(hi.tolist())
    colors = rng.uniform(0.15, 0.95, (len(mins), 3))
    return Scene(np.asarray(mins, np.float32), np.asarray(maxs, np.float32), colors.astype(np.float32), 0.5, 1.0)


def mip360_scene(seed=0, n_center=16, n_far=40):
    rng = np.random.RandomState(seed + 1000)
    mins, maxs = [], []
    for _ in range(n_center):
        c = rng.uniform(-0.4, 0.4, 3); h = rng.uniform(0.04, 0.15, 3)
        mins.append(c - h); maxs.append(c + h)
    for _ in range(n_far):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        c = d * rng.uniform(5.0, 13.0); h = rng.uniform(0.3, 1.2, 3)
        lo, hi = np.maximum(c - h, -15.5), np.minimum(c + h, 15.5)
        mins.append(lo); maxs.append(hi)
    colors = rng.uniform(0.1, 0.9, (len(mins), 3))
    return Scene(np.asarray(mins, np.float32), np.asarray(maxs, np.float32), colors.astype(np.float32), 16.0, 0.0)


def _morton3d_np(x, y, z):
    def spread(v):
        v = v.astype(np.uint32) & 0x3ff
        v = (v | (v << 16)) & 0x030000ff
        v = (v | (v << 8)) & 0x0300f00f
        v = (v | (v << 4)) & 0x030c30c3
        v = (v | (v << 2)) & 0x09249249
        return v
    return spread(x) | (spread(y) << 1) | (spread(z) << 2)


def occupancy_grid(scene, grid_size=128, dilate=1):
    """(cascades, G^3) float32 in Morton order: 1.0 where a cell of cascade k overlaps a box
    (dilated by `dilate` cells), else 0. Cascade k spans [-min(2^(k-1), scale), +...]^3
    (reference networks.py:25,251)."""
    G = grid_size
    out = np.zeros((scene.cascades, G ** 3), np.float32)
    for c in range(scene.cascades):
        s = min(2.0 ** (c - 1), scene.scale)
        cell = 2 * s / G
        occ = np.zeros((G, G, G), bool)
        for lo, hi in zip(scene.box_min, scene.box_max):
            i0 = np.clip(np.floor((lo + s) / cell).astype(int) - dilate, 0, G - 1)
            i1 = np.clip(np.floor((hi + s) / cell).astype(int) + dilate, 0, G - 1)
            if np.any(hi < -s) or np.any(lo > s):
                continue
            occ[i0[0]:i1[0] + 1, i0[1]:i1[1] + 1, i0[2]:i1[2] + 1] = True
        ix, iy, iz = np.nonzero(occ)
        out[c, _morton3d_np(ix, iy, iz)] = 1.0
    return out


def pack_bits(grid, thr=0.5):
    g = (grid.reshape(-1, 8) > thr).astype(np.uint8)
    return (g << np.arange(8, dtype=np.uint8)).sum(1).astype(np.uint8)


def intrinsics(W=800, H=800, fx=1111.11, fy=None):
    fy = fx if fy is None else fy
    return dict(W=W, H=H, fx=fx, fy=fy, cx=W / 2, cy=H / 2)


def ray_directions(K, device="cpu"):
    """(H*W, 3) camera-frame directions, unnormalised (reference datasets/ray_utils.py:33-35)."""
    v, u = torch.meshgrid(torch.arange(K["H"], dtype=torch.float32, device=device),
                          torch.arange(K["W"], dtype=torch.float32, device=device), indexing="ij")
    d = torch.stack([(u - K["cx"] + 0.5) / K["fx"], (v - K["cy"] + 0.5) / K["fy"], torch.ones_like(u)], -1)
    return d.reshape(-1, 3)


def camera_poses(n, radius=1.5, seed=0, upper_only=True):
    """(n,3,4) camera-to-world, camera axes [right, down, forward], looking at the origin."""
    rng = np.random.RandomState(seed + 7)
    poses = np.zeros((n, 3, 4), np.float32)
    for i in range(n):
        d = rng.normal(size=3)
        if upper_only:
            d[2] = abs(d[2]) * 0.8 + 0.1
        d /= np.linalg.norm(d)
        pos = d * radius
        fwd = -d
        up = np.array([0, 0, 1.0])
        right = np.cross(fwd, up); right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        poses[i, :, 0], poses[i, :, 1], poses[i, :, 2], poses[i, :, 3] = right, down, fwd, pos
    return poses


def camera_radius(scene):
    return 1.5 if scene.scale <= 0.5 else 2.5


def get_rays(directions, c2w):
    """directions (N,3), c2w (3,4) or (N,3,4) -> rays_o, rays_d (reference datasets/ray_utils.py:46-70)."""
    if c2w.dim() == 2:
        rays_d = directions @ c2w[:, :3].T
    else:
        rays_d = torch.einsum("nc,nkc->nk", directions, c2w[..., :3])
    rays_o = c2w[..., 3].expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()


@torch.no_grad()
def trace(scene, rays_o, rays_d, chunk=1 << 20):
    """Analytic ground truth: colour of the first box hit, Lambert-ish shading by face, else bg."""
    dev = rays_o.device
    bmin = torch.as_tensor(scene.box_min, device=dev)
    bmax = torch.as_tensor(scene.box_max, device=dev)
    col = torch.as_tensor(scene.colors, device=dev)
    shade = torch.tensor([0.75, 0.9, 1.0], device=dev)
    out = torch.empty(rays_o.shape[0], 3, device=dev)
    for i in range(0, rays_o.shape[0], chunk):
        o, d = rays_o[i:i + chunk, None], rays_d[i:i + chunk, None]
        inv = 1.0 / d
        ta, tb = (bmin - o) * inv, (bmax - o) * inv
        tlo, thi = torch.minimum(ta, tb), torch.maximum(ta, tb)
        t1, axis = tlo.max(-1)
        t2 = thi.min(-1)[0]
        hit = (t1 <= t2) & (t2 > 0) & (t1 > 0)
        t1 = torch.where(hit, t1, torch.full_like(t1, float("inf")))
        tmin, which = t1.min(-1)
        any_hit = torch.isfinite(tmin)
        c = col[which] * shade[axis.gather(1, which[:, None])[:, 0]][:, None]
        out[i:i + chunk] = torch.where(any_hit[:, None], c, torch.full_like(c, scene.bg))
    return out


class RayBank:
    """Training images of a synthetic scene as flat device tensors + the reference's per-step random
    sampling of (image, pixel) pairs with replacement (reference datasets/base.py:22-30)."""

    def __init__(self, scene, n_images=100, K=None, device="cuda", seed=0, store_dtype=torch.uint8, radius=None):
        self.scene = scene
        self.K = K or intrinsics()
        self.device = device
        self.directions = ray_directions(self.K, device)
        # bounded scenes: cameras on the radius-1.5 upper hemisphere (Synthetic-NeRF); unbounded (scale > 0.5): inside the
        # far shell, all around the central content
        radius = camera_radius(scene) if radius is None else radius
        self.poses = torch.as_tensor(camera_poses(n_images, radius=radius, seed=seed, upper_only=scene.scale <= 0.5),
                                     device=device)
        n_pix = self.directions.shape[0]
        self.rgb = torch.empty(n_images, n_pix, 3, device=device, dtype=store_dtype)
        for i in range(n_images):
            o, d = get_rays(self.directions, self.poses[i])
            c = trace(scene, o, d)
            self.rgb[i] = (c * 255).round().to(torch.uint8) if store_dtype == torch.uint8 else c
        self.gen = torch.Generator(device=device).manual_seed(seed + 99)

    def sample(self, batch_size):
        img = torch.randint(self.poses.shape[0], (batch_size,), device=self.device, generator=self.gen)
        pix = torch.randint(self.directions.shape[0], (batch_size,), device=self.device, generator=self.gen)
        rays_o, rays_d = get_rays(self.directions[pix], self.poses[img])
        rgb = self.rgb[img, pix]
        rgb = rgb.float() / 255 if rgb.dtype == torch.uint8 else rgb
        return rays_o, rays_d, rgb
