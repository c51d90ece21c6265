"""GPU parity of the twelve vren operators (through the C ABI) against
  (1) the CPU oracle (oracle/ngp_oracle.c) on the same seeded inputs, and
  (2) the REAL reference kernels (oracle/_ref/vren) when they are present on the box.
Marcher: per-ray sample counts and the t / dt / xyz sequences are BIT-EXACT.
Compositing / distortion: 1e-4 relative (the reference uses __expf and serial fp32 sums).
"""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def T(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def rel_close(a, b, rtol=RTOL, atol=1e-6, what=""):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    err = np.abs(a - b)
    lim = atol + rtol * np.maximum(np.abs(a), np.abs(b))
    bad = err > lim
    assert not bad.any(), "%s: %d/%d outside rtol=%g; worst abs err %g at %s (%g vs %g)" % (
        what, bad.sum(), bad.size, rtol, err.max(), np.unravel_index(err.argmax(), err.shape),
        a.flat[err.argmax()], b.flat[err.argmax()])


def bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    bad = a != b
    assert not bad.any(), "%s: %d/%d elements differ bitwise, first at %s" % (
        what, bad.sum(), bad.size, np.argwhere(bad)[0] if bad.any() else None)


def my_aabb(c):
    from ngp_pl_b200 import vren
    center = torch.zeros(1, 3, device="cuda")
    half = torch.full((1, 3), float(c["scale"]), device="cuda")
    return vren.ray_aabb_intersect(T(c["o"]), T(c["d"]), center, half, 1)


def near_clamp(hits_t):
    t0 = hits_t[:, 0, 0]
    hits_t[:, 0, 0] = torch.where((t0 >= 0) & (t0 < 0.01), torch.full_like(t0, 0.01), t0)
    return hits_t[:, 0].contiguous()


@pytest.mark.parametrize("name", cases.MARCH_CASES)
def test_aabb_and_march_train_vs_oracle(name, oracle):
    from ngp_pl_b200 import vren
    c = cases.march_case(name)
    cnt, hits_t, idx = my_aabb(c)
    hits_o = cases.hits_for(c, oracle)
    hits = near_clamp(hits_t)
    bits_equal(hits.cpu().numpy(), hits_o, "hits_t")
    assert ((hits_o[:, 1] > 0) == (cnt.cpu().numpy() == 1)).all()

    rays_a, xyzs, dirs, deltas, ts, counter = vren.raymarching_train(
        T(c["o"]), T(c["d"]), hits, T(c["bits"]), c["cascades"], float(c["scale"]), float(c["esf"]), T(c["noise"]), 128, 1024)
    ra_o, xyz_o, dir_o, dl_o, ts_o = oracle.march_train(c["o"], c["d"], hits_o, c["bits"], c["cascades"], c["scale"],
                                                        c["esf"], c["noise"], 128, 1024)
    total = int(counter[0])
    assert int(counter[1]) == c["o"].shape[0]
    assert (rays_a.cpu().numpy() == ra_o).all(), "rays_a (ray_idx,start,N) differs from the oracle"
    assert total == ra_o[:, 2].sum()
    bits_equal(ts[:total].cpu().numpy(), ts_o, "ts")
    bits_equal(deltas[:total].cpu().numpy(), dl_o, "deltas")
    bits_equal(xyzs[:total].cpu().numpy(), xyz_o, "xyzs")
    bits_equal(dirs[:total].cpu().numpy(), dir_o, "dirs")
    if name == "full_scale2":
        assert ra_o[:, 2].max() == 1024 and (ra_o[:, 2] == 1024).sum() >= 4  # max_samples saturation is exercised


@pytest.mark.parametrize("name", cases.MARCH_CASES)
def test_march_train_vs_reference_kernels(name, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built on this box")
    from ngp_pl_b200 import vren
    c = cases.march_case(name)
    o, d, bits, noise = T(c["o"]), T(c["d"]), T(c["bits"]), T(c["noise"])
    center = torch.zeros(1, 3, device="cuda")
    half = torch.full((1, 3), float(c["scale"]), device="cuda")
    cnt_r, hits_r, _ = ref.vren.ray_aabb_intersect(o, d, center, half, 1)
    cnt_m, hits_m, _ = vren.ray_aabb_intersect(o, d, center, half, 1)
    bits_equal(hits_m.cpu().numpy(), hits_r.cpu().numpy(), "ray_aabb_intersect hits_t")
    assert (cnt_m == cnt_r).all()
    hits = near_clamp(hits_m)
    args = (o, d, hits, bits, c["cascades"], float(c["scale"]), float(c["esf"]), noise, 128, 1024)
    ra_r, xyz_r, dir_r, dl_r, ts_r, cnt_r = ref.vren.raymarching_train(*args)
    ra_m, xyz_m, dir_m, dl_m, ts_m, cnt_m2 = vren.raymarching_train(*args)
    assert int(cnt_r[0]) == int(cnt_m2[0])
    ra_r = ra_r.cpu().numpy()
    ra_r = ra_r[np.argsort(ra_r[:, 0], kind="stable")]
    ra_m = ra_m.cpu().numpy()
    assert (ra_r[:, 2] == ra_m[:, 2]).all(), "per-ray sample counts differ from the reference kernel"
    sel = np.concatenate([np.arange(s, s + n) for _, s, n in ra_r if n > 0]) if int(cnt_r[0]) else np.zeros(0, np.int64)
    tot = int(cnt_m2[0])
    bits_equal(ts_m[:tot].cpu().numpy(), ts_r.cpu().numpy()[sel], "ts vs reference")
    bits_equal(dl_m[:tot].cpu().numpy(), dl_r.cpu().numpy()[sel], "deltas vs reference")
    bits_equal(xyz_m[:tot].cpu().numpy(), xyz_r.cpu().numpy()[sel], "xyzs vs reference")


@pytest.mark.parametrize("name", cases.MARCH_CASES)
def test_march_test_rounds(name, oracle, ref):
    from ngp_pl_b200 import vren
    c = cases.march_case(name)
    o, d, bits = T(c["o"]), T(c["d"]), T(c["bits"])
    hits_np = cases.hits_for(c, oracle).copy()
    hits_m = T(hits_np)
    hits_r = hits_m.clone()
    n = o.shape[0]
    alive = torch.arange(n, device="cuda")
    for ns in (1, 2, 4, 64):
        out_m = vren.raymarching_test(o, d, hits_m, alive, bits, c["cascades"], float(c["scale"]), float(c["esf"]), 128, 1024, ns)
        out_o = oracle.march_test(c["o"], c["d"], hits_np, np.arange(n), c["bits"], c["cascades"], c["scale"], c["esf"],
                                  128, 1024, ns)
        for a, b, nm in zip(out_m[:4], out_o[:4], ["xyzs", "dirs", "deltas", "ts"]):
            bits_equal(a.cpu().numpy(), b, "raymarching_test %s (N_samples=%d)" % (nm, ns))
        assert (out_m[4].cpu().numpy() == out_o[4]).all()
        bits_equal(hits_m.cpu().numpy(), hits_np, "hits_t after round")
        if ref is not None:
            out_r = ref.vren.raymarching_test(o, d, hits_r, alive, bits, c["cascades"], float(c["scale"]), float(c["esf"]),
                                              128, 1024, ns)
            for a, b, nm in zip(out_m[:4], out_r[:4], ["xyzs", "dirs", "deltas", "ts"]):
                bits_equal(a.cpu().numpy(), b.cpu().numpy(), "raymarching_test %s vs reference" % nm)
            assert (out_m[4] == out_r[4]).all()
            bits_equal(hits_m.cpu().numpy(), hits_r.cpu().numpy(), "hits_t vs reference")


def test_march_large_vs_reference(ref):
    """bit-exactness over many rays (size-independent check at the bench's ray count and beyond)"""
    if ref is None:
        pytest.skip("oracle/_ref not built on this box")
    from ngp_pl_b200 import synth, vren
    for scene, esf, n_rays in ((synth.lego_scene(0), 0.0, 1 << 18), (synth.mip360_scene(0), 1.0 / 256, 1 << 16)):
        bits = T(synth.pack_bits(synth.occupancy_grid(scene)))
        o_np, d_np = cases.rays_from_scene(scene, n_rays, 77)
        o, d = T(o_np), T(d_np)
        center = torch.zeros(1, 3, device="cuda")
        half = torch.full((1, 3), scene.scale, device="cuda")
        _, hits_m, _ = vren.ray_aabb_intersect(o, d, center, half, 1)
        _, hits_r, _ = ref.vren.ray_aabb_intersect(o, d, center, half, 1)
        assert torch.equal(hits_m.view(torch.int32), hits_r.view(torch.int32))
        hits = near_clamp(hits_m)
        noise = torch.rand(n_rays, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
        args = (o, d, hits, bits, scene.cascades, scene.scale, esf, noise, 128, 1024)
        ra_r, xyz_r, _, dl_r, ts_r, cnt_r = ref.vren.raymarching_train(*args)
        ra_m, xyz_m, _, dl_m, ts_m, cnt_m = vren.raymarching_train(*args)
        tot = int(cnt_m[0])
        assert tot == int(cnt_r[0]) and tot > 0
        order = torch.argsort(ra_r[:, 0])
        ra_r = ra_r[order]
        assert torch.equal(ra_r[:, 2], ra_m[:, 2])
        # gather the reference's samples into ray order
        seg = torch.repeat_interleave(torch.arange(n_rays, device="cuda"), ra_m[:, 2])
        within = torch.arange(tot, device="cuda") - ra_m[:, 1][seg]
        src = ra_r[:, 1][seg] + within
        assert torch.equal(ts_m[:tot].view(torch.int32), ts_r[src].view(torch.int32))
        assert torch.equal(dl_m[:tot].view(torch.int32), dl_r[src].view(torch.int32))
        assert torch.equal(xyz_m[:tot].view(torch.int32), xyz_r[src].view(torch.int32))


def test_composite_train_fw_bw(oracle, ref):
    from ngp_pl_b200 import vren
    c = cases.composite_case()
    sig, rgbs, dl, ts, ra = T(c["sigmas"]), T(c["rgbs"]), T(c["deltas"]), T(c["ts"]), T(c["rays_a"])
    thr = float(c["T_thr"])
    total, opacity, depth, rgb, ws = vren.composite_train_fw(sig, rgbs, dl, ts, ra, thr)
    o_total, o_op, o_dp, o_rgb, o_ws = oracle.composite_train_fw(c["sigmas"], c["rgbs"], c["deltas"], c["ts"], c["rays_a"], thr)
    assert (total.cpu().numpy() == o_total).all()
    rel_close(opacity.cpu().numpy(), o_op, what="opacity")
    rel_close(depth.cpu().numpy(), o_dp, what="depth")
    rel_close(rgb.cpu().numpy(), o_rgb, what="rgb")
    rel_close(ws.cpu().numpy(), o_ws, atol=2e-6, what="ws")
    dsig, drgbs = vren.composite_train_bw(T(c["dO"]), T(c["dD"]), T(c["dC"]), T(c["dws"]), sig, rgbs, ws, dl, ts, ra,
                                          opacity, depth, rgb, thr)
    o_dsig, o_drgbs = oracle.composite_train_bw(c["dO"], c["dD"], c["dC"], c["dws"], c["sigmas"], c["rgbs"], o_ws, c["deltas"],
                                                c["ts"], c["rays_a"], o_op, o_dp, o_rgb, thr)
    rel_close(drgbs.cpu().numpy(), o_drgbs, atol=1e-5, what="dL_drgbs")
    # dL_dsigmas is a difference of O(1) terms scaled by delta: absolute floor = 1e-4 * delta * |terms|
    rel_close(dsig.cpu().numpy(), o_dsig, atol=1e-5, what="dL_dsigmas")
    if ref is not None:
        r_total, r_op, r_dp, r_rgb, r_ws = ref.vren.composite_train_fw(sig, rgbs, dl, ts, ra, thr)
        assert torch.equal(r_total, total)
        rel_close(opacity.cpu().numpy(), r_op.cpu().numpy(), what="opacity vs reference")
        rel_close(rgb.cpu().numpy(), r_rgb.cpu().numpy(), what="rgb vs reference")
        rel_close(depth.cpu().numpy(), r_dp.cpu().numpy(), what="depth vs reference")
        rel_close(ws.cpu().numpy(), r_ws.cpu().numpy(), atol=2e-6, what="ws vs reference")
        r_dsig, r_drgbs = ref.vren.composite_train_bw(T(c["dO"]), T(c["dD"]), T(c["dC"]), T(c["dws"]), sig, rgbs, r_ws, dl, ts,
                                                      ra, r_op, r_dp, r_rgb, thr)
        rel_close(drgbs.cpu().numpy(), r_drgbs.cpu().numpy(), atol=1e-5, what="dL_drgbs vs reference")
        rel_close(dsig.cpu().numpy(), r_dsig.cpu().numpy(), atol=1e-5, what="dL_dsigmas vs reference")


def test_composite_test_fw(oracle, ref):
    from ngp_pl_b200 import vren
    rng = np.random.RandomState(4)
    n_rays, n_alive, ns = 200, 150, 8
    alive_np = rng.permutation(n_rays)[:n_alive].astype(np.int64)
    sig = np.exp(rng.normal(0, 2.5, (n_alive, ns))).astype(np.float32)
    rgbs = rng.rand(n_alive, ns, 3).astype(np.float32)
    dl = np.full((n_alive, ns), 0.01, np.float32)
    ts = np.cumsum(dl, 1).astype(np.float32)
    neff = rng.randint(0, ns + 1, n_alive).astype(np.int32)
    op0 = (rng.rand(n_rays) * 0.9).astype(np.float32)
    dp0 = rng.rand(n_rays).astype(np.float32)
    rgb0 = rng.rand(n_rays, 3).astype(np.float32)
    alive_m, op_m, dp_m, rgb_m = T(alive_np), T(op0), T(dp0), T(rgb0)
    hits = torch.zeros(n_rays, 2, device="cuda")
    vren.composite_test_fw(T(sig), T(rgbs), T(dl), T(ts), hits, alive_m, 1e-2, T(neff), op_m, dp_m, rgb_m)
    alive_o, op_o, dp_o, rgb_o = alive_np.copy(), op0.copy(), dp0.copy(), rgb0.copy()
    oracle.composite_test_fw(sig, rgbs, dl, ts, alive_o, 1e-2, neff, op_o, dp_o, rgb_o)
    assert (alive_m.cpu().numpy() == alive_o).all()
    rel_close(op_m.cpu().numpy(), op_o, what="opacity")
    rel_close(dp_m.cpu().numpy(), dp_o, what="depth")
    rel_close(rgb_m.cpu().numpy(), rgb_o, what="rgb")
    if ref is not None:
        alive_r, op_r, dp_r, rgb_r = T(alive_np), T(op0), T(dp0), T(rgb0)
        ref.vren.composite_test_fw(T(sig), T(rgbs), T(dl), T(ts), hits, alive_r, 1e-2, T(neff), op_r, dp_r, rgb_r)
        assert torch.equal(alive_r, alive_m)
        rel_close(op_m.cpu().numpy(), op_r.cpu().numpy(), what="opacity vs reference")
        rel_close(rgb_m.cpu().numpy(), rgb_r.cpu().numpy(), what="rgb vs reference")


def test_distortion_loss(oracle, ref):
    from ngp_pl_b200 import vren
    c = cases.composite_case(seed=9)
    _, _, _, _, ws_np = oracle.composite_train_fw(c["sigmas"], c["rgbs"], c["deltas"], c["ts"], c["rays_a"], 1e-4)
    ws, dl, ts, ra = T(ws_np), T(c["deltas"]), T(c["ts"]), T(c["rays_a"])
    loss, wi, wti = vren.distortion_loss_fw(ws, dl, ts, ra)
    o_loss, o_wi, o_wti = oracle.distortion_fw(ws_np, c["deltas"], c["ts"], c["rays_a"])
    rel_close(wi.cpu().numpy(), o_wi, atol=1e-7, what="ws_inclusive_scan")
    rel_close(wti.cpu().numpy(), o_wti, atol=1e-7, what="wts_inclusive_scan")
    # the per-ray loss is a difference of O(1) prefix products: absolute, not relative, accuracy
    rel_close(loss.cpu().numpy(), o_loss, atol=3e-5, what="distortion loss")
    dL = np.random.RandomState(8).normal(size=ra.shape[0]).astype(np.float32)
    dws = vren.distortion_loss_bw(T(dL), wi, wti, ws, dl, ts, ra)
    o_dws = oracle.distortion_bw(dL, o_wi, o_wti, ws_np, c["deltas"], c["ts"], c["rays_a"])
    rel_close(dws.cpu().numpy(), o_dws, atol=3e-5, what="dL_dws")
    if ref is not None:
        r_loss, r_wi, r_wti = ref.vren.distortion_loss_fw(ws, dl, ts, ra)
        rel_close(loss.cpu().numpy(), r_loss.cpu().numpy(), atol=3e-5, what="distortion loss vs reference")
        r_dws = ref.vren.distortion_loss_bw(T(dL), r_wi, r_wti, ws, dl, ts, ra)
        rel_close(dws.cpu().numpy(), r_dws.cpu().numpy(), atol=3e-5, what="dL_dws vs reference")


def test_packbits_morton(oracle, ref):
    from ngp_pl_b200 import vren
    rng = np.random.RandomState(21)
    for dtype in (torch.float32, torch.float16, torch.float64):
        grid = rng.normal(0, 1, 4096 * 8).astype(np.float32)
        g = T(grid).to(dtype)
        bf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
        vren.packbits(g, 0.25, bf)
        want = oracle.packbits(g.float().cpu().numpy(), 0.25) if dtype != torch.float64 else oracle.packbits(grid, 0.25)
        assert (bf.cpu().numpy() == want).all()
        if ref is not None:
            bf2 = torch.zeros_like(bf)
            ref.vren.packbits(g, 0.25, bf2)
            assert torch.equal(bf, bf2)
    # odd size (not a multiple of 4 bytes) takes the byte path
    grid = rng.normal(0, 1, 1001 * 8).astype(np.float32)
    bf = torch.zeros(1001, dtype=torch.uint8, device="cuda")
    vren.packbits(T(grid), -0.1, bf)
    assert (bf.cpu().numpy() == oracle.packbits(grid, -0.1)).all()

    coords = rng.randint(0, 1024, (5000, 3)).astype(np.int32)
    coords[:3] = [[0, 0, 0], [1023, 1023, 1023], [127, 0, 64]]
    m = vren.morton3D(T(coords))
    assert (m.cpu().numpy() == oracle.morton3D(coords)).all()
    inv = vren.morton3D_invert(m)
    assert (inv.cpu().numpy() == coords).all()
    if ref is not None:
        c128 = T(rng.randint(0, 128, (5000, 3)).astype(np.int32))
        assert torch.equal(vren.morton3D(c128), ref.vren.morton3D(c128))
        assert torch.equal(vren.morton3D_invert(vren.morton3D(c128)), ref.vren.morton3D_invert(ref.vren.morton3D(c128)))


def test_empty_and_error_paths():
    from ngp_pl_b200 import vren
    e3 = torch.zeros(0, 3, device="cuda")
    cnt, hits, idx = vren.ray_aabb_intersect(e3, e3, torch.zeros(1, 3, device="cuda"), torch.ones(1, 3, device="cuda"), 1)
    assert hits.shape == (0, 1, 2)
    out = vren.raymarching_train(e3, e3, torch.zeros(0, 2, device="cuda"), torch.zeros(128 ** 3 // 8, dtype=torch.uint8, device="cuda"),
                                 1, 0.5, 0.0, torch.zeros(0, device="cuda"), 128, 1024)
    assert int(out[5][0]) == 0
    with pytest.raises(RuntimeError):
        vren.morton3D(torch.zeros(4, 3, dtype=torch.int32))  # CPU tensor -> RuntimeError, like TORCH_CHECK(is_cuda)
    with pytest.raises(RuntimeError):
        vren.morton3D(torch.zeros(4, 6, dtype=torch.int32, device="cuda")[:, ::2])  # non-contiguous
