"""The ALGORITHM of the fused compositing kernel (ngp_pl_b200/csrc/train.cu: k_train_composite_loss), replayed on the CPU.

The reference composites a ray serially, sample after sample (volumerendering.cu:6-45 forward, :87-151 backward), in three
launches with the loss in between. The kernel gives a ray to one warp and does forward, NeRFLoss and backward in one pass:
  * 32 samples per trip; the transmittance is a warp-wide inclusive product scan (Hillis-Steele, 5 shuffle steps) times a
    carry from the previous trip; the accumulated colour / depth / opacity are per-lane fma accumulators reduced at the end;
  * the ray stops after the trip in which T drops to <= T_threshold; the sample on which that happens IS composited, later
    ones are not, and total_samples excludes it;
  * the loss gradient of a ray needs only that ray's composited colour / opacity, so the backward follows at once: the same
    trips again with the inclusive colour prefixes (three add scans + carries) giving every sample its "colour behind me";
  * the first CL_CACHE trips are read once and held in registers for both sweeps, longer rays continue from memory.
This file restates exactly that control flow in numpy float32 (scans in the kernel's association order) and holds it to the
C oracle's literal serial loops + the NeRFLoss formulas, on ragged rays that include empty, 1-, 32-, 33- and 700-sample rays
(beyond any cache depth), opaque first samples and rays that never terminate. Test infrastructure only.
"""
import numpy as np
import pytest

import cases

F = np.float32
DSIG_ATOL = 1e-9  # |dL/dsigma| is ~1e-4 here (delta ~ 1e-2 times a 1/n_rays gradient); the two agree to ~7e-11


def _scan(v, op):
    """inclusive Hillis-Steele scan over 32 lanes, the association order of warp_scan_mul / warp_scan_add (common.cuh)"""
    v = v.astype(F).copy()
    o = 1
    while o < 32:
        u = np.concatenate([np.ones(o, F) if op == "mul" else np.zeros(o, F), v[:-o]])
        v = (v * u).astype(F) if op == "mul" else (v + u).astype(F)
        o <<= 1
    return v


def _butterfly_sum(v):
    v = v.astype(F).copy()
    o = 16
    while o > 0:
        v = (v + v[np.arange(32) ^ o]).astype(F)
        o >>= 1
    return v[0]


def _trip_inputs(sg, de, tt, cl, base, n):
    """lane-wise inputs of one trip; lanes past the ray's end are invalid (a = 0)"""
    idx = base + np.arange(32)
    valid = idx < n
    j = np.where(valid, idx, 0)
    a = np.where(valid, (F(1) - np.exp(-(sg[j] * de[j]).astype(F)).astype(F)).astype(F), F(0)).astype(F)
    return idx, valid, j, a


def _warp_composite_loss(sg, de, tt, cl, gt, bg, n_rays, lam, thr, cache):
    """one ray the way a warp of k_train_composite_loss does it -> (rgb, opacity, depth, total_samples, dsig, dcol, n_comp)"""
    n = len(sg)
    # ---- forward sweep ----
    acc = np.zeros((5, 32), F)  # r g b depth opacity, per lane
    T_carry, n_comp, done, base = F(1), 0, False, 0
    trips = 0
    while base < n and not done:
        idx, valid, j, a = _trip_inputs(sg, de, tt, cl, base, n)
        T_inc = (_scan(F(1) - a, "mul") * T_carry).astype(F)
        T_exc = np.concatenate([[T_carry], T_inc[:-1]]).astype(F)
        comp = valid & ((idx == 0) | (T_exc > thr))
        w = np.where(comp, (a * T_exc).astype(F), F(0)).astype(F)
        for k in range(3):
            acc[k] = np.where(comp, (w * cl[j, k] + acc[k]).astype(F), acc[k])
        acc[3] = np.where(comp, (w * tt[j] + acc[3]).astype(F), acc[3])
        acc[4] = np.where(comp, (acc[4] + w).astype(F), acc[4])
        n_comp += int(comp.sum())
        done = bool((valid & ~(T_inc > thr)).any())
        T_carry = T_inc[31]
        base += 32
        trips += 1
    C = np.array([_butterfly_sum(acc[k]) for k in range(3)], F)
    D, O = _butterfly_sum(acc[3]), _butterfly_sum(acc[4])
    total = n_comp - 1 if done else n_comp
    rest = F(1) - O
    out = (C + bg * rest).astype(F)
    # ---- NeRFLoss (losses.py:47-60, lambda_distortion = 0) and its per-ray gradients ----
    inv_n = F(1.0 / n_rays)
    e = (out - gt).astype(F)
    dC = (F(2) * e * inv_n * F(1.0 / 3.0)).astype(F)
    lg = np.log(O + F(1e-10)).astype(F)
    dsig, dcol = np.zeros(n, F), np.zeros((n, 3), F)
    n_comp_b = 0
    if n > 0:
        dO = F(lam * (-lg - F(1)) * inv_n - (dC * bg).sum(dtype=F))
        dO_term = F(dO * (F(1) - O))
        # ---- backward sweep: the same trips (the first `cache` of them come from registers in the kernel) ----
        T_carry, pref, done, base = F(1), np.zeros(3, F), False, 0
        while base < n and not done:
            idx, valid, j, a = _trip_inputs(sg, de, tt, cl, base, n)
            T_inc = (_scan(F(1) - a, "mul") * T_carry).astype(F)
            T_exc = np.concatenate([[T_carry], T_inc[:-1]]).astype(F)
            comp = valid & ((idx == 0) | (T_exc > thr))
            w = np.where(comp, (a * T_exc).astype(F), F(0)).astype(F)
            inc = [(_scan((w * cl[j, k]).astype(F), "add") + pref[k]).astype(F) for k in range(3)]
            g = (dC[0] * (cl[j, 0] * T_inc - (C[0] - inc[0])) + dC[1] * (cl[j, 1] * T_inc - (C[1] - inc[1])) +
                 dC[2] * (cl[j, 2] * T_inc - (C[2] - inc[2])) + dO_term).astype(F)
            sel = np.where(comp)[0]
            dsig[idx[sel]] = (de[j[sel]] * g[sel]).astype(F)
            dcol[idx[sel]] = (dC[None, :] * w[sel, None]).astype(F)
            n_comp_b += int(comp.sum())
            done = bool((valid & ~(T_inc > thr)).any())
            T_carry = T_inc[31]
            pref = np.array([inc[k][31] for k in range(3)], F)
            base += 32
    assert n_comp_b == (n_comp if n > 0 else 0)
    return out, O, D, total, dsig, dcol, n_comp, trips


@pytest.mark.parametrize("seed,bg", [(3, 1.0), (11, 0.0)])
def test_warp_composite_loss_equals_serial_reference(seed, bg, oracle):
    c = cases.composite_case(seed=seed, n_rays=64, max_n=300)
    rng = np.random.RandomState(seed + 100)
    ra = c["rays_a"]
    n_rays = ra.shape[0]
    # a few special rays: an opaque first sample, a fully transparent ray, a ray that terminates exactly at a trip boundary
    sig = c["sigmas"].copy()
    r_opaque, r_clear = ra[5], ra[6]
    sig[r_opaque[1]:r_opaque[1] + 1] = 1e9
    sig[r_clear[1]:r_clear[1] + r_clear[2]] = 0
    gt = rng.rand(n_rays, 3).astype(F)
    bgv = np.full(3, bg, F)
    lam, thr = 1e-3, c["T_thr"]
    # ---- the reference's way: serial forward, loss gradients, serial backward (the C oracle) ----
    total, opacity, depth, rgb, ws = oracle.composite_train_fw(sig, c["rgbs"], c["deltas"], c["ts"], ra, thr)
    out_ref = rgb + bgv[None, :] * (1 - opacity[:, None])
    dC = (2 * (out_ref - gt) / (3 * n_rays)).astype(F)
    dO = (lam * (-np.log(opacity + 1e-10) - 1) / n_rays - (dC * bgv).sum(1)).astype(F)
    dsig_ref, dcol_ref = oracle.composite_train_bw(dO, np.zeros(n_rays, F), dC, np.zeros_like(sig), sig, c["rgbs"], ws, c["deltas"],
                                                   c["ts"], ra, opacity, depth, rgb, thr)
    # ---- the kernel's way, ray by ray (rays_a row i is ray rays_a[i, 0]) ----
    long_rays = 0
    for ray, s, n in ra:
        sl = slice(s, s + n)
        out, O, D, tot, dsig, dcol, n_comp, trips = _warp_composite_loss(sig[sl], c["deltas"][sl], c["ts"][sl], c["rgbs"][sl], gt[ray],
                                                                          bgv, n_rays, lam, thr, cache=4)
        long_rays += trips > 4
        assert tot == total[ray], "total_samples of ray %d" % ray
        np.testing.assert_allclose(out, out_ref[ray], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(O, opacity[ray], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(D, depth[ray], rtol=1e-4, atol=1e-6)
        # gradients: absolute bars as in test_oracle_cpu (a tiny weight has a large RELATIVE exp error), scaled to 1/n_rays
        np.testing.assert_allclose(dcol, dcol_ref[sl], rtol=1e-4, atol=1e-5 / n_rays)
        np.testing.assert_allclose(dsig, dsig_ref[sl], rtol=2e-4, atol=DSIG_ATOL)
        # samples past the terminating one receive exactly zero (the kernel leaves them out of the live list)
        assert not dsig[n_comp:].any() and not dcol[n_comp:].any()
    assert long_rays >= 1, "the case must exercise rays longer than the register cache"
    assert total[r_opaque[0]] == 0 and total[r_clear[0]] == r_clear[2]
