"""The ALGORITHM of the CUDA marcher (ngp_pl_b200/csrc/march.cuh: march_ray_warp), replayed on the CPU.

The kernel does not walk a ray point by point as the reference does. A warp materialises 32 consecutive points of the ray's
step chain, probes all 32 cells at once, finds every empty point's jump target with a binary search over the chain, and gets
the set of points the serial walk visits as the orbit of the entry point under
    succ(j) = stop (past the box) | j + 1 (occupied) | first chain point not below t_target_j (empty)
by pointer jumping; a jump past the block is carried into the next block. This file restates exactly that control flow in
numpy (float32 adds = __fadd_rn; the per-point arithmetic comes from the C oracle's probe) and checks that it emits the SAME
samples as the oracle's literal serial loop -- including the constant-step chain guess + verification, rays that leave the
box inside a block, the carry across blocks and the budget-limited fallback (max_samples reached). Test infrastructure only.
"""
import numpy as np
import pytest

import cases

F = np.float32


def _chain(t, dt_of, const_dt):
    """32 chain points from t (and the 33rd): lane j holds c_j. For a constant step: guess + verify, else serial adds."""
    if const_dt is not None:
        inc = F(F(t + const_dt) - t)
        guess = (np.arange(32, dtype=F) * inc + t).astype(F)  # fma(j, inc, t): exact whenever the check below passes
        ok = np.all((guess[:-1] + const_dt).astype(F) == guess[1:])
        if ok:
            return guess, F(guess[31] + const_dt), True
    p = np.empty(32, F)
    cur = F(t)
    for j in range(32):
        p[j] = cur
        cur = F(cur + dt_of(cur))
    return p, cur, False


def _warp_march(o, d, t_start, t2, probe, dt_of, const_dt, max_new):
    """-> list of (t, dt) in ray order, the way march_ray_warp produces them"""
    out = []
    n = 0
    t = F(t_start)
    pending, skip_to = False, F(0)
    alive = (t >= 0) and (t < t2) and max_new > 0
    stats = {"blocks": 0, "guess_ok": 0, "fallback_blocks": 0}
    while alive:
        p, t_next, ok = _chain(t, dt_of, const_dt)
        stats["blocks"] += 1
        stats["guess_ok"] += int(ok)
        valid = p < t2
        occ, dts, tgt = probe(p)
        occ = occ.astype(bool) & valid
        # successor of every empty point: first lane with p >= t_target (binary search over the increasing chain)
        nxt = np.array([j + 1 + int(np.searchsorted(p[j + 1:], tgt[j], side="left")) for j in range(32)])
        assert np.all(np.diff(p) > 0)
        cur = 0
        if pending:
            m = np.nonzero(~(p < skip_to))[0]
            cur = int(m[0]) if m.size else 32
            pending = m.size == 0
        sample = np.zeros(32, bool)
        room0 = max_new - n
        if room0 > 32:
            succ = np.where(~valid, 32, np.where(occ, np.arange(32) + 1, nxt))
            visited = np.zeros(33, bool)
            if cur < 32:
                visited[cur] = True
            jump = succ.copy()
            for _ in range(5):  # pointer jumping: orbit of `cur`
                src = np.nonzero(visited[:32] & (jump < 32))[0]
                visited[jump[src]] = True
                jump = np.where(jump < 32, jump[np.minimum(jump, 31)], 32)
            visited = visited[:32]
            sample = visited & occ
            exits = np.nonzero(visited & ~valid)[0]
            if exits.size:
                alive = False
            elif visited.any():
                last = int(np.nonzero(visited)[0][-1])
                if not occ[last]:
                    pending, skip_to = True, tgt[last]
        else:  # the literal walk (the sample budget may run out inside this block)
            stats["fallback_blocks"] += 1
            while cur < 32:
                if not valid[cur]:
                    alive = False
                    break
                room = max_new - n - int(sample.sum())
                if room <= 0:
                    alive = False
                    break
                if occ[cur]:
                    run = 0
                    while cur + run < 32 and occ[cur + run]:
                        run += 1
                    run = min(run, room)
                    sample[cur:cur + run] = True
                    cur += run
                else:
                    to = int(nxt[cur])
                    if to >= 32:
                        pending, skip_to = True, tgt[cur]
                    cur = to
        for j in np.nonzero(sample)[0]:
            out.append((p[j], dts[j]))
        n += int(sample.sum())
        if alive:
            t = t_next
            if not (t < t2):
                alive = False
    return out, stats


@pytest.mark.parametrize("name", ["lego", "lego_half_random", "full_scale2", "mip360"])
def test_warp_walk_equals_serial_walk(name, oracle):
    c = cases.march_case(name)
    o, d = c["o"], c["d"]
    n_rays = min(o.shape[0], 48 if name != "lego" else 96)
    o, d, noise = o[:n_rays], d[:n_rays], c["noise"][:n_rays]
    sc = c["scene"]
    half = np.full((1, 3), sc.scale, np.float32)
    hits = oracle.ray_aabb(o, d, np.zeros((1, 3), np.float32), half, near_distance=float(cases.NEAR))
    rays_a, xyzs, dirs, deltas, ts = oracle.march_train(o, d, hits, c["bits"], c["cascades"], c["scale"], c["esf"], noise,
                                                        c["grid_size"], c["max_samples"])
    lo = F(np.sqrt(F(3)) if False else F(1.73205080757)) / F(c["max_samples"])
    hi = F(F(c["scale"]) * F(3.46410161514)) / F(c["grid_size"])
    esf = F(c["esf"])
    dt_of = lambda t: F(max(lo, min(F(t * esf), hi)))
    const_dt = lo if (esf == 0 and lo <= hi) else None
    tot_blocks = tot_ok = tot_fb = 0
    for r in range(n_rays):
        t1, t2 = F(hits[r, 0]), F(hits[r, 1])
        if t1 >= 0:
            t1 = F(np.float32(np.float64(dt_of(t1)) * np.float64(noise[r]) + np.float64(t1)))  # fmaf(dt, noise, t1): exact product fits double
        probe = lambda p, r=r: oracle.march_probe(o[r], d[r], c["bits"], c["cascades"], c["scale"], c["esf"], c["scale"],
                                                  c["grid_size"], c["max_samples"], p)
        got, st = _warp_march(o[r], d[r], t1, t2, probe, dt_of, const_dt, c["max_samples"])
        tot_blocks += st["blocks"]; tot_ok += st["guess_ok"]; tot_fb += st["fallback_blocks"]
        s0, cnt = int(rays_a[r, 1]), int(rays_a[r, 2])
        assert len(got) == cnt, (name, r, len(got), cnt)
        if cnt:
            gt = np.array([g[0] for g in got], F)
            gd = np.array([g[1] for g in got], F)
            assert np.array_equal(gt, ts[s0:s0 + cnt]) and np.array_equal(gd, deltas[s0:s0 + cnt]), (name, r)
    assert tot_blocks > 0
    if const_dt is not None:
        assert tot_ok > 0.5 * tot_blocks  # the verified closed form is the common case, the serial adds the exception
    if name == "full_scale2":
        assert tot_fb > 0  # rays that reach max_samples end in budget-limited blocks
