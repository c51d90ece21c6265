"""Host logic of ngp_pl_b200/models/custom_functions.py on the CPU: the autograd plumbing of the five Functions runs with
the operator module replaced by the C oracle (test infrastructure only -- the product path has no CPU route) and is checked
against torch autograd / finite differences."""
import types

import numpy as np
import torch

import cases


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a))


def _oracle_vren(max_samples_seen):
    from oracle import oracle as O

    def ray_aabb_intersect(rays_o, rays_d, center, half_size, max_hits):
        hits = O.ray_aabb(rays_o.numpy(), rays_d.numpy(), center.numpy(), half_size.numpy())
        n = hits.shape[0]
        cnt = (hits[:, 0] >= 0).astype(np.int32)
        return [_t(cnt), _t(hits.reshape(n, 1, 2)), torch.zeros(n, 1, dtype=torch.int64)]

    def raymarching_train(rays_o, rays_d, hits_t, bits, cascades, scale, esf, noise, grid_size, max_samples):
        max_samples_seen.append(max_samples)
        rays_a, xyzs, dirs, deltas, ts = O.march_train(rays_o.numpy(), rays_d.numpy(), hits_t.numpy(), bits.numpy(), cascades,
                                                        scale, esf, noise.numpy(), grid_size, max_samples)
        tot = xyzs.shape[0]
        pad = 7  # the operator returns capacity-sized tensors; only the first counter[0] rows are defined
        grow = lambda a: np.concatenate([a, np.full((pad,) + a.shape[1:], np.nan, a.dtype)])
        return [_t(rays_a), _t(grow(xyzs)), _t(grow(dirs)), _t(grow(deltas)), _t(grow(ts)), torch.tensor([tot, rays_a.shape[0]], dtype=torch.int32)]

    def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T):
        return [_t(x) for x in O.composite_train_fw(sigmas.numpy(), rgbs.numpy(), deltas.numpy(), ts.numpy(), rays_a.numpy(), T)]

    def composite_train_bw(dO, dD, dC, dws, sigmas, rgbs, ws, deltas, ts, rays_a, opacity, depth, rgb, T):
        return [_t(x) for x in O.composite_train_bw(dO.numpy(), dD.numpy(), dC.numpy(), dws.numpy(), sigmas.numpy(), rgbs.numpy(),
                                                    ws.numpy(), deltas.numpy(), ts.numpy(), rays_a.numpy(), opacity.numpy(),
                                                    depth.numpy(), rgb.numpy(), T)]
    return types.SimpleNamespace(ray_aabb_intersect=ray_aabb_intersect, raymarching_train=raymarching_train,
                                 composite_train_fw=composite_train_fw, composite_train_bw=composite_train_bw)


def test_function_wrappers_on_the_oracle(monkeypatch):
    from ngp_pl_b200 import synth
    from ngp_pl_b200.models import custom_functions as cf
    seen = []
    monkeypatch.setattr(cf, "vren", _oracle_vren(seen))
    scene = synth.lego_scene(0)
    bits = _t(synth.pack_bits(synth.occupancy_grid(scene)))
    n = 96
    o_np, d_np = cases.rays_from_scene(scene, n, 5, extra_edge_cases=True)
    o = _t(o_np).requires_grad_(True)
    d = _t(d_np).requires_grad_(True)
    center, half = torch.zeros(1, 3), torch.full((1, 3), float(scene.scale))

    assert cf.RayAABBIntersector.__name__ == "RayAABBIntersector" and cf.RaySphereIntersector.__name__ == "RaySphereIntersector"
    cnt, hits_t, idx = cf.RayAABBIntersector.apply(o, d, center, half, 1)
    assert hits_t.shape == (n, 1, 2) and cnt.shape == (n,) and idx.shape == (n, 1)
    hits = hits_t[:, 0].detach().contiguous()

    noise = torch.rand(n, generator=torch.Generator().manual_seed(3))
    cf.RayMarcher.noise_override = noise
    try:
        rays_a, xyzs, dirs, deltas, ts, total = cf.RayMarcher.apply(o, d, hits, bits, 1, float(scene.scale), 0.0, 128, 1024)
    finally:
        cf.RayMarcher.noise_override = None
    S = int(total)
    assert seen == [1024] and S > 0 and int(rays_a[:, 2].sum()) == S
    assert xyzs.shape == (S, 3) and ts.shape == (S,) and not torch.isnan(xyzs).any()  # sliced to the defined rows
    # backward of the marcher: x = o + t d, dirs = d  (reference custom_functions.py:102-112)
    gx, gd = torch.randn(S, 3, generator=torch.Generator().manual_seed(4)), torch.randn(S, 3, generator=torch.Generator().manual_seed(5))
    (xyzs * gx).sum().backward(retain_graph=True)
    (dirs * gd).sum().backward()
    owner = torch.repeat_interleave(torch.arange(n), rays_a[:, 2])
    exp_o = torch.zeros(n, 3).index_add_(0, owner, gx)
    exp_d = torch.zeros(n, 3).index_add_(0, owner, gx * ts.detach()[:, None] + gd)
    assert torch.allclose(o.grad, exp_o, atol=1e-5) and torch.allclose(d.grad, exp_d, atol=1e-5)

    # compositing: autograd through VolumeRenderer against central differences of the oracle forward
    g = torch.Generator().manual_seed(6)
    sig = (torch.rand(S, generator=g) * 30).requires_grad_(True)
    rgbs = torch.rand(S, 3, generator=g).requires_grad_(True)
    dl, tt = deltas.detach(), ts.detach()

    def loss_of(s_, c_):
        tot_, opacity, depth, rgb, ws = cf.VolumeRenderer.apply(s_, c_, dl, tt, rays_a, 1e-4)
        w_o, w_d = torch.linspace(0.5, 1.5, n), torch.linspace(1.0, 2.0, n)
        return (opacity * w_o).sum() + (depth * w_d).sum() + (rgb * torch.tensor([1.0, 2.0, 3.0])).sum() + (ws * 0.1).sum(), tot_
    L, tot_ = loss_of(sig, rgbs)
    assert tot_.dim() == 0 and 0 < int(tot_) <= S
    L.backward()
    k = int(torch.argmax(rays_a[:, 2]))  # the longest ray
    s0 = int(rays_a[k, 1])
    for j in (s0, s0 + 1, s0 + 2):
        for tensor, grad, eps in ((sig, sig.grad, 1e-2), (rgbs, rgbs.grad, 1e-2)):
            base = tensor.detach().clone()
            idx = (j,) if tensor is sig else (j, 1)
            hi, lo = base.clone(), base.clone()
            hi[idx] += eps
            lo[idx] -= eps
            with torch.no_grad():
                fd = (loss_of(*((hi, rgbs.detach()) if tensor is sig else (sig.detach(), hi)))[0]
                      - loss_of(*((lo, rgbs.detach()) if tensor is sig else (sig.detach(), lo)))[0]) / (2 * eps)
            assert abs(float(fd) - float(grad[idx])) < 2e-2 * max(1.0, abs(float(fd))), (idx, float(fd), float(grad[idx]))

    # TruncExp: exp forward, derivative evaluated at the clamped argument
    x = torch.tensor([-20.0, -1.0, 0.5, 14.0, 16.0, 40.0], requires_grad=True)
    y = cf.TruncExp.apply(x)
    assert torch.equal(y, torch.exp(x.detach()))
    y.backward(torch.ones_like(y))
    assert torch.allclose(x.grad, torch.exp(x.detach().clamp(-15, 15)))
