"""Host logic of ngp_pl_b200/losses.py on the CPU: the autograd plumbing of DistortionLoss / NeRFLoss is exercised with
the operator module replaced by the C oracle (test infrastructure only -- the product path has no CPU route), and checked
against plain torch autograd of the loss written out per ray."""
import types

import numpy as np
import torch

import cases  # noqa: F401  (puts the repo root on sys.path)


def _oracle_vren():
    from oracle import oracle as O

    def fw(ws, deltas, ts, rays_a):
        loss, ws_inc, wts_inc = O.distortion_fw(ws.detach().numpy(), deltas.numpy(), ts.numpy(), rays_a.numpy())
        return [torch.as_tensor(loss), torch.as_tensor(ws_inc), torch.as_tensor(wts_inc)]

    def bw(dL, ws_inc, wts_inc, ws, deltas, ts, rays_a):
        return torch.as_tensor(O.distortion_bw(dL.numpy(), ws_inc.numpy(), wts_inc.numpy(), ws.detach().numpy(), deltas.numpy(),
                                               ts.numpy(), rays_a.numpy()))
    return types.SimpleNamespace(distortion_loss_fw=fw, distortion_loss_bw=bw)


def _distortion_torch(ws, deltas, ts, rays_a):
    """Mip-NeRF 360 distortion loss per ray, O(n^2) definition: sum_ij w_i w_j |t_i - t_j| + 1/3 sum_i w_i^2 delta_i"""
    out = []
    for _, start, n in rays_a.tolist():
        w, t, d = ws[start:start + n], ts[start:start + n], deltas[start:start + n]
        out.append((w[:, None] * w[None, :] * (t[:, None] - t[None, :]).abs()).sum() + (w * w * d).sum() / 3)
    return torch.stack(out)


def test_losses_module_plumbing_against_torch_autograd(monkeypatch):
    from ngp_pl_b200 import losses
    monkeypatch.setattr(losses, "vren", _oracle_vren())
    rng = np.random.RandomState(0)
    counts = np.array([0, 5, 1, 17, 40, 3], dtype=np.int64)  # ragged, incl. an empty ray
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    rays_a = torch.as_tensor(np.stack([np.arange(len(counts)), starts, counts], 1))
    S = int(counts.sum())
    ts = torch.as_tensor(np.concatenate([np.sort(rng.uniform(0.1, 2.0, c)) for c in counts]).astype(np.float32))
    deltas = torch.as_tensor(rng.uniform(1e-3, 2e-3, S).astype(np.float32))
    ws = torch.as_tensor(rng.uniform(0, 0.2, S).astype(np.float32)).requires_grad_(True)
    res = {"rgb": torch.rand(len(counts), 3), "opacity": torch.rand(len(counts)), "ws": ws, "deltas": deltas, "ts": ts,
           "rays_a": rays_a}
    tgt = {"rgb": torch.rand(len(counts), 3)}
    terms = losses.NeRFLoss(lambda_opacity=1e-3, lambda_distortion=1e-2)(res, tgt)
    assert set(terms) == {"rgb", "opacity", "distortion"}
    assert torch.equal(terms["rgb"], (res["rgb"] - tgt["rgb"]) ** 2)
    o = res["opacity"] + 1e-10
    assert torch.allclose(terms["opacity"], 1e-3 * (-o * torch.log(o)))
    ws2 = ws.detach().clone().requires_grad_(True)
    ref = 1e-2 * _distortion_torch(ws2, deltas, ts, rays_a)
    assert torch.allclose(terms["distortion"], ref, rtol=1e-4, atol=1e-7)
    up = torch.rand(len(counts))
    (terms["distortion"] * up).sum().backward()
    (ref * up).sum().backward()
    assert torch.allclose(ws.grad, ws2.grad, rtol=1e-3, atol=1e-7)
    # without the distortion weight the term (and the operator call) disappears
    assert set(losses.NeRFLoss(lambda_distortion=0)(res, tgt)) == {"rgb", "opacity"}
