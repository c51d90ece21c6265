"""GPU parity of the fused NGP network kernels (hash grid + MLPs + SH) against the fp32 oracle
restatement with identical fp16 rounding points (oracle/oracle.py: torch_ngp_forward on CPU, and the
C restatement oracle_ngp_forward).

tinycudann is absent from /root/reference, so this part of the path is "parity unpinned" against the
reference; what is pinned here is the CUDA kernel against the restated algorithm:
  features / h / rgb : fp16-level error (a few fp16 ulps; accumulation order differs)
  sigmas             : exp of an fp16 value -> same relative error as h0's absolute error
  gradients          : compared with autograd of the oracle, normalised by the gradient's max
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def build_model(scale=0.5, seed=0, table_amp=0.5, **kw):
    from ngp_pl_b200.models.networks import NGP
    m = NGP(scale, **kw).cuda()
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        # a trained-looking table: O(1) features instead of the 1e-4 init, so every level matters
        p = m.xyz_encoder.params
        p[3072:] = (torch.rand(p.numel() - 3072, generator=g) * 2 - 1).cuda() * table_amp
    return m


def oracle_inputs(model):
    enc = model.xyz_encoder.params.detach().cpu()
    rgbp = model.rgb_net.params.detach().cpu()
    return enc, rgbp


def sample_points(n, scale, seed):
    rng = np.random.RandomState(seed)
    x = rng.uniform(-scale, scale, (n, 3)).astype(np.float32)
    x[0] = [-scale, -scale, -scale]
    x[1] = [scale, scale, scale]  # upper boundary: the wrapping +1 corner of the dense levels
    x[2] = [0, 0, 0]
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d *= rng.uniform(0.5, 2.0, (n, 1)).astype(np.float32)  # unnormalised, as the marcher delivers them
    return x, d


@pytest.mark.parametrize("cfg", [dict(scale=0.5), dict(scale=0.5, n_levels=4, log2_hashmap_size=14), dict(scale=16.0)])
def test_forward_vs_oracle(cfg, oracle):
    scale = cfg["scale"]
    model = build_model(**cfg)
    n = 3000 + 7  # not a multiple of the 32-sample warp tile
    x, d = sample_points(n, scale, 1)
    with torch.no_grad():
        sig, rgb = model(torch.as_tensor(x).cuda(), torch.as_tensor(d).cuda())
        sig_d = model.density(torch.as_tensor(x).cuda())
    assert torch.equal(sig, sig_d), "density-only kernel path must agree with the full path"
    enc, rgbp = oracle_inputs(model)
    meta_o, _ = oracle.grid_meta(model.xyz_encoder.n_levels, model.xyz_encoder.log2_hashmap_size, 16,
                                 float(np.float32(model.per_level_scale)))
    # product level table == oracle level table
    for l in range(model.xyz_encoder.n_levels):
        assert model.xyz_encoder.meta.res[l] == meta_o.res[l] and model.xyz_encoder.meta.offset[l] == meta_o.offset[l]
        assert model.xyz_encoder.meta.scale[l] == meta_o.scale[l]
    xyz_min = torch.full((1, 3), -scale)
    xyz_max = torch.full((1, 3), scale)
    sig_o, rgb_o, h_o = oracle.torch_ngp_forward(meta_o, enc, rgbp, xyz_min, xyz_max, torch.as_tensor(x), torch.as_tensor(d))
    sig_c, rgb_c, h_c = oracle.ngp_forward_c(meta_o, enc.numpy(), rgbp.numpy(), xyz_min.numpy(), xyz_max.numpy(), x, d)
    # the two oracle restatements agree with each other (C fmaf chain vs torch matmul: fp16-ulp level)
    assert np.abs(h_c - h_o.numpy()).max() < 0.02 * max(1.0, np.abs(h_c).max())
    h_scale = max(1.0, float(h_o.abs().max()))
    err_sig = (torch.log(sig.cpu()) - torch.log(sig_o)).abs().max().item()
    assert err_sig < 0.01 * h_scale, "log-sigma error %g" % err_sig
    err_rgb = (rgb.cpu() - rgb_o).abs().max().item()
    assert err_rgb < 4e-3, "rgb error %g" % err_rgb
    # typical (median) error is far below the fp16 rounding-flip outliers
    assert (rgb.cpu() - rgb_o).abs().median().item() < 5e-4


def test_backward_vs_oracle_autograd(oracle):
    scale = 0.5
    model = build_model(scale=scale, n_levels=16, log2_hashmap_size=19)
    n = 1500 + 3
    x, d = sample_points(n, scale, 2)
    rng = np.random.RandomState(3)
    dsig = (rng.normal(size=n) * 1e-3).astype(np.float32)
    drgb = (rng.normal(size=(n, 3)) * 1e-2).astype(np.float32)

    sig, rgb = model(torch.as_tensor(x).cuda(), torch.as_tensor(d).cuda())
    loss = (sig * torch.as_tensor(dsig).cuda()).sum() + (rgb * torch.as_tensor(drgb).cuda()).sum()
    loss.backward()
    g_enc = model.xyz_encoder.params.grad.cpu()
    g_rgb = model.rgb_net.params.grad.cpu()

    enc, rgbp = oracle_inputs(model)
    enc = enc.clone().requires_grad_(True)
    rgbp = rgbp.clone().requires_grad_(True)
    meta_o, _ = oracle.grid_meta(16, 19, 16, float(np.float32(model.per_level_scale)))
    sig_o, rgb_o, _ = oracle.torch_ngp_forward(meta_o, enc, rgbp, torch.full((1, 3), -scale), torch.full((1, 3), scale),
                                               torch.as_tensor(x), torch.as_tensor(d))
    loss_o = (sig_o * torch.as_tensor(dsig)).sum() + (rgb_o * torch.as_tensor(drgb)).sum()
    loss_o.backward()

    def check(a, b, what, tol):
        scale_ = b.abs().max().item()
        err = (a - b).abs().max().item()
        assert scale_ > 0, what + ": oracle gradient is identically zero"
        assert err <= tol * scale_, "%s: max err %g vs gradient scale %g" % (what, err, scale_)

    check(g_rgb[:2048], rgbp.grad[:2048], "dW1_rgb", 0.03)
    check(g_rgb[2048:6144], rgbp.grad[2048:6144], "dW2_rgb", 0.03)
    check(g_rgb[6144:6144 + 3 * 64], rgbp.grad[6144:6144 + 3 * 64], "dW3_rgb", 0.03)
    check(g_enc[:2048], enc.grad[:2048], "dW1_density", 0.03)
    check(g_enc[2048:3072], enc.grad[2048:3072], "dW2_density", 0.03)
    check(g_enc[3072:], enc.grad[3072:], "d_table", 0.03)
    # the scatter touches exactly the entries the oracle touches
    nz_m = (g_enc[3072:] != 0)
    nz_o = (enc.grad[3072:] != 0)
    assert (nz_m & ~nz_o).sum().item() == 0
    # linearity (size-independent property): doubling the upstream gradient doubles the parameter gradient
    model.zero_grad()
    sig, rgb = model(torch.as_tensor(x).cuda(), torch.as_tensor(d).cuda())
    (2 * ((sig * torch.as_tensor(dsig).cuda()).sum() + (rgb * torch.as_tensor(drgb).cuda()).sum())).backward()
    g2 = model.rgb_net.params.grad.cpu()
    assert torch.allclose(g2, 2 * g_rgb, rtol=2e-3, atol=1e-3 * g_rgb.abs().max().item())


def test_backward_recompute_matches_saved_features():
    """ngp_net_backward with feat_save == NULL (re-gather) must equal the saved-feature path."""
    from ngp_pl_b200 import _lib
    from ngp_pl_b200.models import networks as N
    model = build_model(scale=0.5)
    n = 1000
    x, d = sample_points(n, 0.5, 5)
    x, d = torch.as_tensor(x).cuda(), torch.as_tensor(d).cuda()
    net, keep = N._net_struct(model)
    smp = N._samples_struct(x, d)
    st = torch.cuda.current_stream().cuda_stream
    sig = torch.empty(n, device="cuda")
    rgb = torch.empty(n, 3, device="cuda")
    feat = torch.empty(N.feat_save_bytes(n), device="cuda", dtype=torch.uint8)
    L = _lib.lib()
    _lib.check(L.ngp_net_forward(C.byref(net), C.byref(smp), 1, sig.data_ptr(), rgb.data_ptr(), None, feat.data_ptr(), st), "fwd")
    dsig = torch.randn(n, device="cuda") * 1e-3
    drgb = torch.randn(n, 3, device="cuda") * 1e-2
    outs = []
    ws = torch.empty(L.ngp_net_backward_workspace(n), device="cuda", dtype=torch.uint8)
    for fs in (feat.data_ptr(), None):
        ge = torch.zeros_like(model.xyz_encoder.params)
        gr = torch.zeros_like(model.rgb_net.params)
        _lib.check(L.ngp_net_backward(C.byref(net), C.byref(smp), dsig.data_ptr(), drgb.data_ptr(), fs, None,
                                      ge.data_ptr(), gr.data_ptr(), ws.data_ptr(), ws.numel(), st), "bwd")
        outs.append((ge, gr))
    torch.cuda.synchronize()
    for a, b in zip(outs[0], outs[1]):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-6 + 1e-3 * a.abs().max().item())


def test_loss_scale_invariance():
    """tiny upstream gradients (fp16 underflow territory) survive thanks to the dynamic loss scale"""
    model = build_model(scale=0.5)
    n = 2048
    x, d = sample_points(n, 0.5, 6)
    x, d = torch.as_tensor(x).cuda(), torch.as_tensor(d).cuda()
    w = torch.randn(n, 3, device="cuda")
    grads = []
    for s in (1.0, 1e-7):
        model.zero_grad()
        sig, rgb = model(x, d)
        ((rgb * w).sum() * s).backward()
        grads.append(model.rgb_net.params.grad.clone() / s)
    assert torch.allclose(grads[0], grads[1], rtol=2e-2, atol=2e-3 * grads[0].abs().max().item())
