"""The performance-grade tinycudann stand-in of the reference arm (oracle/tcnn_fast.py) against the checker-grade one
(oracle/tcnn_standin.py): same parameters -> outputs equal to fp16 resolution, parameter gradients to 2 % of the
gradient's max. Runs on the CPU (small batch) and, marked gpu, on the device at the bench's batch size."""
import numpy as np
import pytest
import torch

from oracle import tcnn_fast as F
from oracle import tcnn_standin as S

B = float(np.exp(np.log(2048 * 0.5 / 16) / 15))
ENC = {"otype": "Grid", "type": "Hash", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19,
       "base_resolution": 16, "per_level_scale": B, "interpolation": "Linear"}
NET1 = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None", "n_neurons": 64, "n_hidden_layers": 1}
NET2 = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 64, "n_hidden_layers": 2}


def _compare(dev, n):
    torch.manual_seed(0)
    a, f = S.NetworkWithInputEncoding(3, 16, ENC, NET1).to(dev), F.NetworkWithInputEncoding(3, 16, ENC, NET1).to(dev)
    with torch.no_grad():
        a.params[3072:].uniform_(-0.5, 0.5)
        f.params.copy_(a.params)
    x = torch.rand(n, 3, device=dev)
    ya, yf = a(x), f(x)
    assert yf.dtype == torch.float16 and yf.shape == (n, 16)
    assert (ya.float() - yf.float()).abs().max().item() < 2e-3 * max(1.0, ya.float().abs().max().item())
    g = torch.randn(n, 16, device=dev)
    (ya.float() * g).sum().backward()
    (yf.float() * g).sum().backward()
    ga, gf = a.params.grad, f.params.grad
    for lo, hi in ((0, 2048), (2048, 3072), (3072, ga.numel())):
        assert (ga[lo:hi] - gf[lo:hi]).abs().max().item() < 2e-2 * ga[lo:hi].abs().max().item()
    ra, rf = S.Network(32, 3, NET2).to(dev), F.Network(32, 3, NET2).to(dev)
    xi = torch.randn(n, 32, device=dev).half()
    oa, of = ra(xi), rf(xi)
    assert (oa.float() - of.float()).abs().max().item() < 2e-3
    go = torch.randn(n, 3, device=dev)
    (oa.float() * go).sum().backward()
    (of.float() * go).sum().backward()
    assert (ra.params.grad - rf.params.grad).abs().max().item() < 2e-2 * ra.params.grad.abs().max().item()
    d = torch.rand(n, 3, device=dev)
    ea, ef = S.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4}), F.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4})
    assert torch.equal(ea(d), ef(d))


def test_fast_standin_matches_checker_cpu():
    _compare("cpu", 2048)


@pytest.mark.gpu
def test_fast_standin_matches_checker_gpu():
    _compare("cuda", 200_000)
