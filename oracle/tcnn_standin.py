"""Stand-in for `tinycudann` -- TEST / BASELINE INFRASTRUCTURE ONLY (never imported by ngp_pl_b200/).

tinycudann is not vendored by the reference (README.md:39), not pinned, and not installable here, so
the reference arm (`bench.py --impl reference`) and the end-to-end parity tests run the reference's
UNMODIFIED models/networks.py against this module: the same three classes
(NetworkWithInputEncoding / Encoding / Network, reference networks.py:36-77) implemented with plain
PyTorch ops on the GPU (index gathers + trilinear weights + F.linear), fp16 outputs, one flat fp32
`params` Parameter each, tinycudann's parameter layout. Any number reported with it must be labelled
"reference vren + tinycudann STAND-IN".
"""
import math

import torch
from torch import nn

from . import oracle as _o


class _Meta:
    pass


def _rt(x):
    # fp16 rounding of the value, fp32 straight-through gradient
    return x + (x.half().float() - x).detach()


class NetworkWithInputEncoding(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        e = encoding_config
        self.meta, self.n_entries = _o.grid_meta(int(e["n_levels"]), int(e["log2_hashmap_size"]), int(e["base_resolution"]),
                                                 float(torch.tensor(float(e["per_level_scale"]), dtype=torch.float32)))
        g = torch.Generator().manual_seed(seed)
        p = torch.empty(3072 + 2 * self.n_entries)
        p[:2048].uniform_(-math.sqrt(6 / 96), math.sqrt(6 / 96), generator=g)
        p[2048:3072].uniform_(-math.sqrt(6 / 80), math.sqrt(6 / 80), generator=g)
        p[3072:].uniform_(-1e-4, 1e-4, generator=g)
        self.params = nn.Parameter(p)

    def forward(self, x01):
        with torch.autocast("cuda", enabled=False):
            p = _rt(self.params)
            feat = _o.torch_grid_encode(self.meta, p[3072:].view(-1, 2), x01.float())
            hid = _rt(torch.relu(feat @ p[:2048].view(64, 32).t()))
            return (hid @ p[2048:3072].view(16, 64).t()).half()


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config):
        super().__init__()
        self.params = nn.Parameter(torch.zeros(0))

    def forward(self, u):
        with torch.autocast("cuda", enabled=False):
            return _o.torch_sh4(u.float() * 2 - 1).half()


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1338):
        super().__init__()
        self.n_in, self.n_out = n_input_dims, n_output_dims
        self.sigmoid = network_config.get("output_activation", "None") == "Sigmoid"
        g = torch.Generator().manual_seed(seed)
        p = torch.empty(7168)
        p[:2048].uniform_(-math.sqrt(6 / 96), math.sqrt(6 / 96), generator=g)
        p[2048:6144].uniform_(-math.sqrt(6 / 128), math.sqrt(6 / 128), generator=g)
        p[6144:].uniform_(-math.sqrt(6 / 80), math.sqrt(6 / 80), generator=g)
        self.params = nn.Parameter(p)

    def forward(self, x):
        with torch.autocast("cuda", enabled=False):
            p = _rt(self.params)
            x = _rt(x.float())
            r1 = _rt(torch.relu(x @ p[:2048].view(64, 32).t()))
            r2 = _rt(torch.relu(r1 @ p[2048:6144].view(64, 64).t()))
            out = (r2 @ p[6144:].view(16, 64).t())[:, :self.n_out]
            if self.sigmoid:
                out = torch.sigmoid(out)
            return out.half()
