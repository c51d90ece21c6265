"""Load the UNMODIFIED reference (staged by oracle/build_ref.py under oracle/_ref/) -- TEST / BASELINE
INFRASTRUCTURE ONLY.

    ref = load_reference()            # ref.vren, ref.NGP, ref.render, ref.custom_functions, ref.losses
runs the reference's own models/{custom_functions,networks,rendering}.py with
    vren        = the reference's CUDA extension compiled from /root/reference/models/csrc
    tinycudann  = oracle/tcnn_standin.py (tinycudann itself is unavailable, see that file)
    torch_scatter.segment_csr = a torch restatement (only used when rays are optimised)
Needs a GPU to *run* (the reference has no CPU path).
"""
import glob
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


class Reference:
    pass


def available():
    return bool(glob.glob(os.path.join(REF, "vren*.so"))) and os.path.isdir(os.path.join(REF, "ngp_pl", "models"))


_cached = None


def load_reference():
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.py where /root/reference exists)")
    import torch  # noqa: F401  (must be imported before the extension)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    vren = importlib.import_module("vren")
    from . import tcnn_standin
    sys.modules["tinycudann"] = tcnn_standin
    ts = types.ModuleType("torch_scatter")

    def segment_csr(src, indptr):
        import torch
        out = torch.zeros((indptr.numel() - 1,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
        counts = (indptr[1:] - indptr[:-1])
        seg = torch.repeat_interleave(torch.arange(counts.numel(), device=src.device), counts)
        return out.index_add_(0, seg, src)

    ts.segment_csr = segment_csr
    sys.modules["torch_scatter"] = ts
    pkg = os.path.join(REF, "ngp_pl")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    # the reference's packages are top-level `models`, `losses`, `metrics`
    for name in ("models", "models.custom_functions", "models.rendering", "models.networks", "losses", "metrics"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(pkg):
            del sys.modules[name]
    r = Reference()
    r.vren = vren
    r.custom_functions = importlib.import_module("models.custom_functions")
    r.rendering = importlib.import_module("models.rendering")
    r.networks = importlib.import_module("models.networks")
    r.losses = importlib.import_module("losses")
    r.NGP = r.networks.NGP
    r.render = r.rendering.render
    _cached = r
    return r
