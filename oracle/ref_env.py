"""Load the UNMODIFIED reference (staged by oracle/build_ref.py under oracle/_ref/) -- TEST / BASELINE
INFRASTRUCTURE ONLY.

    ref = load_reference()            # ref.vren, ref.NGP, ref.render, ref.custom_functions, ref.losses
runs the reference's own models/{custom_functions,networks,rendering}.py with
    vren        = the reference's CUDA extension compiled from /root/reference/models/csrc
    tinycudann  = oracle/tcnn_standin.py (checker) or oracle/tcnn_fast.py (performance-grade); tinycudann itself is unavailable
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


def python_available():
    return os.path.isdir(os.path.join(REF, "ngp_pl", "models"))


_cached = {}


def load_reference(drop_in=False, tcnn="standin"):
    """drop_in=False: the reference's Python on the REFERENCE's compiled vren + a tinycudann stand-in:
                   tcnn="standin" the checker-grade restatement (oracle/tcnn_standin.py),
                   tcnn="fast"    the performance-grade one (oracle/tcnn_fast.py; what `bench.py --impl reference` times).
    drop_in=True : the reference's Python on ngp_pl_b200.vren + ngp_pl_b200.tcnn (the drop-in claim under test)."""
    key = (drop_in, tcnn)
    if key in _cached:
        return _cached[key]
    import torch  # noqa: F401  (must be imported before the extension)
    if drop_in:
        if not python_available():
            raise RuntimeError("reference python is not staged under oracle/_ref/ngp_pl")
        import ngp_pl_b200.tcnn as tcnn_impl
        import ngp_pl_b200.vren as vren
        saved = {k: sys.modules.get(k) for k in ("vren", "tinycudann")}
        sys.modules["vren"] = vren
        sys.modules["tinycudann"] = tcnn_impl
    else:
        if not available():
            raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.py where /root/reference exists)")
        if REF not in sys.path:
            sys.path.insert(0, REF)
        saved = {k: sys.modules.get(k) for k in ("vren", "tinycudann")}
        for k in ("vren",):
            if k in sys.modules and getattr(sys.modules[k], "__name__", "") != "vren":
                del sys.modules[k]
        if "vren" in sys.modules and getattr(sys.modules["vren"], "__file__", "") and "ngp_pl_b200" in sys.modules["vren"].__file__:
            del sys.modules["vren"]
        vren = importlib.import_module("vren")
        if tcnn == "fast":
            from . import tcnn_fast as tcnn_mod
        else:
            from . import tcnn_standin as tcnn_mod
        sys.modules["tinycudann"] = tcnn_mod
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
    # the reference's packages are top-level `models`, `losses`, `metrics`; always import a FRESH copy so that the
    # two bindings (reference kernels / our kernels) get separate module objects
    for name in ("models", "models.custom_functions", "models.rendering", "models.networks", "losses", "metrics"):
        sys.modules.pop(name, None)
    r = Reference()
    r.vren = vren
    r.custom_functions = importlib.import_module("models.custom_functions")
    r.rendering = importlib.import_module("models.rendering")
    r.networks = importlib.import_module("models.networks")
    r.losses = importlib.import_module("losses")
    r.NGP = r.networks.NGP
    r.render = r.rendering.render
    # leave no reference-named modules behind (the loaded copies keep their own globals)
    for name in ("models", "models.custom_functions", "models.rendering", "models.networks", "losses", "metrics"):
        sys.modules.pop(name, None)
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    _cached[key] = r
    return r
