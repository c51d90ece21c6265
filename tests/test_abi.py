"""CPU-side checks of the drop-in boundary: libngp_b200.so loads, exports every symbol that
include/ngp_b200.h declares (no compute calls without a GPU), the ctypes table agrees with the header,
and the Python shim exposes the reference's twelve `vren` names."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ngp_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ngp_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from ngp_pl_b200 import build, _lib
    build.build()
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libngp_b200.so does not export " + s
    assert sorted(_lib.SIGNATURES.keys()) == syms, (
        "ctypes table and header disagree: %s" % (set(_lib.SIGNATURES.keys()) ^ set(syms)))
    assert lib.ngp_abi_version() == _lib.ABI_VERSION


def test_header_is_plain_c():
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "ngp_b200.h"\nint main(void){return sizeof(NgpGridMeta)>0?0:1;}\n')
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.dirname(HEADER), "-c", c, "-o",
                            os.path.join(d, "t.o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        assert r.returncode == 0, r.stdout.decode()


def test_struct_layouts_match_header():
    """sizeof of the by-value structs as the C compiler sees them == ctypes mirrors"""
    import ctypes, subprocess, tempfile
    from ngp_pl_b200 import _lib
    names = ["NgpGridMeta", "NgpNet", "NgpSamples", "NgpTrainCfg"]
    present = [n for n in names if re.search(r"\}\s*%s\s*;" % n, open(HEADER).read())]
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        body = "".join('printf("%%zu\\n", sizeof(%s));' % n for n in present)
        open(c, "w").write('#include <stdio.h>\n#include "ngp_b200.h"\nint main(void){%s return 0;}\n' % body)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).decode().split()]
    for n, s in zip(present, sizes):
        assert ctypes.sizeof(getattr(_lib, n)) == s, "%s: C says %d, ctypes says %d" % (n, s, ctypes.sizeof(getattr(_lib, n)))


def test_vren_surface():
    from ngp_pl_b200 import vren
    for name in ["ray_aabb_intersect", "ray_sphere_intersect", "packbits", "morton3D", "morton3D_invert",
                 "raymarching_train", "raymarching_test", "composite_train_fw", "composite_train_bw",
                 "composite_test_fw", "distortion_loss_fw", "distortion_loss_bw"]:
        assert callable(getattr(vren, name)), name


def test_grid_meta_tables_match_oracle(oracle):
    """level tables (host code, no GPU): product C-ABI vs the oracle restatement, BASELINE configs 1, 2, 5"""
    import numpy as np
    from ngp_pl_b200 import _lib
    for L, log2_T, scale in [(4, 14, 0.5), (16, 19, 0.5), (16, 19, 16.0), (16, 19, 2.0)]:
        b = float(np.exp(np.log(2048 * scale / 16) / (L - 1)))
        m, total = _lib.grid_meta(L, log2_T, 16, b)
        mo, total_o = oracle.grid_meta(L, log2_T, 16, float(np.float32(b)))
        assert total == total_o
        for l in range(L):
            assert m.res[l] == mo.res[l] and m.offset[l] == mo.offset[l] and m.scale[l] == mo.scale[l]
        assert m.hashed_mask == mo.hashed_mask
    m, total = _lib.grid_meta(16, 19, 16, float(np.exp(np.log(2048 * 0.5 / 16) / 15)))
    assert total == 5722520 and 2 * total + 3072 == 11448112
    assert list(m.res)[:6] == [16, 22, 28, 37, 49, 65] and m.res[15] == 1025


def test_ops_fail_loudly_without_gpu_or_library(monkeypatch):
    import torch
    from ngp_pl_b200 import vren, _lib
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            vren.morton3D(torch.zeros(4, 3, dtype=torch.int32))  # CPU tensor: no CPU fallback
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libngp_b200.so")
    with pytest.raises(RuntimeError):
        _lib.lib()


def test_graft_entry_build_runs_on_cpu():
    """the driver's "does it build" check: __graft_entry__.build() compiles (cached) and imports everything without a GPU"""
    import importlib
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    entry = importlib.import_module("__graft_entry__")
    entry.build()
