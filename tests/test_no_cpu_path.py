"""The product path has no CPU route: every operator of the drop-in module refuses CPU tensors with the reference's error
type (RuntimeError, like TORCH_CHECK(is_cuda), reference models/csrc/include/utils.h:4-6), the tinycudann-shaped modules and
the fused trainer refuse to run off the GPU. Runs without a GPU."""
import pytest
import torch

import cases  # noqa: F401  (repo root on sys.path)


def test_every_vren_operator_rejects_cpu_tensors():
    from ngp_pl_b200 import vren
    f = lambda *s: torch.zeros(*s)
    i64 = lambda *s: torch.zeros(*s, dtype=torch.int64)
    i32 = lambda *s: torch.zeros(*s, dtype=torch.int32)
    u8 = lambda *s: torch.zeros(*s, dtype=torch.uint8)
    rays_a = i64(2, 3)
    calls = {
        "ray_aabb_intersect": (f(2, 3), f(2, 3), f(1, 3), f(1, 3), 1),
        "ray_sphere_intersect": (f(2, 3), f(2, 3), f(1, 3), f(1), 1),
        "packbits": (f(1, 4096), 0.5, u8(512)),
        "morton3D": (i32(4, 3),),
        "morton3D_invert": (i32(4),),
        "raymarching_train": (f(2, 3), f(2, 3), f(2, 2), u8(128 ** 3 // 8), 1, 0.5, 0.0, f(2), 128, 1024),
        "raymarching_test": (f(2, 3), f(2, 3), f(2, 2), i64(2), u8(128 ** 3 // 8), 1, 0.5, 0.0, 128, 1024, 4),
        "composite_train_fw": (f(5), f(5, 3), f(5), f(5), rays_a, 1e-4),
        "composite_train_bw": (f(2), f(2), f(2, 3), f(5), f(5), f(5, 3), f(5), f(5), f(5), rays_a, f(2), f(2), f(2, 3), 1e-4),
        "composite_test_fw": (f(2, 4), f(2, 4, 3), f(2, 4), f(2, 4), f(2, 2), i64(2), 1e-4, i32(2), f(2), f(2), f(2, 3)),
        "distortion_loss_fw": (f(5), f(5), f(5), rays_a),
        "distortion_loss_bw": (f(2), f(5), f(5), f(5), f(5), f(5), rays_a),
    }
    exported = [n for n in dir(vren) if not n.startswith("_") and callable(getattr(vren, n)) and getattr(vren, n).__module__ == vren.__name__]
    assert sorted(exported) == sorted(calls), "the 12 names of reference models/csrc/binding.cpp:234-250"
    for name, args in calls.items():
        with pytest.raises(RuntimeError):
            getattr(vren, name)(*args)


def test_modules_and_trainer_refuse_to_run_on_the_cpu():
    from ngp_pl_b200.models.networks import NGP
    from ngp_pl_b200.trainer import Trainer
    model = NGP(0.5)  # construction (parameter init, level table) is host work
    with pytest.raises(RuntimeError):
        model(torch.zeros(4, 3), torch.zeros(4, 3))
    with pytest.raises(RuntimeError):
        model.density(torch.zeros(4, 3))
    with pytest.raises(RuntimeError):
        Trainer(model, n_rays=64)


def test_tcnn_shaped_modules_refuse_cpu_tensors():
    from ngp_pl_b200 import tcnn
    enc = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4})
    with pytest.raises(RuntimeError):
        enc(torch.zeros(4, 3))
    net = tcnn.Network(32, 3, {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "Sigmoid", "n_neurons": 64,
                               "n_hidden_layers": 2})
    with pytest.raises(RuntimeError):
        net(torch.zeros(4, 32, dtype=torch.float16))
