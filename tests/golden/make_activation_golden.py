"""Golden vectors for the fused activations from the reference's own GaussianModel (importable in the CPU
container with `plyfile` stubbed; /root/reference is absent on the GPU box, so the vectors are committed):
  activations_ref.npz   get_opacity_with_3D_filter / get_scaling_with_3D_filter / get_rotation
                        (scene/gaussian_model.py:207-217,237-249) with the `.float()` casts of
                        gaussian_renderer/__init__.py:137-138, and the torch-autograd gradients of
                        sum(cot * output) w.r.t. _opacity, _scaling, _rotation.
Run:  python tests/golden/make_activation_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SFGS_REFERENCE", "/root/reference")


def load_gaussian_model():
    # scene/__init__.py pulls in dataset readers (plyfile, PIL ...); load the one module we need directly.
    sys.path.insert(0, os.path.join(ROOT, "skyfall-gs_b200"))   # simple_knn._C of this repository satisfies the import
    sys.path.insert(0, REF)
    stub = types.ModuleType("plyfile"); stub.PlyData = object; stub.PlyElement = object
    sys.modules.setdefault("plyfile", stub)
    pkg = types.ModuleType("scene"); pkg.__path__ = [os.path.join(REF, "scene")]
    sys.modules.setdefault("scene", pkg)
    spec = importlib.util.spec_from_file_location("scene.gaussian_model", os.path.join(REF, "scene", "gaussian_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.GaussianModel


def main():
    GaussianModel = load_gaussian_model()
    rng = np.random.default_rng(77)
    P = 768
    g = GaussianModel(3, False, 4, 32)
    o = rng.normal(0, 2, size=(P, 1)).astype(np.float32)
    s = rng.normal(-1.0, 1.2, size=(P, 3)).astype(np.float32)
    q = rng.normal(0, 1, size=(P, 4)).astype(np.float32)
    q[:4] *= 1e-3                                            # tiny but non-degenerate quaternions
    f = np.abs(rng.normal(0.2, 0.3, size=(P, 1))).astype(np.float64)
    f[:8] = 0.0                                              # no filter: coef == 1, scales == exp(s)
    g._opacity = torch.tensor(o, requires_grad=True)
    g._scaling = torch.tensor(s, requires_grad=True)
    g._rotation = torch.tensor(q, requires_grad=True)
    g.filter_3D = torch.tensor(f)
    opacity = g.get_opacity_with_3D_filter.float()
    scales = g.get_scaling_with_3D_filter.float()
    rot = g.get_rotation
    co = rng.normal(size=(P, 1)).astype(np.float32)
    cs = rng.normal(size=(P, 3)).astype(np.float32)
    cq = rng.normal(size=(P, 4)).astype(np.float32)
    ((opacity * torch.tensor(co)).sum() + (scales * torch.tensor(cs)).sum() + (rot * torch.tensor(cq)).sum()).backward()
    np.savez_compressed(os.path.join(HERE, "activations_ref.npz"), opacity_raw=o, scaling_raw=s, rotation_raw=q,
                        filter_3D=f, opacity=opacity.detach().numpy(), scales=scales.detach().numpy(),
                        rotations=rot.detach().numpy(), cot_opacity=co, cot_scales=cs, cot_rotations=cq,
                        g_opacity_raw=g._opacity.grad.numpy(), g_scaling_raw=g._scaling.grad.numpy(),
                        g_rotation_raw=g._rotation.grad.numpy())
    print("wrote activations_ref.npz", opacity.dtype, scales.dtype, rot.dtype)


if __name__ == "__main__":
    main()
