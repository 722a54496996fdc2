"""Small fwd+bwd + sibling-op run for compute-sanitizer (memcheck / racecheck / synccheck). Not a pytest file.

    compute-sanitizer --tool racecheck python tests/gpu_sanitize_case.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "skyfall-gs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import helpers as Hh
from sfgs import synthetic as S

dev = torch.device("cuda:0")
bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
# (a) ordinary small scene, odd image size; (b) everything piled on a few tiles -> long lists (heavy sort + radix paths)
cases = [(S.blob_scene(4000, seed=3), S.simple_camera(200, 120)),
         (S.blob_scene(6000, seed=4, spread=0.15), S.simple_camera(96, 64, distance=3.0))]
for scene, cam in cases:
    d = Hh.to_torch(scene, cam, dev)
    cot = [torch.from_numpy(c).to(dev) for c in S.cotangents(cam.width, cam.height, seed=2)]
    f = Hh.run_ours_forward(d, cam, 3, bg)
    b = Hh.run_ours_backward(d, cam, 3, bg, f, cot)
    torch.cuda.synchronize()
    rg = Hh.our_internals(f, scene.P, cam.height, cam.width)["ranges"].cpu().numpy().astype(np.int64)
    print("case", scene.P, cam.width, cam.height, "R", int(f["num_rendered"]), "max tile list", int((rg[:, 1] - rg[:, 0]).max()))
# fused normalize path + siblings + activations
import diff_gauss
from diff_gauss import GaussianRasterizationSettings, GaussianRasterizer
from fused_ssim import fused_ssim
from simple_knn._C import distCUDA2
from sfgs.activations import fused_activations
scene, cam = cases[0]
d = Hh.to_torch(scene, cam, dev)
leaves = {k: d[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "shs", "scales", "rotations")}
m2d = torch.zeros((scene.P, 3), device=dev, requires_grad=True)
rs = GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, 0.1, torch.zeros(1, device=dev), bg, 1.0,
                                   d["viewmatrix"], d["projmatrix"], 3, d["campos"], False, False)
color, depth, norm, alpha, radii, _ = GaussianRasterizer(rs)(leaves["means3D"], m2d, leaves["opacities"], shs=leaves["shs"],
                                                             scales=leaves["scales"], rotations=leaves["rotations"])
gt = torch.rand_like(color)
loss = (color - gt).abs().mean() + 0.2 * (1 - fused_ssim(color[None], gt[None])) + 0.01 * norm.mean() + 0.01 * depth.mean()
loss.backward()
print("dist2", float(distCUDA2(d["means3D"]).mean()))
o, s, q = (torch.randn(1000, k, device=dev, requires_grad=True) for k in (1, 3, 4))
op, sc, rt = fused_activations(o, s, q, torch.rand(1000, 1, device=dev, dtype=torch.float64))
(op.sum() + sc.sum() + rt.sum()).backward()
torch.cuda.synchronize()
print("sanitize case done")
