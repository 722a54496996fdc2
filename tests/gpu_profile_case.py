"""Run a few fwd+bwd iterations of one configuration (for ncu launch lists / captures). Not a pytest file."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "skyfall-gs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import helpers as Hh
from sfgs import synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="ours", choices=["ours", "ref"])
ap.add_argument("--P", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--cam", default="jax", choices=["jax", "orbit"])
a = ap.parse_args()
dev = torch.device("cuda:0")
scene = S.city_scene(a.P, seed=0)
cam = S.jax004_camera() if a.cam == "jax" else S.orbit_camera()
d = Hh.to_torch(scene, cam, dev)
bg = torch.zeros(3, device=dev)
cot = [torch.from_numpy(c).to(dev) for c in S.cotangents(cam.width, cam.height)]
for it in range(a.iters):
    if a.impl == "ours":
        f = Hh.run_ours_forward(d, cam, 3, bg)
        b = Hh.run_ours_backward(d, cam, 3, bg, f, cot)
    else:
        f = Hh.run_ref_forward(d, cam, 3, bg)
        b = Hh.run_ref_backward(d, cam, 3, bg, f, cot)
    torch.cuda.synchronize()
print("done", f["num_rendered"])
