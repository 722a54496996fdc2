"""Golden vectors for the fused 3D-filter kernel, produced by the reference's OWN GaussianModel.compute_3D_filter
(scene/gaussian_model.py:254-308) on CPU.  Run where /root/reference exists."""
import math
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SFGS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", "..", "skyfall-gs_b200"))
for name in ("plyfile", "OpenEXR"):
    m = types.ModuleType(name); m.__getattr__ = lambda k: (lambda *a, **k2: None); sys.modules.setdefault(name, m)
from scene.gaussian_model import GaussianModel  # noqa: E402

rng = np.random.default_rng(11)
P, C = 20000, 37
xyz = np.concatenate([rng.uniform(-60, 60, (P, 2)), np.abs(rng.normal(0, 8, (P, 1)))], 1).astype(np.float32)
cams = []
for c in range(C):
    ang, el, rad = rng.uniform(0, 2 * math.pi), rng.uniform(0.5, 1.4), rng.uniform(60, 160)
    eye = rad * np.array([math.cos(el) * math.cos(ang), math.cos(el) * math.sin(ang), math.sin(el)])
    fwd = -eye / np.linalg.norm(eye)
    right = np.cross(fwd, [0, 0, 1.0]); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4); c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
    w2c = np.linalg.inv(c2w)
    W, H = int(rng.choice([640, 1024, 1920])), int(rng.choice([360, 1024, 1080]))
    fx = (W / 2) / math.tan(math.radians(rng.uniform(15, 50)) / 2)
    cams.append(types.SimpleNamespace(R=w2c[:3, :3].T.copy(), T=w2c[:3, 3].copy(), focal_x=fx, focal_y=fx * rng.uniform(0.95, 1.05),
                                      cx=rng.uniform(-0.05, 0.05), cy=rng.uniform(-0.05, 0.05), image_width=W, image_height=H))
pc = GaussianModel(3, False, 4, 32)
pc._xyz = torch.from_numpy(xyz)
pc.compute_3D_filter(cams)
out = pc.filter_3D.numpy()
np.savez_compressed(os.path.join(HERE, "filter3d_ref.npz"), xyz=xyz, filter_3D=out,
                    cams=np.stack([np.concatenate([np.asarray(c.R).ravel(), c.T, [c.focal_x, c.focal_y, c.cx, c.cy, c.image_width, c.image_height]])
                                   for c in cams]))
print("wrote filter3d_ref.npz", out.shape, out.dtype, float(out.min()), float(out.max()), "never seen:", int((out == out.max()).sum()))
