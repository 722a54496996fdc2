"""Fused `GaussianModel.compute_3D_filter` (scene/gaussian_model.py:254-308).

    pc.filter_3D = compute_3D_filter(pc.get_xyz, cameras)          # instead of pc.compute_3D_filter(cameras)

`cameras` are the reference's Camera objects (attributes R, T, focal_x, focal_y, cx, cy, image_width, image_height).
Returns the [P,1] float64 tensor the reference stores.  One kernel pass over the Gaussians instead of ~15 float64
torch kernels per camera.
"""
from __future__ import annotations

import numpy as np
import torch

from . import native as N


def camera_table(cameras) -> np.ndarray:
    rows = []
    for cam in cameras:
        R = np.asarray(cam.R, dtype=np.float64).reshape(3, 3)
        T = np.asarray(cam.T, dtype=np.float64).reshape(3)
        w, h = float(cam.image_width), float(cam.image_height)
        cx_ori = float(cam.cx) / 2 * w + w / 2            # scene/gaussian_model.py:284-285
        cy_ori = float(cam.cy) / 2 * h + h / 2
        rows.append(np.concatenate([R.ravel(), T, [float(cam.focal_x), float(cam.focal_y), cx_ori, cy_ori, w, h]]))
    return np.stack(rows) if rows else np.zeros((0, 18))


@torch.no_grad()
def compute_3D_filter(xyz: torch.Tensor, cameras) -> torch.Tensor:
    L = N.lib()
    dev = xyz.device
    P = int(xyz.shape[0])
    table = camera_table(cameras)
    focal_max = 0.0
    for cam in cameras:                                   # "if focal_length < camera.focal_x" (:300-301)
        focal_max = max(focal_max, float(cam.focal_x))
    out = torch.empty((P,), dtype=torch.float64, device=dev)
    if P == 0:
        return out[..., None]
    cams = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
    scratch = torch.empty(1, dtype=torch.int64, device=dev)
    x = xyz.detach().contiguous().float()
    with torch.cuda.device(dev):
        N.check(L.sfgs_compute_3d_filter(P, x.data_ptr(), int(table.shape[0]), cams.data_ptr() if table.shape[0] else None,
                                         float(focal_max), out.data_ptr(), scratch.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "sfgs_compute_3d_filter")
    return out[..., None]
