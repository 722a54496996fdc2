"""numpy restatement of GaussianModel.compute_3D_filter (scene/gaussian_model.py:254-308).  TEST INFRASTRUCTURE: only
tests/ import it.  Pinned by tests/golden/filter3d_ref.npz, which the reference's own method produced
(tests/golden/make_filter3d_golden.py)."""
import numpy as np


def compute_3D_filter(xyz, cameras):
    xyz = np.asarray(xyz, dtype=np.float64)                               # :258
    distance = np.ones(xyz.shape[0]) * 1e8                                # :260
    valid_points = np.zeros(xyz.shape[0], dtype=bool)                     # :261
    focal_length = 0.0
    for cam in cameras:
        R, T = np.asarray(cam.R, np.float64), np.asarray(cam.T, np.float64)
        xyz_cam = xyz @ R + T[None, :]                                    # :271
        valid_depth = xyz_cam[:, 2] > 0.2                                 # :276
        x, y, z = xyz_cam[:, 0], xyz_cam[:, 1], np.maximum(xyz_cam[:, 2], 0.001)   # :279-280
        cx_ori = cam.cx / 2 * cam.image_width + cam.image_width / 2       # :283-284
        cy_ori = cam.cy / 2 * cam.image_height + cam.image_height / 2
        x = x / z * cam.focal_x + cx_ori
        y = y / z * cam.focal_y + cy_ori
        in_screen = (x >= -0.15 * cam.image_width) & (x <= cam.image_width * 1.15) & \
                    (y >= -0.15 * cam.image_height) & (y <= 1.15 * cam.image_height)   # :290
        valid = valid_depth & in_screen
        distance[valid] = np.minimum(distance[valid], z[valid])           # :296
        valid_points |= valid
        focal_length = max(focal_length, cam.focal_x)                     # :298-299
    distance[~valid_points] = distance[valid_points].max()                # :301
    return (distance / focal_length * (0.2 ** 0.5))[:, None]              # :305-308
