"""Seeded synthetic scenes and cameras "of the named shape" (SURVEY.md §8d).

No dataset ships with the reference, so every configuration runs on synthetic
Gaussians: a JAX_004-shaped urban height field normalised like
scene/dataset_readers.py:383-390 (radius-256 disc), flat-ish surfels, random
unit quaternions, sigmoid(N(0,2^2)) opacities, SH degree 3.  Cameras follow the
reference's conventions exactly: row-vector ("transposed") world_view and
full_proj matrices (scene/cameras.py:64-73, utils/graphics_utils.py:38-126) and
the OpenGL->COLMAP flip of render_video.py:96-107.

Everything is generated with numpy's PCG64 so CPU oracle, golden fixtures and
GPU runs see bit-identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

SH_C0 = 0.28209479177387814

# frame 0 of camera_paths/JAX/004/r488_e50_fov20.json (camera_to_world, OpenGL axes), fov 20 deg
JAX004_FRAME0_C2W = np.array([
    0.0, -0.7677176206816371, 0.6407883073956063, 313.0,
    1.0, 0.0, -0.0, 0.0,
    -0.0, 0.6407883073956063, 0.7677176206816371, 433.40000000000003,
    0.0, 0.0, 0.0, 1.0], dtype=np.float64).reshape(4, 4)
JAX004_FOV_DEG = 20.0


@dataclass
class Camera:
    width: int
    height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # [4,4] float32, row-vector convention (world_view_transform)
    projmatrix: np.ndarray   # [4,4] float32, full_proj_transform
    campos: np.ndarray       # [3] float32


def _projection(znear, zfar, fovx, fovy):
    t, r = math.tan(fovy / 2) * znear, math.tan(fovx / 2) * znear
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2.0 * znear / (2 * r)
    P[1, 1] = 2.0 * znear / (2 * t)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_from_w2c(R_c2w_cols: np.ndarray, T: np.ndarray, fovx: float, fovy: float, width: int, height: int,
                    znear: float = 0.01, zfar: float = 100.0) -> Camera:
    """R, T in the reference's CameraInfo convention: R = transpose(w2c[:3,:3]), T = w2c[:3,3]."""
    Rt = np.zeros((4, 4), dtype=np.float64)
    Rt[:3, :3] = R_c2w_cols.T
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    w2c = Rt.astype(np.float32)
    view = np.ascontiguousarray(w2c.T)                       # world_view_transform
    proj = np.ascontiguousarray(_projection(znear, zfar, fovx, fovy).T)
    full = (view @ proj).astype(np.float32)                  # full_proj_transform
    campos = np.linalg.inv(view.astype(np.float64))[3, :3].astype(np.float32)
    return Camera(width, height, math.tan(fovx * 0.5), math.tan(fovy * 0.5), view, np.ascontiguousarray(full), campos)


def camera_from_c2w_opengl(c2w: np.ndarray, fov_deg: float, width: int, height: int) -> Camera:
    c2w = np.array(c2w, dtype=np.float64).reshape(4, 4).copy()
    c2w[:3, 1:3] *= -1
    w2c = np.linalg.inv(c2w)
    R = w2c[:3, :3].T
    T = w2c[:3, 3]
    focal = (height / 2.0) / math.tan(math.radians(fov_deg) / 2.0)
    fovx = 2 * math.atan(width / (2 * focal))
    fovy = 2 * math.atan(height / (2 * focal))
    return camera_from_w2c(R, T, fovx, fovy, width, height)


def jax004_camera(width: int = 1920, height: int = 1080) -> Camera:
    return camera_from_c2w_opengl(JAX004_FRAME0_C2W, JAX004_FOV_DEG, width, height)


def orbit_camera(target=(0.0, 0.0, 0.0), elevation_deg=85.0, azimuth_deg=0.0, radius=300.0, fov_deg=60.0,
                 width: int = 1920, height: int = 1080) -> Camera:
    """Look-at camera on an orbit (z up), COLMAP axes (x right, y down, z forward)."""
    el, az = math.radians(elevation_deg), math.radians(azimuth_deg)
    tgt = np.array(target, dtype=np.float64)
    eye = tgt + radius * np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
    fwd = tgt - eye
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 0.0, 1.0])
    if abs(np.dot(fwd, up)) > 0.999:
        up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
    w2c = np.linalg.inv(c2w)
    R = w2c[:3, :3].T
    T = w2c[:3, 3]
    focal = (height / 2.0) / math.tan(math.radians(fov_deg) / 2.0)
    return camera_from_w2c(R, T, 2 * math.atan(width / (2 * focal)), 2 * math.atan(height / (2 * focal)), width, height)


def simple_camera(width: int, height: int, fov_deg: float = 60.0, distance: float = 8.0) -> Camera:
    """Small pinhole camera on the -z axis looking at the origin (config 1 / unit tests)."""
    c2w = np.eye(4)
    c2w[2, 3] = -distance
    w2c = np.linalg.inv(c2w)
    focal = (height / 2.0) / math.tan(math.radians(fov_deg) / 2.0)
    return camera_from_w2c(w2c[:3, :3].T, w2c[:3, 3], 2 * math.atan(width / (2 * focal)),
                           2 * math.atan(height / (2 * focal)), width, height)


@dataclass
class Scene:
    means3D: np.ndarray    # [P,3]
    scales: np.ndarray     # [P,3] activated
    rotations: np.ndarray  # [P,4] unit quaternions (w,x,y,z)
    opacities: np.ndarray  # [P,1] activated
    shs: np.ndarray        # [P,M,3]
    sh_degree: int

    @property
    def P(self):
        return self.means3D.shape[0]


def city_scene(P: int = 1_000_000, seed: int = 0, sh_degree: int = 3, extent: float = 256.0) -> Scene:
    """JAX_004-shaped scene: xy ~ U(-extent, extent), z ~ |N(0,15)| clipped to [0,120]."""
    rng = np.random.default_rng(seed)
    M = (sh_degree + 1) ** 2
    xy = rng.uniform(-extent, extent, size=(P, 2))
    z = np.clip(np.abs(rng.normal(0.0, 15.0, size=(P, 1))), 0.0, 120.0)
    means = np.concatenate([xy, z], axis=1).astype(np.float32)
    scales = np.exp(rng.normal(math.log(0.5), 0.6, size=(P, 3)))
    scales[:, 2] *= 0.3
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = 1.0 / (1.0 + np.exp(-rng.normal(0.0, 2.0, size=(P, 1))))
    shs = np.zeros((P, M, 3))
    shs[:, 0, :] = (rng.uniform(0.0, 1.0, size=(P, 3)) - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:, :] = rng.normal(0.0, 0.05, size=(P, M - 1, 3))
    return Scene(means, scales.astype(np.float32), q.astype(np.float32), opac.astype(np.float32),
                 shs.astype(np.float32), sh_degree)


def blob_scene(P: int = 2000, seed: int = 0, sh_degree: int = 3, spread: float = 2.5, scale: float = 0.15) -> Scene:
    """Compact cloud around the origin for small cameras (config 1 and unit tests)."""
    rng = np.random.default_rng(seed)
    M = (sh_degree + 1) ** 2
    means = rng.normal(0.0, spread, size=(P, 3)).astype(np.float32)
    scales = np.exp(rng.normal(math.log(scale), 0.5, size=(P, 3)))
    scales[:, 2] *= 0.3
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = 1.0 / (1.0 + np.exp(-rng.normal(0.0, 2.0, size=(P, 1))))
    shs = np.zeros((P, M, 3))
    shs[:, 0, :] = (rng.uniform(0.0, 1.0, size=(P, 3)) - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:, :] = rng.normal(0.0, 0.1, size=(P, M - 1, 3))
    return Scene(means, scales.astype(np.float32), q.astype(np.float32), opac.astype(np.float32),
                 shs.astype(np.float32), sh_degree)


def cotangents(width: int, height: int, seed: int = 1):
    """Fixed N(0,1) pixel cotangents for colour, depth, normal and alpha (backward inputs)."""
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(3, height, width)).astype(np.float32),
            rng.normal(size=(1, height, width)).astype(np.float32),
            rng.normal(size=(3, height, width)).astype(np.float32),
            rng.normal(size=(1, height, width)).astype(np.float32))
