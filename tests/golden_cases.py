"""Seeded definitions of the golden cases shared by make_golden.py (reference CUDA, on the GPU box) and the tests."""
import numpy as np

from sfgs import synthetic as S

CASES = ["blob_sh3", "blob_sh1_mod", "blob_precomp", "city_small"]


def build_case(name):
    """-> (scene, camera, bg[3] float32, kwargs)."""
    if name == "blob_sh3":
        scene, cam = S.blob_scene(1500, seed=3, sh_degree=3), S.simple_camera(200, 136)   # 136 is not a multiple of 16
        kw = dict(sh_degree=3, kernel_size=0.1, scale_modifier=1.0)
        bg = np.array([0.1, 0.2, 0.3], np.float32)
    elif name == "blob_sh1_mod":
        scene, cam = S.blob_scene(1200, seed=7, sh_degree=3), S.simple_camera(160, 160, fov_deg=50.0, distance=7.0)
        kw = dict(sh_degree=1, kernel_size=0.3, scale_modifier=1.4)   # active degree below the stored one
        bg = np.array([1.0, 1.0, 1.0], np.float32)
    elif name == "blob_precomp":
        scene, cam = S.blob_scene(1000, seed=9, sh_degree=0), S.simple_camera(128, 96)
        rng = np.random.default_rng(99)
        kw = dict(sh_degree=0, kernel_size=0.1, scale_modifier=1.0, colors_precomp=True,
                  colors=rng.uniform(0, 1, size=(scene.P, 3)).astype(np.float32))
        bg = np.array([0.0, 0.0, 0.0], np.float32)
    elif name == "city_small":
        scene, cam = S.city_scene(6_000, seed=5, sh_degree=3, extent=36.0), S.jax004_camera(320, 180)
        scene.scales *= 2.0
        kw = dict(sh_degree=3, kernel_size=0.1, scale_modifier=1.0)
        bg = np.array([0.0, 0.0, 0.0], np.float32)
    else:
        raise KeyError(name)
    kw["cot"] = S.cotangents(cam.width, cam.height, seed=17)
    return scene, cam, bg, kw
