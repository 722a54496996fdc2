"""The CPU oracle against the reference's own Python code (golden vectors) and against itself (no GPU needed)."""
import math
import os

import numpy as np
import pytest

from oracle import cpu_oracle as O
from oracle import knn_oracle, ssim_oracle
from sfgs import synthetic as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fwd(scene, cam, bg=(0, 0, 0), **kw):
    return O.forward(scene.means3D, scene.scales, scene.rotations, scene.opacities, scene.shs,
                     kw.pop("sh_degree", scene.sh_degree), cam.viewmatrix, cam.projmatrix, cam.campos, cam.width,
                     cam.height, cam.tanfovx, cam.tanfovy, np.array(bg, np.float32), **kw)


def test_sh_matches_reference_eval_sh():
    """SH->RGB of the oracle vs utils/sh_utils.py::eval_sh (golden vectors made by tests/golden/make_python_golden.py)."""
    g = np.load(os.path.join(GOLD, "sh_ref.npz"))
    keep = g["dirs"][:, 2] > 0.3
    dirs, coef = g["dirs"][keep], g["coef"][keep]
    P = dirs.shape[0]
    cam = S.camera_from_w2c(np.eye(3), np.zeros(3), 2 * math.atan(4.0), 2 * math.atan(4.0), 256, 256)
    assert np.allclose(cam.campos, 0)
    means = (dirs * 5.0).astype(np.float32)
    shs = np.ascontiguousarray(np.transpose(coef, (0, 2, 1)))   # [P,16,3]
    sc = S.Scene(means, np.full((P, 3), 0.05, np.float32), np.tile(np.array([1, 0, 0, 0], np.float32), (P, 1)),
                 np.full((P, 1), 0.5, np.float32), shs, 3)
    for deg in range(4):
        f = fwd(sc, cam, sh_degree=deg)
        assert (f["radii"] > 0).all()
        want = np.maximum(g[f"deg{deg}"][keep] + 0.5, 0.0)
        assert np.abs(f["rgb"] - want).max() < 2e-6
        assert np.array_equal(f["clamped"].astype(bool), (g[f"deg{deg}"][keep] + 0.5) < 0) or \
            np.abs((g[f"deg{deg}"][keep] + 0.5)[f["clamped"].astype(bool) != ((g[f"deg{deg}"][keep] + 0.5) < 0)]).max() < 1e-6


def test_ssim_oracle_matches_reference_python():
    g = np.load(os.path.join(GOLD, "ssim_ref.npz"))
    m, d1, d2, d3 = ssim_oracle.ssim_forward(g["img1"], g["img2"])
    assert abs(m.mean() - g["mean"]) < 1e-7
    grad = ssim_oracle.ssim_backward(g["img1"], g["img2"], np.full_like(m, 1.0 / m.size), d1, d2, d3)
    assert np.abs(grad - g["grad"]).max() < 1e-9


def test_knn_oracle_on_a_lattice():
    x, y, z = np.meshgrid(np.arange(5), np.arange(4), np.arange(3), indexing="ij")
    pts = np.stack([x, y, z], -1).reshape(-1, 3).astype(np.float32) * 2.0
    d = knn_oracle.dist2_knn3(pts)
    assert np.allclose(d, 4.0)   # three nearest lattice neighbours at distance 2
    pts2 = np.concatenate([pts, pts[:1]])   # a duplicate: distance 0 counts
    d2 = knn_oracle.dist2_knn3(pts2)
    assert np.isclose(d2[0], (0 + 4 + 4) / 3) and np.isclose(d2[-1], (0 + 4 + 4) / 3)


def test_binning_is_sorted_and_stable():
    scene, cam = S.blob_scene(800, seed=1), S.simple_camera(96, 80)
    scene.means3D[100:110] = scene.means3D[100]          # identical depth keys -> tie broken by Gaussian id
    scene.scales[100:110] = scene.scales[100]
    scene.rotations[100:110] = scene.rotations[100]
    f = fwd(scene, cam)
    keys, plist, ranges = f["keys"], f["point_list"], f["ranges"]
    assert f["num_rendered"] == int(f["tiles_touched"].sum()) == keys.shape[0]
    assert np.all(keys[1:] >= keys[:-1])
    same = keys[1:] == keys[:-1]
    assert same.any() and np.all(plist[1:][same] > plist[:-1][same])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        lo, hi = ranges[t]
        assert np.all(tiles[lo:hi] == t) and (lo == 0 or tiles[lo - 1] != t) and (hi == len(tiles) or tiles[hi] != t)
    empty = np.setdiff1d(np.arange(ranges.shape[0]), np.unique(tiles))
    assert np.all(ranges[empty] == 0)
    assert np.array_equal(f["point_offsets"], np.cumsum(f["tiles_touched"]))


def test_forward_semantics_small():
    """Background only affects colour; alpha = 1 - T; one opaque splat in front hides everything."""
    scene, cam = S.blob_scene(300, seed=2), S.simple_camera(64, 48)
    a = fwd(scene, cam, bg=(0, 0, 0))
    b = fwd(scene, cam, bg=(1, 0.5, 0.25))
    assert np.array_equal(a["depth"], b["depth"]) and np.array_equal(a["alpha"], b["alpha"])
    T = 1 - a["alpha"][0]
    assert np.allclose(b["color"] - a["color"], T[None] * np.array([1, 0.5, 0.25], np.float32)[:, None, None], atol=1e-6)
    assert a["alpha"].max() <= 1.0 and a["alpha"].min() >= 0.0
    assert a["n_contrib"].max() <= (a["ranges"][:, 1] - a["ranges"][:, 0]).max()
    # culling: a camera looking the other way sees nothing
    cam2 = S.simple_camera(64, 48, distance=-8.0)
    cam2.viewmatrix[:, 2] *= -1.0
    vis = O.mark_visible(scene.means3D, cam.viewmatrix)
    assert vis.sum() > 250


def test_backward_matches_finite_differences():
    """The analytic adjoint of the oracle (the reference's backward formulas) against central differences of its
    own forward, for parameters the forward is smooth in."""
    scene, cam = S.blob_scene(40, seed=5, spread=1.2, scale=0.35), S.simple_camera(48, 40, distance=6.0)
    scene.opacities[:] = np.clip(scene.opacities, 0.05, 0.6)     # stay away from the 0.99 clamp and T < 1e-4 stop
    cot = S.cotangents(cam.width, cam.height, seed=3)

    def loss(sc):
        f = fwd(sc, cam, bg=(0.2, 0.1, 0.3))
        return float(sum((c.astype(np.float64) * f[k].astype(np.float64)).sum()
                         for c, k in zip(cot, ("color", "depth", "norm", "alpha")))), f

    L0, f0 = loss(scene)
    g = O.backward(f0, *cot)
    rng = np.random.default_rng(0)
    vis = np.nonzero(f0["radii"] > 0)[0]
    assert len(vis) > 20
    checks = [("opacities", "opacity", 2e-3), ("shs", "sh", 2e-3)]
    for attr, gname, eps in checks:
        arr = getattr(scene, attr)
        for _ in range(6):
            i = int(rng.choice(vis))
            idx = (i,) + tuple(int(rng.integers(0, s)) for s in arr.shape[1:])
            old = arr[idx]
            arr[idx] = old + eps
            Lp, _ = loss(scene)
            arr[idx] = old - eps
            Lm, _ = loss(scene)
            arr[idx] = old
            num = (Lp - Lm) / (2 * eps)
            ana = float(g[gname][idx])
            assert abs(num - ana) <= 2e-2 * max(1.0, abs(ana), abs(num)), (attr, idx, num, ana)


def test_oracle_is_deterministic_forward():
    scene, cam = S.blob_scene(500, seed=8), S.simple_camera(80, 64)
    a, b = fwd(scene, cam), fwd(scene, cam)
    for k in ("color", "depth", "alpha", "radii", "point_list", "n_contrib"):
        assert np.array_equal(a[k], b[k])
