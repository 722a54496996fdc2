"""SURVEY.md 8f rank 3 — GaussianModel.compute_3D_filter fused into one kernel pass.  Golden vectors from the
reference's own method (tests/golden/make_filter3d_golden.py)."""
import os
import types

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter3d_ref.npz")


def _cams(table):
    return [types.SimpleNamespace(R=r[:9].reshape(3, 3), T=r[9:12], focal_x=float(r[12]), focal_y=float(r[13]), cx=float(r[14]),
                                  cy=float(r[15]), image_width=int(r[16]), image_height=int(r[17])) for r in table]


def test_oracle_matches_the_reference_method():
    from oracle import filter3d_oracle
    g = np.load(GOLD)
    got = filter3d_oracle.compute_3D_filter(g["xyz"], _cams(g["cams"]))
    assert got.shape == g["filter_3D"].shape and got.dtype == np.float64
    assert np.abs(got - g["filter_3D"]).max() <= 1e-12 * np.abs(g["filter_3D"]).max()


@pytest.mark.gpu
def test_fused_kernel_matches_the_reference_method(cuda_device):
    from sfgs.filter3d import compute_3D_filter
    g = np.load(GOLD)
    cams = _cams(g["cams"])
    got = compute_3D_filter(torch.from_numpy(g["xyz"]).to(cuda_device), cams)
    assert got.shape == (g["xyz"].shape[0], 1) and got.dtype == torch.float64
    err = np.abs(got.cpu().numpy() - g["filter_3D"])
    assert err.max() <= 1e-12 * np.abs(g["filter_3D"]).max(), float(err.max())
    # a subset of the cameras, and a single camera
    from oracle import filter3d_oracle
    for sub in (cams[:5], cams[7:8]):
        want = filter3d_oracle.compute_3D_filter(g["xyz"], sub)
        got = compute_3D_filter(torch.from_numpy(g["xyz"]).to(cuda_device), sub).cpu().numpy()
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
