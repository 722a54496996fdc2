"""Fused per-Gaussian activations: oracle vs the reference's own GaussianModel (golden vectors),
CUDA kernels vs both."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "activations_ref.npz")


def _close(a, b, rtol, atol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= atol + rtol * np.abs(b).max(), float(np.abs(a - b).max())


def test_oracle_matches_reference_gaussian_model():
    from oracle import activations_oracle as A
    g = np.load(GOLD)
    op, sc, rt = A.forward(g["opacity_raw"], g["scaling_raw"], g["rotation_raw"], g["filter_3D"])
    # forward: same float32/float64 promotion as torch -> equal to a couple of float32 ulps
    assert np.abs(op - g["opacity"]).max() <= 2e-7
    assert np.abs(sc - g["scales"]).max() <= 2e-7 * max(1.0, np.abs(g["scales"]).max())
    assert np.abs(rt - g["rotations"]).max() <= 2e-7
    d_o, d_s, d_q = A.backward(g["opacity_raw"], g["scaling_raw"], g["rotation_raw"], g["filter_3D"],
                               g["cot_opacity"], g["cot_scales"], g["cot_rotations"])
    for got, key in ((d_o, "g_opacity_raw"), (d_s, "g_scaling_raw"), (d_q, "g_rotation_raw")):
        ok, err = _close(got, g[key], 2e-6, 1e-6)
        assert ok, (key, err)
    # no filter -> coef == 1, scales == exp(s)
    assert np.allclose(sc[:8], np.exp(g["scaling_raw"][:8]), rtol=1e-6)


@pytest.mark.gpu
def test_cuda_activations_match_reference_and_oracle(cuda_device):
    import torch
    from oracle import activations_oracle as A
    from sfgs.activations import fused_activations
    dev = cuda_device
    g = np.load(GOLD)
    t = lambda k, **kw: torch.tensor(g[k], device=dev, **kw)  # noqa: E731
    o, s, q = t("opacity_raw", requires_grad=True), t("scaling_raw", requires_grad=True), t("rotation_raw", requires_grad=True)
    op, sc, rt = fused_activations(o, s, q, t("filter_3D"))
    assert op.dtype == torch.float32 and op.shape == (o.shape[0], 1) and sc.shape == s.shape and rt.shape == q.shape
    assert (op.detach().cpu().numpy() - g["opacity"]).__abs__().max() <= 3e-7
    assert np.abs(sc.detach().cpu().numpy() - g["scales"]).max() <= 3e-7 * max(1.0, np.abs(g["scales"]).max())
    assert np.abs(rt.detach().cpu().numpy() - g["rotations"]).max() <= 3e-7
    ((op * t("cot_opacity")).sum() + (sc * t("cot_scales")).sum() + (rt * t("cot_rotations")).sum()).backward()
    for leaf, key in ((o, "g_opacity_raw"), (s, "g_scaling_raw"), (q, "g_rotation_raw")):
        ok, err = _close(leaf.grad.cpu().numpy(), g[key], 5e-6, 1e-6)
        assert ok, (key, err)
    # a larger seeded case against the oracle, float32 filter accepted, [P] filter accepted
    rng = np.random.default_rng(5)
    P = 100_003
    o2 = rng.normal(0, 2, (P, 1)).astype(np.float32); s2 = rng.normal(-1, 1, (P, 3)).astype(np.float32)
    q2 = rng.normal(0, 1, (P, 4)).astype(np.float32); f2 = np.abs(rng.normal(0.1, 0.2, P)).astype(np.float32)
    want = A.forward(o2, s2, q2, f2.astype(np.float64))
    got = fused_activations(torch.tensor(o2, device=dev), torch.tensor(s2, device=dev), torch.tensor(q2, device=dev),
                            torch.tensor(f2, device=dev))
    for a, b in zip(got, want):
        assert np.abs(a.cpu().numpy() - b).max() <= 3e-7 * max(1.0, np.abs(b).max())
