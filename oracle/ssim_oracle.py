"""numpy restatement of the reference fused SSIM (TEST INFRASTRUCTURE).

Follows SSIM/ssim.cu:62-278 (forward: separable 11-tap sigma=1.5 window of ssim.cu:12-24, zero "same"
padding, the SSIM map and its three partial derivatives) and :286-427 (backward: dL/dimg1 =
conv(dm_dmu1*dL) + 2*img1*conv(dm_dsigma1_sq*dL) + img2*conv(dm_dsigma12*dL)).  Pinned by
tests/golden/ssim_ref.npz, generated from the reference's own PyTorch implementation
(utils/loss_utils.py:33-63, autograd for the gradient) by tests/golden/make_python_golden.py.
Only tests/ may import this module.
"""
import numpy as np

GAUSS = np.array([0.001028380123898387, 0.0075987582094967365, 0.036000773310661316, 0.10936068743467331,
                  0.21300552785396576, 0.26601171493530273, 0.21300552785396576, 0.10936068743467331,
                  0.036000773310661316, 0.0075987582094967365, 0.001028380123898387], dtype=np.float64)


def _blur(x):
    """separable 11x11 blur with zero padding over the last two axes (float64)."""
    H, W = x.shape[-2:]
    p = np.pad(x, [(0, 0)] * (x.ndim - 2) + [(5, 5), (5, 5)])
    h = sum(GAUSS[k] * p[..., :, k:k + W] for k in range(11))
    return sum(GAUSS[k] * h[..., k:k + H, :] for k in range(11))


def ssim_forward(img1, img2, C1=0.01 ** 2, C2=0.03 ** 2):
    x, y = img1.astype(np.float64), img2.astype(np.float64)
    mu1, mu2 = _blur(x), _blur(y)
    s1 = _blur(x * x) - mu1 * mu1
    s2 = _blur(y * y) - mu2 * mu2
    s12 = _blur(x * y) - mu1 * mu2
    A = mu1 * mu1 + mu2 * mu2 + C1
    B = s1 + s2 + C2
    C_ = 2 * mu1 * mu2 + C1
    D_ = 2 * s12 + C2
    m = (C_ * D_) / (A * B)
    dm_dmu1 = (mu2 * 2 * D_) / (A * B) - (mu2 * 2 * C_) / (A * B) - (mu1 * 2 * C_ * D_) / (A * A * B) + \
              (mu1 * 2 * C_ * D_) / (A * B * B)
    dm_ds1 = (-C_ * D_) / (A * B * B)
    dm_ds12 = (2 * C_) / (A * B)
    return m, dm_dmu1, dm_ds1, dm_ds12


def ssim_backward(img1, img2, dL_dmap, dm_dmu1, dm_ds1, dm_ds12):
    g = dL_dmap.astype(np.float64)
    return _blur(dm_dmu1 * g) + 2 * img1.astype(np.float64) * _blur(dm_ds1 * g) + img2.astype(np.float64) * _blur(dm_ds12 * g)
