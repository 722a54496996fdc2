"""fused_ssim — drop-in for SSIM/fused_ssim/__init__.py:15-49 on top of libsfgs.so.

fused_ssim(img1, img2, padding="same"|"valid", train=True) -> scalar mean SSIM, differentiable w.r.t. img1 only
(11-tap sigma=1.5 separable Gaussian window, C1=0.01^2, C2=0.03^2, zero "same" padding).
"""
from __future__ import annotations

import os
import sys

import torch

_pkg_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)

from sfgs import native as _N  # noqa: E402

_N.lib()
allowed_padding = ["same", "valid"]


def fusedssim(C1, C2, img1, img2, train):
    """Same signature/returns as fused_ssim_cuda.fusedssim (SSIM/ssim.h:7-14)."""
    L = _N.lib()
    img1c, img2c = img1.contiguous(), img2.contiguous()
    B, CH, H, W = img1c.shape
    ssim_map = torch.empty_like(img1c)
    if train:
        d1, d2, d3 = torch.empty_like(img1c), torch.empty_like(img1c), torch.empty_like(img1c)
    else:
        d1 = d2 = d3 = torch.empty(0, device=img1c.device)
    with torch.cuda.device(img1c.device):
        _N.check(L.sfgs_fusedssim_forward(float(C1), float(C2), B, CH, H, W, img1c.data_ptr(), img2c.data_ptr(),
                                          int(bool(train)), ssim_map.data_ptr(),
                                          d1.data_ptr() if train else None, d2.data_ptr() if train else None,
                                          d3.data_ptr() if train else None,
                                          torch.cuda.current_stream(img1c.device).cuda_stream),
                 "sfgs_fusedssim_forward")
    return ssim_map, d1, d2, d3


def fusedssim_backward(C1, C2, img1, img2, dL_dmap, dm_dmu1, dm_dsigma1_sq, dm_dsigma12):
    """Same signature/returns as fused_ssim_cuda.fusedssim_backward (SSIM/ssim.h:16-26)."""
    L = _N.lib()
    img1c, img2c, g = img1.contiguous(), img2.contiguous(), dL_dmap.contiguous()
    B, CH, H, W = img1c.shape
    out = torch.empty_like(img1c)
    with torch.cuda.device(img1c.device):
        _N.check(L.sfgs_fusedssim_backward(float(C1), float(C2), B, CH, H, W, img1c.data_ptr(), img2c.data_ptr(),
                                           g.data_ptr(), dm_dmu1.data_ptr(), dm_dsigma1_sq.data_ptr(),
                                           dm_dsigma12.data_ptr(), out.data_ptr(),
                                           torch.cuda.current_stream(img1c.device).cuda_stream),
                 "sfgs_fusedssim_backward")
    return out


class FusedSSIMMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, C1, C2, img1, img2, padding="same", train=True):
        ssim_map, d1, d2, d3 = fusedssim(C1, C2, img1, img2, train)
        if padding == "valid":
            ssim_map = ssim_map[:, :, 5:-5, 5:-5]
        ctx.save_for_backward(img1.detach(), img2, d1, d2, d3)
        ctx.C1, ctx.C2, ctx.padding = C1, C2, padding
        return ssim_map

    @staticmethod
    def backward(ctx, opt_grad):
        img1, img2, d1, d2, d3 = ctx.saved_tensors
        dL_dmap = opt_grad
        if ctx.padding == "valid":
            dL_dmap = torch.zeros_like(img1)
            dL_dmap[:, :, 5:-5, 5:-5] = opt_grad
        grad = fusedssim_backward(ctx.C1, ctx.C2, img1, img2, dL_dmap, d1, d2, d3)
        return None, None, grad, None, None, None


def fused_ssim(img1, img2, padding="same", train=True):
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    assert padding in allowed_padding
    return FusedSSIMMap.apply(C1, C2, img1.contiguous(), img2, padding, train).mean()
