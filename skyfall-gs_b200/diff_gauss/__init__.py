"""diff_gauss — drop-in for the reference's autograd wrapper package.

Public surface (what gaussian_renderer/__init__.py:14 imports and what
RAST/diff_gauss/__init__.py:182-254 defines):

    GaussianRasterizationSettings   NamedTuple, same 14 fields in the same order
    GaussianRasterizer(settings)    nn.Module; .forward(...) and .markVisible(positions)
    rasterize_gaussians(...)        functional form

Returned tuple: (color[3,H,W], depth[1,H,W], unit normal[3,H,W], alpha[1,H,W],
radii[P] int32, extra[F,H,W] | empty).  `means2D` is a dummy [P,3] leaf whose
.grad receives (dL/dx_ndc * W/2, dL/dy_ndc * H/2, sum |.|) — the densification
statistics consumed by scene/gaussian_model.py:744-749.
"""
from __future__ import annotations

import os
import sys
from typing import NamedTuple

import torch
import torch.nn as nn

_pkg_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _pkg_root not in sys.path:  # make the sibling `sfgs` package importable when only this dir is on sys.path
    sys.path.insert(0, _pkg_root)

from . import _C  # noqa: E402

MAX_EXTRA_DIMS = 34  # RAST/cuda_rasterizer/auxiliary.h:20

# The reference normalises the blended normal map with torch ops inside the autograd graph
# (RAST/diff_gauss/__init__.py:48: ~4 elementwise kernels forward, ~8 backward, three passes over 3·H·W floats).
# Here the blend kernel writes the unit normals itself and the blend adjoint applies the normalize adjoint while
# it loads each pixel's cotangent.  Set SFGS_FUSE_NORMALIZE=0 to get the torch post-op back.
FUSE_NORMALIZE = os.environ.get("SFGS_FUSE_NORMALIZE", "1") != "0"


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    kernel_size: float
    subpixel_offset: torch.Tensor  # accepted for API compatibility; the reference never forwards it either
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _snapshot(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


def _guarded(fn, args, debug, dump_name, label):
    """Run a native call; in debug mode keep a CPU copy of the inputs and dump it if the call throws
    (same contract as the reference's snapshot_fw.dump / snapshot_bw.dump)."""
    if not debug:
        return fn(*args)
    saved = _snapshot(args)
    try:
        return fn(*args)
    except Exception:
        torch.save(saved, dump_name)
        print(f"\nAn error occured in {label}. Writing {dump_name} for debugging.\n")
        raise


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                norm3Ds_precomp, extra_attrs, raster_settings, fuse_normalize=False):
        rs = raster_settings
        n_extra = extra_attrs.shape[1] if extra_attrs.shape[0] != 0 else 0
        assert n_extra <= MAX_EXTRA_DIMS
        native_args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                       cov3Ds_precomp, norm3Ds_precomp, extra_attrs, n_extra, rs.viewmatrix, rs.projmatrix,
                       rs.tanfovx, rs.tanfovy, rs.kernel_size, rs.image_height, rs.image_width, sh, rs.sh_degree,
                       rs.campos, rs.prefiltered, rs.debug)
        ctx.fused = bool(fuse_normalize)
        if ctx.fused:
            (num_rendered, color, depth, norm, alpha, radii, extra, geom_buf, binning_buf, img_buf,
             norm_raw) = _guarded(_C.rasterize_gaussians_fused, native_args, rs.debug, "snapshot_fw.dump", "forward")
        else:
            (num_rendered, color, depth, norm, alpha, radii, extra,
             geom_buf, binning_buf, img_buf) = _guarded(_C.rasterize_gaussians, native_args, rs.debug,
                                                        "snapshot_fw.dump", "forward")
            norm_raw = torch.empty(0, device=means3D.device)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, norm3Ds_precomp, radii,
                              extra_attrs, sh, geom_buf, binning_buf, img_buf, alpha, norm_raw)
        ctx.mark_non_differentiable(radii)
        return color, depth, norm, alpha, radii, extra

    @staticmethod
    def backward(ctx, g_color, g_depth, g_norm, g_alpha, _g_radii, g_extra):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, norm3Ds_precomp, radii, extra_attrs, sh,
         geom_buf, binning_buf, img_buf, alpha, norm_raw) = ctx.saved_tensors
        native_args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, extra_attrs, rs.scale_modifier,
                       cov3Ds_precomp, norm3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                       rs.kernel_size, g_color, g_depth, g_norm, g_alpha, g_extra, sh, rs.sh_degree, rs.campos,
                       geom_buf, ctx.num_rendered, binning_buf, img_buf, alpha, rs.debug)
        bwd = _C.rasterize_gaussians_backward
        if ctx.fused:
            bwd = lambda *a: _C.rasterize_gaussians_backward_fused(*a, norm_raw=norm_raw)  # noqa: E731
        (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_norm3D, g_sh, g_scales, g_rot,
         g_extra_attrs) = _guarded(bwd, native_args, rs.debug, "snapshot_bw.dump", "backward")
        # order of forward's inputs; the settings tuple and the fusion flag get no gradient
        return (g_means3D, g_means2D, g_sh, g_colors, g_opac, g_scales, g_rot, g_cov3D, g_norm3D, g_extra_attrs,
                None, None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        norm3Ds_precomp, extra_attrs, raster_settings):
    color, depth, norm, alpha, radii, extra = _RasterizeGaussians.apply(
        means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, norm3Ds_precomp,
        extra_attrs, raster_settings, FUSE_NORMALIZE)
    if not FUSE_NORMALIZE:
        # the reference normalises the blended normal map in torch, inside the autograd graph (diff_gauss/__init__.py:48)
        norm = torch.nn.functional.normalize(norm, p=2, dim=0)
    return color, depth, norm, alpha, radii, extra


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of the points in front of the near plane (view depth > 0.2)."""
        rs = self.raster_settings
        with torch.no_grad():
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3Ds_precomp=None, norm3Ds_precomp=None, extra_attrs=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        has_sr = scales is not None or rotations is not None
        if ((scales is None or rotations is None) and cov3Ds_precomp is None) or (has_sr and cov3Ds_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if scales is None:
            raise ValueError('To support norm and depth prediction, scales == None is not allowed')
        if rotations is None:
            raise ValueError('To support norm and depth prediction, rotations == None is not allowed')

        def _or_empty(t):
            return torch.Tensor([]) if t is None else t

        return rasterize_gaussians(means3D, means2D, _or_empty(shs), _or_empty(colors_precomp), opacities, scales,
                                   rotations, _or_empty(cov3Ds_precomp), _or_empty(norm3Ds_precomp),
                                   _or_empty(extra_attrs), self.raster_settings)
