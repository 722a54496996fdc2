"""Stand-in for the reference's compiled module `diff_gauss._C` (RAST/ext.cpp:15-18).

Exposes the same three callables with the same positional signatures; the work
is done by libsfgs.so through the C ABI in include/sfgs.h.
"""
from sfgs.rasterizer import mark_visible, rasterize_gaussians, rasterize_gaussians_backward  # noqa: F401
from sfgs.native import lib as _lib


def rasterize_gaussians_fused(*args):
    """`rasterize_gaussians` with the F.normalize post-op fused into the blend kernel; returns the reference's
    10-tuple (normal map already unit length) + the un-normalised blend needed by the backward."""
    return rasterize_gaussians(*args, fuse_normalize=True)


def rasterize_gaussians_backward_fused(*args, norm_raw):
    return rasterize_gaussians_backward(*args, norm_raw=norm_raw)


_lib()  # fail at import time if the CUDA library has not been built
