"""Stand-in for the reference's compiled module `diff_gauss._C` (RAST/ext.cpp:15-18).

Exposes the same three callables with the same positional signatures; the work
is done by libsfgs.so through the C ABI in include/sfgs.h.
"""
from sfgs.rasterizer import mark_visible, rasterize_gaussians, rasterize_gaussians_backward  # noqa: F401
from sfgs.native import lib as _lib

_lib()  # fail at import time if the CUDA library has not been built
