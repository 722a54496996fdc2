"""Stand-in for `simple_knn._C` (KNN/ext.cpp:15-17): distCUDA2(points[P,3] cuda f32) -> f32[P],
the mean squared distance to the 3 nearest neighbours (KNN/simple_knn.cu:186-222)."""
from __future__ import annotations

import os
import sys

import torch

_pkg_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)

from sfgs import native as _N  # noqa: E402

_N.lib()


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    L = _N.lib()
    pts = points.contiguous()
    if pts.dtype != torch.float32 or not pts.is_cuda:
        raise TypeError("distCUDA2 expects a CUDA float32 tensor [P,3]")
    P = int(pts.shape[0])
    out = torch.zeros((P,), dtype=torch.float32, device=pts.device)
    if P == 0:
        return out
    held = []

    def _alloc(_u, n):
        t = torch.empty(int(n), dtype=torch.uint8, device=pts.device)
        held.append(t)
        return t.data_ptr()

    cb = _N.ALLOC_FN(_alloc)
    with torch.cuda.device(pts.device):
        _N.check(L.sfgs_dist2_knn3(P, pts.data_ptr(), out.data_ptr(), cb, None,
                                   torch.cuda.current_stream(pts.device).cuda_stream), "sfgs_dist2_knn3")
    # `held` (the scratch) was allocated on the current stream and the kernels were queued on it: returning the blocks
    # to torch's caching allocator here is stream-ordered, exactly as for any torch op — no host synchronisation.
    return out
