"""Brute-force restatement of SimpleKNN::knn (KNN/simple_knn.cu:148-222) — TEST INFRASTRUCTURE.

For every point: the mean of the three smallest squared distances to the *other* points (by index; coincident
points count with distance 0).  The reference finds them with Morton-ordered boxes and box rejection, which is
exact, so a brute-force search returns the same set.  Distances are accumulated in float32 in the reference's
order ((dx*dx + dy*dy) + dz*dz up to FMA contraction), pinned on the GPU against oracle/_ref (ref_knn).
Only tests/ may import this module.
"""
import numpy as np


def dist2_knn3(points: np.ndarray, chunk: int = 2048) -> np.ndarray:
    p = np.ascontiguousarray(points, dtype=np.float32)
    P = p.shape[0]
    out = np.zeros(P, np.float32)
    for s in range(0, P, chunk):
        q = p[s:s + chunk]
        d = q[:, None, :].astype(np.float32) - p[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]).astype(np.float32)
        idx = np.arange(s, min(s + chunk, P))
        d2[np.arange(len(idx)), idx] = np.inf
        best = np.partition(d2, 2, axis=1)[:, :3]
        out[s:s + chunk] = ((best[:, 0] + best[:, 1] + best[:, 2]) / np.float32(3.0)).astype(np.float32)
    return out
