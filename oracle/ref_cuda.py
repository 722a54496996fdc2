"""ctypes face of oracle/_ref/libref_rasterizer.so — the UNMODIFIED reference CUDA rasterizer.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py may
import this module; the product (skyfall-gs_b200/) never does.  It is the
oracle the parity tests are anchored on, and the "reference CUDA path" timed
next to the product in bench.py.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_rasterizer.so")
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


class GeomView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("depths", "clamped", "means2D", "cov3D", "conic_opacity", "rgb", "norm3D",
                                          "tiles_touched", "point_offsets")]


class BinningView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("keys_unsorted", "keys", "list_unsorted", "list")]


class ImageView(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("n_contrib", "ranges")]


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError(f"{LIB_PATH} missing: run `python oracle/build_ref.py` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.ref_forward.argtypes = [ALLOC_FN, vp, ALLOC_FN, vp, ALLOC_FN, vp, ci, ci, ci, ci, vp, ci, ci] + [vp] * 4 + \
            [vp, cf, vp, vp, vp, vp, vp, vp, vp, cf, cf, cf, ci, vp, vp, vp, vp, vp, vp, ci]
        L.ref_forward.restype = ci
        L.ref_backward.argtypes = [ci, ci, ci, ci, ci, vp, ci, ci, vp, vp, vp, vp, cf, vp, vp, vp, vp, vp, vp, vp,
                                   cf, cf, cf, vp, vp, vp, vp] + [vp] * 18 + [ci]
        L.ref_backward.restype = ci
        L.ref_mark_visible.argtypes = [ci, vp, vp, vp, vp]
        L.ref_geom_layout.argtypes = [vp, ci, C.POINTER(GeomView)]
        L.ref_binning_layout.argtypes = [vp, ci, C.POINTER(BinningView)]
        L.ref_image_layout.argtypes = [vp, ci, C.POINTER(ImageView)]
        L.ref_knn.argtypes = [ci, vp, vp]
        L.ref_copy_d2d.argtypes = [vp, vp, C.c_size_t]
        L.ref_copy_d2d.restype = ci
        _lib = L
    return _lib


class _Arena:
    def __init__(self, device):
        self.device = device
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.cb = ALLOC_FN(self._alloc)

    def _alloc(self, _u, n):
        self.tensor = torch.empty(int(n), dtype=torch.uint8, device=self.device)
        return self.tensor.data_ptr()


def _p(t):
    return None if (t is None or t.numel() == 0) else t.data_ptr()


def _wrap(ptr, shape, dtype, device):
    """View raw device memory as a tensor copy (for the parity tests)."""
    n = 1
    for s in shape:
        n *= s
    out = torch.empty(shape, dtype=dtype, device=device)
    if n:
        rc = lib().ref_copy_d2d(out.data_ptr(), ptr, n * out.element_size())
        if rc != 0:
            raise RuntimeError(f"cudaMemcpy failed with {rc}")
    return out


def _order(dev):
    """The reference launches on the legacy default stream, like torch's default stream: same-stream ordering needs
    no host synchronisation.  Only when the caller switched torch to another stream is a full sync required."""
    cur = torch.cuda.current_stream(dev)
    if cur != torch.cuda.default_stream(dev):
        torch.cuda.synchronize(dev)


def forward(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, norm3D_precomp, extra,
            viewmatrix, projmatrix, tan_fovx, tan_fovy, kernel_size, H, W, sh, degree, campos, prefiltered=False,
            debug=False):
    """Returns dict with the reference's outputs and raw state buffers. The reference launches on the legacy
    default stream, so synchronise around the call."""
    L = lib()
    dev = means3D.device
    P = means3D.shape[0]
    F = extra.shape[1] if (extra is not None and extra.numel()) else 0
    M = sh.shape[1] if (sh is not None and sh.numel()) else 0
    f = dict(dtype=torch.float32, device=dev)
    out_color = torch.zeros((3, H, W), **f); out_depth = torch.zeros((1, H, W), **f)
    out_alpha = torch.zeros((1, H, W), **f); out_norm = torch.zeros((3, H, W), **f)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    out_extra = torch.zeros((F, H, W), **f) if F else torch.empty(0, **f)
    g, b, i = _Arena(dev), _Arena(dev), _Arena(dev)
    _order(dev)
    with torch.cuda.device(dev):
        R = L.ref_forward(g.cb, None, b.cb, None, i.cb, None, P, int(degree), M, F, _p(bg), W, H, _p(means3D), _p(sh),
                          _p(colors), _p(opacity), _p(scales), float(scale_modifier), _p(rotations),
                          _p(cov3D_precomp), _p(norm3D_precomp), _p(extra), _p(viewmatrix), _p(projmatrix),
                          _p(campos), float(tan_fovx), float(tan_fovy), float(kernel_size), int(prefiltered),
                          _p(out_color), _p(out_depth), _p(out_norm), _p(out_alpha), _p(out_extra), _p(radii),
                          int(debug))
    _order(dev)
    if R < 0:
        raise RuntimeError("reference forward threw")
    return dict(num_rendered=R, color=out_color, depth=out_depth, norm=out_norm, alpha=out_alpha, radii=radii,
                extra=out_extra, geom=g.tensor, binning=b.tensor, img=i.tensor)


def backward(bg, means3D, radii, colors, scales, rotations, extra, scale_modifier, cov3D_precomp, norm3D_precomp,
             viewmatrix, projmatrix, tan_fovx, tan_fovy, kernel_size, dL_color, dL_depth, dL_norm, dL_alpha, dL_extra,
             sh, degree, campos, geom, R, binning, img, out_alpha, debug=False):
    L = lib()
    dev = means3D.device
    P = means3D.shape[0]
    H, W = dL_color.shape[1], dL_color.shape[2]
    F = extra.shape[1] if (extra is not None and extra.numel()) else 0
    M = sh.shape[1] if (sh is not None and sh.numel()) else 0
    z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
    o = dict(means3D=z(P, 3), means2D=z(P, 3), colors=z(P, 3), depths=z(P, 1), conic=z(P, 2, 2), opacity=z(P, 1),
             cov3D=z(P, 6), norm3D=z(P, 3), sh=z(P, M, 3), scales=z(P, 3), rot=z(P, 4),
             extra=z(P, F) if F else torch.empty(0, device=dev))
    _order(dev)
    with torch.cuda.device(dev):
        rc = L.ref_backward(P, int(degree), M, int(R), F, _p(bg), W, H, _p(means3D), _p(sh), _p(colors), _p(scales),
                            float(scale_modifier), _p(rotations), _p(cov3D_precomp), _p(norm3D_precomp), _p(extra),
                            _p(viewmatrix), _p(projmatrix), _p(campos), float(tan_fovx), float(tan_fovy),
                            float(kernel_size), _p(radii), _p(geom), _p(binning), _p(img), _p(out_alpha),
                            _p(dL_color.contiguous()), _p(dL_depth.contiguous()), _p(dL_norm.contiguous()),
                            _p(dL_alpha.contiguous()), _p(dL_extra) if F else None,
                            _p(o["means2D"]), _p(o["conic"]), _p(o["opacity"]), _p(o["colors"]), _p(o["depths"]),
                            _p(o["means3D"]), _p(o["cov3D"]), _p(o["norm3D"]), _p(o["sh"]), _p(o["scales"]),
                            _p(o["rot"]), _p(o["extra"]), int(debug))
    _order(dev)
    if rc != 0:
        raise RuntimeError("reference backward threw")
    return o


def internals(fwd: dict, P: int, H: int, W: int):
    """Copies of the reference's intermediate arrays (the quantities the bit-exactness gate is stated on)."""
    L = lib()
    dev = fwd["geom"].device
    R = fwd["num_rendered"]
    gv, bv, iv = GeomView(), BinningView(), ImageView()
    L.ref_geom_layout(fwd["geom"].data_ptr(), P, C.byref(gv))
    L.ref_image_layout(fwd["img"].data_ptr(), W * H, C.byref(iv))
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    out = dict(
        depths=_wrap(gv.depths, (P,), torch.float32, dev), means2D=_wrap(gv.means2D, (P, 2), torch.float32, dev),
        cov3D=_wrap(gv.cov3D, (P, 6), torch.float32, dev),
        conic_opacity=_wrap(gv.conic_opacity, (P, 4), torch.float32, dev),
        rgb=_wrap(gv.rgb, (P, 3), torch.float32, dev), norm3D=_wrap(gv.norm3D, (P, 3), torch.float32, dev),
        clamped=_wrap(gv.clamped, (P, 3), torch.uint8, dev),
        tiles_touched=_wrap(gv.tiles_touched, (P,), torch.int32, dev),
        point_offsets=_wrap(gv.point_offsets, (P,), torch.int32, dev),
        n_contrib=_wrap(iv.n_contrib, (H * W,), torch.int32, dev),
        ranges=_wrap(iv.ranges, (tiles, 2), torch.int32, dev))
    if R > 0:
        L.ref_binning_layout(fwd["binning"].data_ptr(), R, C.byref(bv))
        out["keys"] = _wrap(bv.keys, (R,), torch.int64, dev)
        out["point_list"] = _wrap(bv.list, (R,), torch.int32, dev)
    else:
        out["keys"] = torch.empty(0, dtype=torch.int64, device=dev)
        out["point_list"] = torch.empty(0, dtype=torch.int32, device=dev)
    return out


def mark_visible(means3D, viewmatrix, projmatrix):
    L = lib()
    P = means3D.shape[0]
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    _order(None)
    L.ref_mark_visible(P, _p(means3D), _p(viewmatrix), _p(projmatrix), _p(present))
    _order(None)
    return present


def knn(points):
    L = lib()
    out = torch.zeros((points.shape[0],), dtype=torch.float32, device=points.device)
    _order(None)
    L.ref_knn(points.shape[0], _p(points.contiguous()), _p(out))
    _order(None)
    return out
