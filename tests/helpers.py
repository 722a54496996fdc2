"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np
import torch

from sfgs import synthetic as S


def to_torch(scene: S.Scene, cam: S.Camera, device):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
    return dict(means3D=t(scene.means3D), scales=t(scene.scales), rotations=t(scene.rotations),
                opacities=t(scene.opacities), shs=t(scene.shs), viewmatrix=t(cam.viewmatrix),
                projmatrix=t(cam.projmatrix), campos=t(cam.campos))


def run_ours_forward(d, cam, sh_degree, bg, kernel_size=0.1, scale_modifier=1.0, colors=None, debug=False):
    from sfgs import rasterizer as R
    e = torch.empty(0, device=d["means3D"].device)
    sh = e if colors is not None else d["shs"]
    col = colors if colors is not None else e
    out = R.rasterize_gaussians(bg, d["means3D"], col, d["opacities"], d["scales"], d["rotations"], scale_modifier,
                                e, e, e, 0, d["viewmatrix"], d["projmatrix"], cam.tanfovx, cam.tanfovy, kernel_size,
                                cam.height, cam.width, sh, sh_degree, d["campos"], False, debug)
    keys = ("num_rendered", "color", "depth", "norm", "alpha", "radii", "extra", "geom", "binning", "img")
    return dict(zip(keys, out))


def run_ours_backward(d, cam, sh_degree, bg, fwd, cot, kernel_size=0.1, scale_modifier=1.0, colors=None, debug=False,
                      **kw):
    from sfgs import rasterizer as R
    e = torch.empty(0, device=d["means3D"].device)
    sh = e if colors is not None else d["shs"]
    col = colors if colors is not None else e
    out = R.rasterize_gaussians_backward(bg, d["means3D"], fwd["radii"], col, d["scales"], d["rotations"], e,
                                         scale_modifier, e, e, d["viewmatrix"], d["projmatrix"], cam.tanfovx,
                                         cam.tanfovy, kernel_size, cot[0], cot[1], cot[2], cot[3], e, sh, sh_degree,
                                         d["campos"], fwd["geom"], fwd["num_rendered"], fwd["binning"], fwd["img"],
                                         fwd["alpha"], debug, **kw)
    if kw.get("phase") == 1:
        return out                      # the [P,16] blend-adjoint sums
    keys = ("means2D", "colors", "opacity", "means3D", "cov3D", "norm3D", "sh", "scales", "rot", "extra")
    return dict(zip(keys, out))


def run_ref_forward(d, cam, sh_degree, bg, kernel_size=0.1, scale_modifier=1.0, colors=None):
    from oracle import ref_cuda
    e = torch.empty(0, device=d["means3D"].device)
    sh = e if colors is not None else d["shs"]
    col = colors if colors is not None else e
    return ref_cuda.forward(bg, d["means3D"], col, d["opacities"], d["scales"], d["rotations"], scale_modifier, e, e,
                            e, d["viewmatrix"], d["projmatrix"], cam.tanfovx, cam.tanfovy, kernel_size, cam.height,
                            cam.width, sh, sh_degree, d["campos"])


def run_ref_backward(d, cam, sh_degree, bg, fwd, cot, kernel_size=0.1, scale_modifier=1.0, colors=None):
    from oracle import ref_cuda
    e = torch.empty(0, device=d["means3D"].device)
    sh = e if colors is not None else d["shs"]
    col = colors if colors is not None else e
    return ref_cuda.backward(bg, d["means3D"], fwd["radii"], col, d["scales"], d["rotations"], e, scale_modifier, e,
                             e, d["viewmatrix"], d["projmatrix"], cam.tanfovx, cam.tanfovy, kernel_size, cot[0],
                             cot[1], cot[2], cot[3], e, sh, sh_degree, d["campos"], fwd["geom"],
                             fwd["num_rendered"], fwd["binning"], fwd["img"], fwd["alpha"])


def our_internals(fwd, P, H, W):
    """Unpack this library's scratch buffers into the reference's vocabulary."""
    import ctypes as C
    from sfgs import native as N
    L = N.lib()
    dev = fwd["geom"].device
    gv, iv, bv = N.GeomView(), N.ImageView(), N.BinningView()
    N.check(L.sfgs_geom_layout(fwd["geom"].data_ptr(), P, C.byref(gv)), "geom_layout")
    N.check(L.sfgs_image_layout(fwd["img"].data_ptr(), W, H, C.byref(iv)), "image_layout")
    # the binning layout depends on the capacity THIS forward used; it is recorded in the image header (words 2,3)
    hdr = fwd["img"][:32].view(torch.int32).cpu().numpy().astype(np.uint32)
    assert fwd["img"].data_ptr() % 256 == 0
    cap = int(hdr[2]) | (int(hdr[3]) << 32)
    N.check(L.sfgs_binning_layout(fwd["binning"].data_ptr(), cap, C.byref(bv)), "binning_layout")
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    R = fwd["num_rendered"]

    def view(ptr, base_t, shape, dtype):
        off = ptr - base_t.data_ptr()
        n = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        return base_t[off:off + n].view(dtype).view(*shape).clone()

    rec = view(gv.rec, fwd["geom"], (P, 16), torch.float32)
    out = dict(rec=rec, means2D=rec[:, 0:2], conic_opacity=torch.cat([rec[:, 2:5], rec[:, 5:6]], 1),
               depths=rec[:, 6], rgb=rec[:, 8:11], norm3D=rec[:, 11:14],
               cov3D=view(gv.cov3D, fwd["geom"], (P, 6), torch.float32),
               clamped=view(gv.clamped, fwd["geom"], (P,), torch.uint8),
               tiles_touched=view(gv.tiles_touched, fwd["geom"], (P,), torch.int32),
               n_contrib=view(iv.n_contrib, fwd["img"], (H * W,), torch.int32),
               ranges=view(iv.ranges, fwd["img"], (tiles, 2), torch.int32),
               keys=view(bv.keys, fwd["binning"], (R,), torch.int64),
               point_list=view(bv.point_list, fwd["binning"], (R,), torch.int32))
    return out


def ref_style_keys(ranges, keys_depth_id):
    """Rebuild the reference's (tile << 32 | depth_bits) keys from our per-tile (depth_bits << 32 | id) keys."""
    R = keys_depth_id.shape[0]
    counts = (ranges[:, 1] - ranges[:, 0]).long()
    tile_of = torch.repeat_interleave(torch.arange(ranges.shape[0], device=ranges.device), counts)
    assert tile_of.shape[0] == R
    depth_bits = (keys_depth_id >> 32) & 0xFFFFFFFF
    return (tile_of << 32) | depth_bits
