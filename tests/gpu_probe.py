"""First-light diagnostic (run on the GPU box): ours vs the reference CUDA rasterizer, prints mismatch statistics.
Not a pytest file; the assertions live in tests/test_gpu_*.py."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "skyfall-gs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import helpers as Hh
from sfgs import synthetic as S


def stats(name, a, b):
    a = a.float().flatten(); b = b.float().flatten()
    d = (a - b).abs()
    print(f"  {name:16s} max|d|={d.max().item():.3e} mean|d|={d.mean().item():.3e} max|ref|={b.abs().max().item():.3e} "
          f"n(d>1e-5)={(d > 1e-5).sum().item()} / {d.numel()}")


def bits_equal(name, a, b):
    if a.dtype.is_floating_point:
        a = a.contiguous().view(torch.int32); b = b.contiguous().view(torch.int32)
    ne = (a != b)
    print(f"  {name:16s} bit-mismatches={ne.sum().item()} / {ne.numel()}")
    return ne


def case(tag, scene, cam, dev, timing=False):
    print(f"== {tag}: P={scene.P} {cam.width}x{cam.height}")
    d = Hh.to_torch(scene, cam, dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    ref = Hh.run_ref_forward(d, cam, scene.sh_degree, bg)
    our = Hh.run_ours_forward(d, cam, scene.sh_degree, bg, debug=True)
    torch.cuda.synchronize()
    P, H, W = scene.P, cam.height, cam.width
    print(f"  num_rendered ref={ref['num_rendered']} ours={our['num_rendered']}  visible ref={(ref['radii']>0).sum().item()} ours={(our['radii']>0).sum().item()}")
    from oracle import ref_cuda
    ri = ref_cuda.internals(ref, P, H, W)
    oi = Hh.our_internals(our, P, H, W)
    vis = ref["radii"] > 0
    bits_equal("radii", our["radii"], ref["radii"])
    bits_equal("tiles_touched", oi["tiles_touched"], ri["tiles_touched"])
    bits_equal("depths[vis]", oi["depths"][vis], ri["depths"][vis])
    bits_equal("means2D[vis]", oi["means2D"][vis], ri["means2D"][vis])
    bits_equal("cov3D[vis]", oi["cov3D"][vis], ri["cov3D"][vis])
    bits_equal("conic_op[vis]", oi["conic_opacity"][vis], ri["conic_opacity"][vis])
    bits_equal("rgb[vis]", oi["rgb"][vis], ri["rgb"][vis])
    bits_equal("norm3D[vis]", oi["norm3D"][vis], ri["norm3D"][vis])
    if ref["num_rendered"] == our["num_rendered"]:
        bits_equal("ranges", oi["ranges"], ri["ranges"])
        bits_equal("point_list", oi["point_list"], ri["point_list"])
        bits_equal("keys", Hh.ref_style_keys(oi["ranges"], oi["keys"]), ri["keys"])
        bits_equal("n_contrib", oi["n_contrib"], ri["n_contrib"])
    for k in ("color", "depth", "norm", "alpha"):
        stats(k, our[k], ref[k])
        bits_equal(k + " bits", our[k], ref[k])
    cot = [torch.from_numpy(c).to(dev) for c in S.cotangents(W, H)]
    rb = Hh.run_ref_backward(d, cam, scene.sh_degree, bg, ref, cot)
    rb2 = Hh.run_ref_backward(d, cam, scene.sh_degree, bg, ref, cot)
    ob = Hh.run_ours_backward(d, cam, scene.sh_degree, bg, our, cot, debug=True)
    torch.cuda.synchronize()
    for k in ("means2D", "colors", "opacity", "means3D", "cov3D", "norm3D", "sh", "scales", "rot"):
        stats("d_" + k, ob[k], rb[k])
        stats("  ref-vs-ref", rb2[k], rb[k])
    if timing:
        def t(fn, n=10):
            for _ in range(3): fn()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(n): fn()
            torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
        tf_o = t(lambda: Hh.run_ours_forward(d, cam, scene.sh_degree, bg))
        tb_o = t(lambda: Hh.run_ours_backward(d, cam, scene.sh_degree, bg, our, cot))
        tf_r = t(lambda: Hh.run_ref_forward(d, cam, scene.sh_degree, bg))
        tb_r = t(lambda: Hh.run_ref_backward(d, cam, scene.sh_degree, bg, ref, cot))
        print(f"  wall ms: ours fwd {tf_o:.3f} bwd {tb_o:.3f} | ref fwd {tf_r:.3f} bwd {tb_r:.3f}")


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    case("blob-small", S.blob_scene(2000, seed=3), S.simple_camera(200, 136), dev)
    case("blob-10k", S.blob_scene(10000, seed=4), S.simple_camera(512, 512), dev)
    case("city-100k", S.city_scene(100_000, seed=0), S.jax004_camera(1920, 1080), dev, timing=True)
    case("city-1M", S.city_scene(1_000_000, seed=0), S.jax004_camera(1920, 1080), dev, timing=True)
    case("city-1M-orbit", S.city_scene(1_000_000, seed=0), S.orbit_camera(), dev, timing=True)
