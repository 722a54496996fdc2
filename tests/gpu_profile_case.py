"""Run a few fwd+bwd iterations of one configuration (for ncu launch lists / captures). Not a pytest file."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "skyfall-gs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import helpers as Hh
from sfgs import synthetic as S

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="ours", choices=["ours", "ref"])
ap.add_argument("--P", type=int, default=1_000_000)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--cam", default="jax", choices=["jax", "orbit"])
ap.add_argument("--time", action="store_true", help="print CUDA-event ms per fwd+bwd iteration (median)")
ap.add_argument("--stages", action="store_true", help="ours only: print the median per-stage CUDA-event times")
ap.add_argument("--siblings", action="store_true", help="also run fused-ssim fwd/bwd, the appearance kernel and compute_3D_filter once per iteration")
a = ap.parse_args()
dev = torch.device("cuda:0")
scene = S.city_scene(a.P, seed=0)
cam = S.jax004_camera() if a.cam == "jax" else S.orbit_camera()
d = Hh.to_torch(scene, cam, dev)
bg = torch.zeros(3, device=dev)
cot = [torch.from_numpy(c).to(dev) for c in S.cotangents(cam.width, cam.height)]
times = []
stage_hist = {}
if a.stages and a.impl == "ours":
    from sfgs import native
for it in range(a.iters):
    if a.stages and a.impl == "ours":
        native.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if a.impl == "ours":
        f = Hh.run_ours_forward(d, cam, 3, bg)
        b = Hh.run_ours_backward(d, cam, 3, bg, f, cot)
    else:
        f = Hh.run_ref_forward(d, cam, 3, bg)
        b = Hh.run_ref_backward(d, cam, 3, bg, f, cot)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    if a.siblings and a.impl == "ours":
        import types
        import numpy as np
        import fused_ssim as fs
        from sfgs import appearance as AP
        from sfgs import filter3d as F3
        g = torch.Generator(device="cpu").manual_seed(2)
        img1 = torch.rand((1, 3, cam.height, cam.width), generator=g).to(dev)
        img2 = (img1 + 0.1 * torch.randn((1, 3, cam.height, cam.width), generator=g).to(dev)).clamp(0, 1)
        m, d1, d2, d3 = fs.fusedssim(1e-4, 9e-4, img1, img2, True)
        fs.fusedssim_backward(1e-4, 9e-4, img1, img2, torch.full_like(img1, 1e-6), d1, d2, d3)
        P = d["means3D"].shape[0]
        lin = [torch.nn.Linear(59, 128), torch.nn.Linear(128, 128), torch.nn.Linear(128, 6)]
        Wt = tuple(x.to(dev) for l in lin for x in (l.weight.detach(), l.bias.detach()))
        AP.fused_appearance_colors(d["shs"], torch.rand((P, 24), device=dev), torch.rand(32, device=dev), Wt, d["means3D"],
                                   d["campos"], 3)
        cams = []
        for k in range(16):
            c = S.orbit_camera(azimuth_deg=22.5 * k)
            w2c = c.viewmatrix.T.astype(np.float64)
            cams.append(types.SimpleNamespace(R=w2c[:3, :3].T.copy(), T=w2c[:3, 3].copy(), focal_x=c.width / (2 * c.tanfovx),
                                              focal_y=c.height / (2 * c.tanfovy), cx=0.0, cy=0.0, image_width=c.width, image_height=c.height))
        F3.compute_3D_filter(d["means3D"], cams)
        torch.cuda.synchronize()
    if a.stages and a.impl == "ours":
        for k, v in native.profile_read().items():
            stage_hist.setdefault(k, []).append(v[0] / v[1] if v[1] else 0.0)
print("done", f["num_rendered"])
if a.time:
    srt = sorted(times[1:] or times)
    V = int((f["radii"] > 0).sum())
    print(f"impl={a.impl} cam={a.cam} P={a.P} V={V} R={int(f['num_rendered'])} median_ms={srt[len(srt) // 2]:.3f} min_ms={srt[0]:.3f} "
          f"Mpix/s={cam.width * cam.height / srt[len(srt) // 2] / 1e3:.1f}")
if stage_hist:
    import statistics
    print("stages_ms " + " ".join(f"{k}={statistics.median(v[1:] or v):.4f}" for k, v in stage_hist.items()))
