"""BASELINE.json configs[2] ("train.py-equivalent loop") as a throughput measurement — needs a GPU.  Not a pytest file.

    python tests/train_loop_bench.py [--P 1000000] [--iters 200] [--impl ours|reference|both]

One iteration = what the reference's train.py does per step around the rasterizer (train.py:195-260):
activations -> render one of 8 orbit views at 1920x1080 -> 0.8 L1 + 0.2 (1 - SSIM) against an image of a hidden
synthetic scene -> backward -> Adam step on (xyz, scaling, rotation, opacity, SH).  Synthetic data (no JAX_068 here).
  ours       fused activations + drop-in diff_gauss (fused normal post-op) + fused_ssim of this repository
  reference  the same loop with the unmodified reference CUDA rasterizer (oracle/_ref) behind a small autograd
             wrapper and torch activations; fused_ssim is this repository's in both arms (the reference's own
             fused-ssim extension is not built here), so the difference is the rasterizer path + its pre/post-ops.
Prints iterations/s (CUDA events over the timed iterations, after 10 warm-up iterations).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tests/: it uses oracle/_ref)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "skyfall-gs_b200"))
import numpy as np
import torch

from sfgs import synthetic as S


class _RefRasterize(torch.autograd.Function):
    """Autograd wrapper around the unmodified reference CUDA rasterizer (oracle/_ref) — measurement harness only."""

    @staticmethod
    def forward(ctx, means3D, opacity, scales, rotations, shs, bg, view, proj, campos, cam):
        from oracle import ref_cuda
        e = torch.empty(0, device=means3D.device)
        f = ref_cuda.forward(bg, means3D, e, opacity, scales, rotations, 1.0, e, e, e, view, proj, cam.tanfovx,
                             cam.tanfovy, 0.1, cam.height, cam.width, shs, 3, campos)
        ctx.f, ctx.cam = f, cam
        ctx.save_for_backward(means3D, scales, rotations, shs, bg, view, proj, campos)
        return f["color"], f["depth"], f["norm"], f["alpha"]

    @staticmethod
    def backward(ctx, g_color, g_depth, g_norm, g_alpha):
        from oracle import ref_cuda
        means3D, scales, rotations, shs, bg, view, proj, campos = ctx.saved_tensors
        f, cam = ctx.f, ctx.cam
        e = torch.empty(0, device=means3D.device)
        g = ref_cuda.backward(bg, means3D, f["radii"], e, scales, rotations, e, 1.0, e, e, view, proj, cam.tanfovx,
                              cam.tanfovy, 0.1, g_color.contiguous(), g_depth.contiguous(), g_norm.contiguous(),
                              g_alpha.contiguous(), e, shs, 3, campos, f["geom"], f["num_rendered"], f["binning"],
                              f["img"], f["alpha"])
        return g["means3D"], g["opacity"], g["scales"], g["rot"], g["sh"], None, None, None, None, None


def run(impl, scene, cams, targets, iters, dev):
    from diff_gauss import GaussianRasterizationSettings, GaussianRasterizer
    from fused_ssim import fused_ssim
    from sfgs.activations import fused_activations
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    rng = np.random.default_rng(1)
    P = scene.P
    xyz = t((scene.means3D + rng.normal(0, 0.05, scene.means3D.shape)).astype(np.float32)).requires_grad_(True)
    scaling = torch.log(t(scene.scales)).requires_grad_(True)
    rotation = t(scene.rotations).clone().requires_grad_(True)
    opacity = torch.logit(t(scene.opacities).clamp(1e-4, 1 - 1e-4)).reshape(P, 1).clone().requires_grad_(True)
    shs = (t(scene.shs) * 0.9).requires_grad_(True)
    filter_3D = torch.full((P, 1), 0.05, dtype=torch.float64, device=dev)
    opt = torch.optim.Adam([{"params": [xyz], "lr": 1.6e-4}, {"params": [scaling], "lr": 5e-3},
                            {"params": [rotation], "lr": 1e-3}, {"params": [opacity], "lr": 5e-2},
                            {"params": [shs], "lr": 2.5e-3}], eps=1e-15)
    bg = torch.zeros(3, device=dev)
    cam_t = [(c, t(c.viewmatrix), t(c.projmatrix), t(c.campos)) for c in cams]
    m2d = torch.zeros((P, 3), device=dev, requires_grad=True)

    def step(i):
        cam, view, proj, campos = cam_t[i % len(cam_t)]
        if impl == "ours":
            op, sc, rot = fused_activations(opacity, scaling, rotation, filter_3D)
            rs = GaussianRasterizationSettings(cam.height, cam.width, cam.tanfovx, cam.tanfovy, 0.1,
                                               torch.zeros(1, device=dev), bg, 1.0, view, proj, 3, campos, False, False)
            color, depth, norm, alpha, radii, _ = GaussianRasterizer(rs)(xyz, m2d, op, shs=shs, scales=sc, rotations=rot)
        else:
            s = torch.exp(scaling)                                   # scene/gaussian_model.py:207-249 in torch
            sq = torch.square(s)
            a = sq + torch.square(filter_3D)
            sc = torch.sqrt(a).float()
            op = (torch.sigmoid(opacity) * torch.sqrt(sq.prod(dim=1) / a.prod(dim=1))[..., None]).float()
            rot = torch.nn.functional.normalize(rotation)
            color, depth, norm, alpha = _RefRasterize.apply(xyz, op, sc, rot, shs, bg, view, proj, campos, cam)
            norm = torch.nn.functional.normalize(norm, p=2, dim=0)   # diff_gauss/__init__.py:48
        gt = targets[i % len(targets)]
        loss = 0.8 * torch.nn.functional.l1_loss(color, gt) + 0.2 * (1.0 - fused_ssim(color[None], gt[None]))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for i in range(10):
        step(i)
    torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    first = last = None
    for i in range(iters):
        loss = step(i)
        if i == 0:
            first = loss.detach()
        last = loss.detach()
    b.record()
    torch.cuda.synchronize(dev)
    ms = a.elapsed_time(b)
    return {"impl": impl, "iters": iters, "it_per_s": round(iters / (ms / 1e3), 1), "ms_per_it": round(ms / iters, 3),
            "loss_first": round(float(first), 5), "loss_last": round(float(last), 5)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--impl", default="both", choices=["ours", "reference", "both"])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    scene = S.city_scene(a.P, seed=0, sh_degree=3)
    cams = [S.orbit_camera(azimuth_deg=45.0 * k) for k in range(8)]
    from sfgs import rasterizer as R
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    e = torch.empty(0, device=dev)
    targets = []
    with torch.no_grad():
        for c in cams:
            f = R.rasterize_gaussians(torch.zeros(3, device=dev), t(scene.means3D), e, t(scene.opacities), t(scene.scales),
                                      t(scene.rotations), 1.0, e, e, e, 0, t(c.viewmatrix), t(c.projmatrix), c.tanfovx,
                                      c.tanfovy, 0.1, c.height, c.width, t(scene.shs), 3, t(c.campos), False, False)
            targets.append(f[1].clone())
    out = []
    from oracle import ref_cuda
    for impl in (["ours", "reference"] if a.impl == "both" else [a.impl]):
        if impl == "reference" and not ref_cuda.available():
            out.append({"impl": "reference", "unavailable": "oracle/_ref not built"})
            continue
        out.append(run(impl, scene, cams, targets, a.iters, dev))
    print(json.dumps({"config": f"train-loop: {a.P} Gaussians, 8 orbit views 1920x1080, SH 3, L1+SSIM, Adam", "results": out}))


if __name__ == "__main__":
    main()
