"""Adaptive density control on the library's kernels.

Drop-ins for the two places the reference's training loop touches per-Gaussian statistics and topology:

    # train.py:314-315 — every iteration below densify_until_iter
    add_densification_stats(gaussians, viewspace_point_tensor, radii)          # replaces both statements
    # train.py:321 — every densification_interval iterations
    densify_and_prune(gaussians, opt.densify_grad_threshold, 0.005, scene.cameras_extent, size_threshold)

`gaussians` is the reference's GaussianModel (scene/gaussian_model.py), used as it is: its parameters, Adam state,
statistics tensors and `percent_dense` are read and replaced exactly as `densify_and_prune` (:694-735),
`densification_postfix` (:626-651) and `prune_points` (:584-601) leave them, including the row order.  The functional
forms below them take plain tensors.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import nn

from . import native as N

FIELDS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")      # optimizer group names, gaussian_model.py:358-363
ATTRS = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling",
         "rotation": "_rotation", "embeddings": "_embeddings"}


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise TypeError(f"{what} must be a contiguous CUDA float32 tensor")
    return t


@torch.no_grad()
def densification_stats(viewspace_grad: torch.Tensor, radii: torch.Tensor, max_radii2D: torch.Tensor,
                        grad_accum: torch.Tensor, grad_accum_abs: torch.Tensor, grad_accum_abs_max: torch.Tensor,
                        denom: torch.Tensor) -> None:
    """In place, for every Gaussian with radii > 0 (train.py:314-315, gaussian_model.py:744-749)."""
    L = N.lib()
    P = int(radii.shape[0])
    if tuple(viewspace_grad.shape) != (P, 4):
        raise ValueError(f"viewspace gradient is {tuple(viewspace_grad.shape)}, expected ({P}, 4)")
    if radii.dtype != torch.int32 or not radii.is_contiguous():
        raise TypeError("radii must be a contiguous int32 tensor")
    for t, what in ((viewspace_grad, "viewspace gradient"), (max_radii2D, "max_radii2D"), (grad_accum, "xyz_gradient_accum"),
                    (grad_accum_abs, "xyz_gradient_accum_abs"), (grad_accum_abs_max, "xyz_gradient_accum_abs_max"), (denom, "denom")):
        _f32(t, what)
    for t in (max_radii2D, grad_accum, grad_accum_abs, grad_accum_abs_max, denom):
        if t.numel() != P:
            raise ValueError("statistics tensors must hold one value per Gaussian")
    dev = radii.device
    with torch.cuda.device(dev):
        N.check(L.sfgs_densification_stats(P, viewspace_grad.data_ptr(), radii.data_ptr(), max_radii2D.data_ptr(),
                                           grad_accum.data_ptr(), grad_accum_abs.data_ptr(), grad_accum_abs_max.data_ptr(),
                                           denom.data_ptr(), _stream(dev)), "sfgs_densification_stats")


def add_densification_stats(pc, viewspace_point_tensor: torch.Tensor, radii: torch.Tensor) -> None:
    densification_stats(viewspace_point_tensor.grad, radii, pc.max_radii2D, pc.xyz_gradient_accum, pc.xyz_gradient_accum_abs,
                        pc.xyz_gradient_accum_abs_max, pc.denom)


def gradient_thresholds(grad_accum, grad_accum_abs, denom, max_grad: float) -> torch.Tensor:
    """The quantile Q of gaussian_model.py:695-713 as a device scalar (same torch calls, same fall-backs)."""
    grads = grad_accum / denom
    grads[grads.isnan()] = 0.0
    grads_abs = grad_accum_abs / denom
    grads_abs[grads_abs.isnan()] = 0.0
    fallback = torch.full((), 0.99, dtype=torch.float32, device=grad_accum.device)
    if grads_abs.numel() > 0 and not torch.isinf(grads_abs).any() and not torch.isnan(grads_abs).any():
        ratio = (torch.norm(grads, dim=-1) >= max_grad).float().mean()
        try:
            return torch.quantile(grads_abs.reshape(-1), 1 - ratio).float()
        except Exception:                                   # torch.quantile refuses more than 16 M elements (:709-711)
            return fallback
    return fallback


@torch.no_grad()
def densify_tensors(params: Dict[str, torch.Tensor], exp_avg: Optional[Dict[str, torch.Tensor]],
                    exp_avg_sq: Optional[Dict[str, torch.Tensor]], grad_accum, grad_accum_abs, denom, *, max_grad: float,
                    min_opacity: float, extent: float, max_screen_size, percent_dense: float,
                    noise: Optional[torch.Tensor] = None, abs_threshold: Optional[torch.Tensor] = None,
                    extra: Sequence[str] = ()):
    """densify_and_prune on plain tensors.  params: name -> [P, ...] for FIELDS (+ `extra` names); the Adam moment dicts
    are both given or both None.  Returns (new_params, new_exp_avg, new_exp_avg_sq, totals) with totals =
    {K, KC, S, KS, C} (see include/sfgs.h)."""
    L = N.lib()
    names = list(FIELDS) + list(extra)
    xyz = params["xyz"]
    dev = xyz.device
    P = int(xyz.shape[0])
    for n in names:
        _f32(params[n], n)
        if int(params[n].shape[0]) != P:
            raise ValueError(f"{n} has {params[n].shape[0]} rows, xyz has {P}")
    if (exp_avg is None) != (exp_avg_sq is None):
        raise ValueError("give both Adam moments or neither")
    widths = [math.prod(params[n].shape[1:]) for n in names]
    if abs_threshold is None:
        abs_threshold = gradient_thresholds(grad_accum, grad_accum_abs, denom, max_grad)
    abs_threshold = abs_threshold.reshape(1).float().contiguous()
    nblk = L.sfgs_densify_plan_blocks(P)
    action = torch.empty(max(P, 1), dtype=torch.uint8, device=dev)
    offsets = torch.empty(max(5 * nblk, 1), dtype=torch.int32, device=dev)
    totals_dev = torch.empty(8, dtype=torch.int32, device=dev)
    totals = (C.c_int * 5)()
    screen_test = bool(max_screen_size)
    with torch.cuda.device(dev):
        N.check(L.sfgs_densify_plan(P, _f32(grad_accum, "xyz_gradient_accum").data_ptr(), _f32(grad_accum_abs, "xyz_gradient_accum_abs").data_ptr(),
                                    _f32(denom, "denom").data_ptr(), params["scaling"].data_ptr(), params["opacity"].data_ptr(),
                                    float(max_grad), abs_threshold.data_ptr(), float(percent_dense * extent), float(min_opacity),
                                    int(screen_test), float(max_screen_size) if screen_test else 0.0, float(0.1 * extent),
                                    action.data_ptr(), offsets.data_ptr(), totals_dev.data_ptr(), totals, _stream(dev)),
                "sfgs_densify_plan")
    K, KC, S, KS, Csel = (int(v) for v in totals)
    newP = K + KC + 2 * KS
    if noise is None:
        # the reference draws torch.normal(mean=0, std=stds) over [2S, 3]: the same generator calls as this randn
        noise = torch.randn((2 * S, 3), device=dev, dtype=torch.float32)
    if tuple(noise.shape) != (2 * S, 3):
        raise ValueError(f"noise is {tuple(noise.shape)}, the plan has {S} split sources: expected ({2 * S}, 3)")
    noise = _f32(noise.contiguous(), "noise")

    def out_like(t):
        return torch.empty((newP,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev)

    new_p = {n: out_like(params[n]) for n in names}
    new_m = {n: out_like(params[n]) for n in names} if exp_avg is not None else None
    new_v = {n: out_like(params[n]) for n in names} if exp_avg is not None else None

    def ptrs(d):
        if d is None:
            return None
        return (C.c_void_p * len(names))(*[_f32(d[n], n).data_ptr() for n in names])

    if P:
        with torch.cuda.device(dev):
            N.check(L.sfgs_densify_apply(P, action.data_ptr(), offsets.data_ptr(), totals, noise.data_ptr() if S else None,
                                         len(names), (C.c_int * len(names))(*widths), ptrs(params), ptrs(exp_avg), ptrs(exp_avg_sq),
                                         ptrs(new_p), ptrs(new_m), ptrs(new_v), _stream(dev)), "sfgs_densify_apply")
    return new_p, new_m, new_v, dict(K=K, KC=KC, S=S, KS=KS, C=Csel, P=P, new_P=newP)


@torch.no_grad()
def densify_and_prune(pc, max_grad: float, min_opacity: float, extent: float, max_screen_size,
                      noise: Optional[torch.Tensor] = None) -> Tuple[int, int, int]:
    """GaussianModel.densify_and_prune (gaussian_model.py:694-735) for the reference's model object `pc`."""
    extra = ("embeddings",) if getattr(pc, "appearance_enabled", False) else ()
    names = list(FIELDS) + list(extra)
    groups = {g["name"]: g for g in pc.optimizer.param_groups if g.get("name") in names}
    missing = [n for n in names if n not in groups]
    if missing:
        raise ValueError(f"optimizer has no parameter group named {missing}")
    params, m, v = {}, {}, {}
    for n in names:
        p = groups[n]["params"][0]
        params[n] = p.data
        st = pc.optimizer.state.get(p, None)
        if st is not None and "exp_avg" in st:
            m[n], v[n] = st["exp_avg"], st["exp_avg_sq"]
    if m and len(m) != len(names):
        raise ValueError("some parameter groups have Adam state and some do not")
    new_p, new_m, new_v, t = densify_tensors(params, m or None, v or None, pc.xyz_gradient_accum, pc.xyz_gradient_accum_abs, pc.denom,
                                             max_grad=max_grad, min_opacity=min_opacity, extent=extent,
                                             max_screen_size=max_screen_size, percent_dense=pc.percent_dense, noise=noise,
                                             extra=extra)
    for n in names:                                           # cat_tensors_to_optimizer / _prune_optimizer, :564-624
        g = groups[n]
        old = g["params"][0]
        st = pc.optimizer.state.get(old, None)
        fresh = nn.Parameter(new_p[n].requires_grad_(True))
        if st is not None:
            if new_m is not None:
                st["exp_avg"], st["exp_avg_sq"] = new_m[n], new_v[n]
            del pc.optimizer.state[old]
            pc.optimizer.state[fresh] = st
        g["params"][0] = fresh
        setattr(pc, ATTRS[n], fresh)
    dev = new_p["xyz"].device
    newP = t["new_P"]
    pc.xyz_gradient_accum = torch.zeros((newP, 1), device=dev)
    pc.xyz_gradient_accum_abs = torch.zeros((newP, 1), device=dev)
    pc.xyz_gradient_accum_abs_max = torch.zeros((newP, 1), device=dev)
    pc.denom = torch.zeros((newP, 1), device=dev)
    pc.max_radii2D = torch.zeros((newP,), device=dev)
    pruned = (t["P"] + t["C"] + t["S"]) - newP               # `split - prune` of :733-735
    return t["C"], t["S"], pruned
