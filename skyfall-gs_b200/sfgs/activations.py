"""Fused per-Gaussian activations — the torch pre-ops of every `render()` call.

    opacity, scales, rotations = fused_activations(model._opacity, model._scaling, model._rotation, model.filter_3D)

replaces `pc.get_opacity_with_3D_filter`, `pc.get_scaling_with_3D_filter` and `pc.get_rotation`
(scene/gaussian_model.py:207-217,237-249; called at gaussian_renderer/__init__.py:61,71,72) including the
`.float()` casts of gaussian_renderer/__init__.py:137-138: one CUDA kernel forward, one backward, float32 results.
Gradients flow to the three raw parameters; `filter_3D` ([P,1] or [P], float64 like the reference's, float32
accepted) is a constant.
"""
from __future__ import annotations

import torch

from . import native as N


def _chk(t, name, shape_tail):
    if not t.is_cuda or t.dtype != torch.float32:
        raise ValueError(f"{name} must be a CUDA float32 tensor")
    if tuple(t.shape[1:]) != shape_tail:
        raise ValueError(f"{name} must have shape [P{''.join(',' + str(s) for s in shape_tail)}]")
    return t.contiguous()


class _FusedActivations(torch.autograd.Function):
    @staticmethod
    def forward(ctx, opacity_raw, scaling_raw, rotation_raw, filter_3D):
        o = _chk(opacity_raw, "opacity_raw", (1,))
        s = _chk(scaling_raw, "scaling_raw", (3,))
        q = _chk(rotation_raw, "rotation_raw", (4,))
        P = int(o.shape[0])
        if s.shape[0] != P or q.shape[0] != P or filter_3D.numel() != P:
            raise ValueError("fused_activations: inconsistent number of Gaussians")
        f = filter_3D.detach().reshape(P).to(device=o.device, dtype=torch.float64).contiguous()
        with torch.cuda.device(o.device):
            opacity, scales, rot = torch.empty_like(o), torch.empty_like(s), torch.empty_like(q)
            N.check(N.lib().sfgs_activations_forward(P, o.data_ptr(), s.data_ptr(), q.data_ptr(), f.data_ptr(),
                                                     opacity.data_ptr(), scales.data_ptr(), rot.data_ptr(),
                                                     torch.cuda.current_stream(o.device).cuda_stream),
                    "sfgs_activations_forward")
        ctx.save_for_backward(o, s, q, f)
        return opacity, scales, rot

    @staticmethod
    def backward(ctx, g_opacity, g_scales, g_rot):
        o, s, q, f = ctx.saved_tensors
        P = int(o.shape[0])
        go, gs, gq = (g.contiguous().float() for g in (g_opacity, g_scales, g_rot))
        with torch.cuda.device(o.device):
            d_o, d_s, d_q = torch.empty_like(o), torch.empty_like(s), torch.empty_like(q)
            N.check(N.lib().sfgs_activations_backward(P, o.data_ptr(), s.data_ptr(), q.data_ptr(), f.data_ptr(),
                                                      go.data_ptr(), gs.data_ptr(), gq.data_ptr(), d_o.data_ptr(),
                                                      d_s.data_ptr(), d_q.data_ptr(),
                                                      torch.cuda.current_stream(o.device).cuda_stream),
                    "sfgs_activations_backward")
        return d_o, d_s, d_q, None


def fused_activations(opacity_raw, scaling_raw, rotation_raw, filter_3D):
    """(opacity[P,1], scales[P,3], rotations[P,4]) — see the module docstring."""
    return _FusedActivations.apply(opacity_raw, scaling_raw, rotation_raw, filter_3D)
