"""Host-side mirror of the reference's nn/losses.py: same names, argument order, return types
(rank-0 float32 tensor) and error behaviour, backed by the sm_100a kernels through the C ABI.

Differentiable w.r.t. the prediction operand only -- `y` of relaxed_emd / moment_matching and `x`
of self_similarity -- which are the only operands that receive gradients in the reference driver
(run_strotss.py:24,35-39); asking for the other gradient raises.
"""
from __future__ import annotations

import torch

from . import _lib
from .runtime import reshape_2d, shared_handle

__all__ = ["relaxed_emd", "moment_matching", "self_similarity", "dist_metrics", "reshape_2d"]

# keys only: the distance matrices are never materialised (nn/losses.py:27-28)
dist_metrics = {"cosine": "cosine", "l2": "l2", "both": "both"}


def _no_target_grad(name):
    raise NotImplementedError(
        f"{name}: gradient w.r.t. the target operand is outside the hot path (the reference driver "
        "never requests it, run_strotss.py:30,95-96)")


class _RelaxedEMD(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, distance):
        need = y.requires_grad
        out, grad, _, _ = shared_handle(y.device).relaxed_emd(x.detach(), y.detach(), distance, need)
        ctx.x_needs = x.requires_grad
        ctx.save_for_backward(grad if need else None)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        if ctx.x_needs:
            _no_target_grad("relaxed_emd")
        (grad,) = ctx.saved_tensors
        return None, (grad * g if grad is not None else None), None


class _MomentMatching(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        need = y.requires_grad
        out, grad = shared_handle(y.device).moment_matching(x.detach(), y.detach(), need)
        ctx.x_needs = x.requires_grad
        ctx.save_for_backward(grad if need else None)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        if ctx.x_needs:
            _no_target_grad("moment_matching")
        (grad,) = ctx.saved_tensors
        return None, (grad * g if grad is not None else None)


class _SelfSimilarity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        need = x.requires_grad
        out, grad = shared_handle(x.device).self_similarity(x.detach(), y.detach(), need)
        ctx.y_needs = y.requires_grad
        ctx.save_for_backward(grad if need else None)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        if ctx.y_needs:
            _no_target_grad("self_similarity")
        (grad,) = ctx.saved_tensors
        return (grad * g if grad is not None else None), None


def relaxed_emd(x: torch.Tensor, y: torch.Tensor, distance: str = "cosine") -> torch.Tensor:
    """nn/losses.py:69-80.  max(mean_i min_j C_ij, mean_j min_i C_ij) with C = dist_metrics[distance](x, y)."""
    x = reshape_2d(x)
    y = reshape_2d(y)
    if distance not in dist_metrics:
        raise KeyError(distance)
    return _RelaxedEMD.apply(x, y, distance)


def moment_matching(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """nn/losses.py:39-52.  mae(cov_x, cov_y) + mae(mean_x, mean_y)."""
    return _MomentMatching.apply(reshape_2d(x), reshape_2d(y))


def self_similarity(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """nn/losses.py:55-66.  N * mae(Xd / colsum(Xd), Yd / colsum(Yd))."""
    return _SelfSimilarity.apply(reshape_2d(x), reshape_2d(y))
