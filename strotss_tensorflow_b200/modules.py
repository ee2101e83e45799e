"""Mirror of the loss wrappers in run_strotss.py:21-40 plus the fused per-iteration evaluation.

ContentLoss / StyleLoss keep the reference's constructor and call signatures.  StyleLoss caches the
style-side operands (normalised bf16 rows, mean, covariance, YUV) once per target, which the
reference recomputes every iteration (nn/losses.py:43,49; run_strotss.py:37).  StrotssLoss is the
single-call evaluation of run_strotss.py:138-140 (one launch sequence for all four terms).
"""
from __future__ import annotations

import torch

from . import _lib
from .losses import self_similarity
from .runtime import Handle, reshape_2d


class ContentLoss(torch.nn.Module):
    def forward(self, target: torch.Tensor, prediction: torch.Tensor) -> torch.Tensor:
        # run_strotss.py:23-24: note the swapped argument order
        return self_similarity(prediction, target)


class _StyleLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, module):
        need = prediction.requires_grad
        scalars, grad = module.handle.style_loss(prediction.detach(), module.alpha, need)
        module.last_scalars = scalars
        ctx.save_for_backward(grad if need else None)
        return scalars[_lib.S_LOSS_S].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g if grad is not None else None), None


class StyleLoss(torch.nn.Module):
    """StyleLoss(target, alpha)(prediction) = l_m + l_remd + l_palette / max(alpha, 1)."""

    def __init__(self, target: torch.Tensor, alpha: float):
        super().__init__()
        self.target = reshape_2d(target)
        self.alpha = float(alpha)
        self.inv_alpha = 1 / max(alpha, 1)
        self.handle = Handle(self.target.device)
        self.handle.set_style_target(self.target)
        self.last_scalars = None

    def forward(self, prediction: torch.Tensor) -> torch.Tensor:
        return _StyleLossFn.apply(reshape_2d(prediction), self)


class _TotalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, content, module):
        need = prediction.requires_grad
        scalars, grad, _, _ = module.handle.eval(prediction.detach(), content.detach(), module.alpha, need)
        module.last_scalars = scalars
        ctx.save_for_backward(grad if need else None)
        return scalars[_lib.S_TOTAL].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g if grad is not None else None), None, None


class StrotssLoss(torch.nn.Module):
    """loss = (alpha * ContentLoss(content, pred) + StyleLoss(pred)) / (2 + alpha + 1/max(alpha, 1)).

    After a call, `.last_scalars` holds the device scalar block (indices in _lib.S_*): the three
    values the reference prints each iteration are TOTAL, LOSS_C, LOSS_S (run_strotss.py:150-152).
    """

    def __init__(self, target: torch.Tensor, alpha: float):
        super().__init__()
        self.alpha = float(alpha)
        self.style_features = reshape_2d(target).detach()
        self.handle = Handle(self.style_features.device)
        self.handle.set_style_target(self.style_features)
        self.last_scalars = None

    def forward(self, content: torch.Tensor, prediction: torch.Tensor) -> torch.Tensor:
        return _TotalFn.apply(reshape_2d(prediction), reshape_2d(content), self)
