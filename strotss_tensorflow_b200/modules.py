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
from .runtime import acquire_handle, release_handle, reshape_2d


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
        self.handle = acquire_handle(self.target.device)
        self.handle.set_style_target(self.target)
        self.last_scalars = None

    def __del__(self):
        release_handle(getattr(self, "handle", None))

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
        self.handle = acquire_handle(self.style_features.device)
        self.handle.set_style_target(self.style_features)
        self.last_scalars = None

    def __del__(self):
        release_handle(getattr(self, "handle", None))

    def forward(self, content: torch.Tensor, prediction: torch.Tensor) -> torch.Tensor:
        return _TotalFn.apply(reshape_2d(prediction), reshape_2d(content), self)


class _MaskedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, R, *feats):
        contents, preds = feats[:R], feats[R:]
        need = any(p.requires_grad for p in preds)
        scalars, region, grads = module.handle.eval_grouped([p.detach() for p in preds], [c.detach() for c in contents],
                                                            module.alpha, need)
        module.last_scalars, module.last_region_scalars = scalars, region
        ctx.R = R
        ctx.save_for_backward(*(grads if need else []))
        return scalars[_lib.S_TOTAL].clone()

    @staticmethod
    def backward(ctx, g):
        grads = ctx.saved_tensors
        out = tuple(gr * g for gr in grads) if grads else (None,) * ctx.R
        return (None, None) + (None,) * ctx.R + out


class MaskedStrotssLoss(torch.nn.Module):
    """The masked train_step's loss (run_strotss.py:97-125): one StyleLoss target per region, per-region
    content/prediction samples, `loss = mean_r (alpha * loss_c_r + loss_s_r) / loss_denom`.

    forward(contents, predictions) takes two lists of (N_r, D) hypercolumn matrices (what
    `sampling.bilinear(content_feat, pred, mask=content_masks[r])` returns for each region).  After a call
    `.last_scalars[TOTAL|LOSS_C|LOSS_S]` are the three values train_step reports (:123-125) and
    `.last_region_scalars` the per-region blocks.
    """

    def __init__(self, targets, alpha: float):
        super().__init__()
        self.alpha = float(alpha)
        self.targets = [reshape_2d(t).detach() for t in targets]
        self.handle = acquire_handle(self.targets[0].device)
        self.handle.set_style_targets_grouped(self.targets)
        self.last_scalars = None
        self.last_region_scalars = None

    def __del__(self):
        release_handle(getattr(self, "handle", None))

    def forward(self, contents, predictions) -> torch.Tensor:
        R = len(self.targets)
        if len(contents) != R or len(predictions) != R:
            raise ValueError(f"expected {R} regions")
        feats = [reshape_2d(c) for c in contents] + [reshape_2d(p) for p in predictions]
        return _MaskedFn.apply(self, R, *feats)
