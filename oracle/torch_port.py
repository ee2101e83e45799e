"""torch-CPU port of the reference op sequence with framework autodiff.  TEST INFRASTRUCTURE ONLY.

Parity status as for strotss_oracle.py: held to the golden vectors of the reference's own loss code
(tests/golden/ref_*.npz, <= 1e-10), TensorFlow's op semantics unpinned.  This file exists for two purposes:
  1. cross-checking the hand-written gradients of strotss_oracle.py against an independent
     autodiff on inputs without ties (torch and TF differ on ties: torch.min(dim) picks one
     index and torch.maximum splits 0.5/0.5 -- the NumPy oracle encodes TF's rules);
  2. the CPU baseline timed by bench.py (`cpu_baseline`, `--impl reference`): it materialises
     every N x M / N x N / D x D matrix and lets autograd run the dense backward GEMMs, i.e. it
     does the work the reference does on CPU TensorFlow (nn/losses.py:8-80,
     nn/strotss_utils.py:166-167, run_strotss.py:21-40,140), fp32.
It is never imported by the product package.
"""
from __future__ import annotations

import torch

_RGB_TO_YUV = torch.tensor(
    [[0.299, -0.14714119, 0.61497538],
     [0.587, -0.28886916, -0.51496512],
     [0.114, 0.43601035, -0.10001026]], dtype=torch.float64)


def l2_normalize(x, eps=1e-12):
    ss = (x * x).sum(dim=1, keepdim=True)
    return x * torch.rsqrt(torch.clamp_min(ss, eps))


def cosine_distance(x, y):                      # nn/losses.py:12-15
    return 1 - l2_normalize(x) @ l2_normalize(y).T


def l2_distance(x, y):                          # nn/losses.py:18-24
    x_sq = (x ** 2).sum(dim=1).reshape(-1, 1)
    y_sq = (y ** 2).sum(dim=1).reshape(1, -1)
    m = x_sq + y_sq - 2.0 * (x @ y.T)
    m = torch.clamp_min(m, 1e-06) / x.shape[1]
    return torch.sqrt(m)


dist_metrics = {'cosine': cosine_distance, 'l2': l2_distance,
                'both': lambda x, y: cosine_distance(x, y) + l2_distance(x, y)}


def mae(x, y):                                  # nn/losses.py:8-9
    return (x - y).abs().mean()


def moment_matching(x, y):                      # nn/losses.py:39-52
    xm = x.mean(dim=0, keepdim=True)
    ym = y.mean(dim=0, keepdim=True)
    cx = x - xm
    cy = y - ym
    xv = cx.T @ cx / x.shape[0]
    yv = cy.T @ cy / y.shape[0]
    return mae(xv, yv) + mae(xm, ym)


def self_similarity(x, y):                      # nn/losses.py:55-66
    xd = cosine_distance(x, x)
    xd = xd / torch.clamp_min(xd.sum(dim=0), 1e-12)
    yd = cosine_distance(y, y)
    yd = yd / torch.clamp_min(yd.sum(dim=0), 1e-12)
    return mae(xd, yd) * y.shape[0]


def relaxed_emd(x, y, distance='cosine'):       # nn/losses.py:69-80
    C = dist_metrics[distance](x, y)
    R_X = C.min(dim=1).values.mean()
    R_Y = C.min(dim=0).values.mean()
    return torch.maximum(R_X, R_Y)


def convert_rgb_to_yuv(x):                      # nn/strotss_utils.py:166-167
    return x[:, :3] @ _RGB_TO_YUV.to(device=x.device, dtype=x.dtype)


def style_loss(target, prediction, alpha):      # run_strotss.py:27-40
    inv_alpha = 1.0 / max(alpha, 1.0)
    l_m = moment_matching(target, prediction)
    l_remd = relaxed_emd(target, prediction)
    l_palette = relaxed_emd(convert_rgb_to_yuv(target), convert_rgb_to_yuv(prediction), 'both')
    return l_m + l_remd + inv_alpha * l_palette, (l_m, l_remd, l_palette)


def total_loss_and_grad(style, content, pred, alpha=16.0):
    """One evaluation as the reference's train_step does it (run_strotss.py:131-142)."""
    pred = pred.detach().requires_grad_(True)
    loss_c = self_similarity(pred, content)     # ContentLoss swaps the arguments (:24)
    loss_s, parts = style_loss(style, pred, alpha)
    denom = 2.0 + alpha + 1.0 / max(alpha, 1.0)
    loss = (alpha * loss_c + loss_s) / denom
    (grad,) = torch.autograd.grad(loss, pred)
    return loss.detach(), grad, dict(loss_c=loss_c.detach(), loss_s=loss_s.detach(),
                                     l_m=parts[0].detach(), l_remd=parts[1].detach(), l_palette=parts[2].detach())
