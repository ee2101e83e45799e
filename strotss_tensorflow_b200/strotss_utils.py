"""Mirror of the one hot-path function of nn/strotss_utils.py."""
from __future__ import annotations

import torch

from .runtime import shared_handle

# tf.image.rgb_to_yuv kernel; rows = R, G, B
_RGB_TO_YUV = [[0.299, -0.14714119, 0.61497538],
               [0.587, -0.28886916, -0.51496512],
               [0.114, 0.43601035, -0.10001026]]


class _RgbToYuv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        return shared_handle(x.device).convert_rgb_to_yuv(x.detach())

    @staticmethod
    def backward(ctx, g):
        k = torch.tensor(_RGB_TO_YUV, device=g.device, dtype=g.dtype)
        out = torch.zeros(ctx.shape, device=g.device, dtype=g.dtype)
        out[:, :3] = g @ k.T
        return out


def convert_rgb_to_yuv(x: torch.Tensor) -> torch.Tensor:
    """nn/strotss_utils.py:166-167: tf.image.rgb_to_yuv(x[:, :3])."""
    return _RgbToYuv.apply(x)
