"""Mirror of `Sampling` (nn/strotss_utils.py:20-136), SURVEY.md section 8(f) "next #1": the hypercolumn
gather that feeds the loss path, as one fused CUDA kernel (and its scatter-add backward) instead of
40 tf.gather calls + 9 concats per call.

Feature maps are NHWC float32 CUDA tensors of shape (1, h, w, c) like the reference's Keras outputs
(`[img] + vgg(img)`, run_strotss.py:95-96,135).  The index generator keeps the reference's logic (strided
grid with a random offset, mask filter with the empty-region fallback, pair-wise shuffle, first
`sample_size` points); its random draws come from a torch.Generator instead of tf_rng (nn/rand.py:21), so
individual draws differ from a TensorFlow run while the distribution is the same.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Tuple, Union

import torch

from .runtime import _check_features, _ptr, _stream, shared_handle


def _shapes(xs: List[torch.Tensor]):
    out = []
    for x in xs:
        if x.dim() == 4:
            if x.shape[0] != 1:
                raise ValueError("feature maps must have batch size 1 (the reference squeezes them, strotss_utils.py:39)")
            out.append((int(x.shape[1]), int(x.shape[2]), int(x.shape[3])))
        elif x.dim() == 3:
            out.append((int(x.shape[0]), int(x.shape[1]), int(x.shape[2])))
        else:
            raise ValueError(f"feature map of rank {x.dim()}; expected (1, h, w, c) or (h, w, c)")
    return out


def _call_sample(xs, indices, bilinear, out=None, grads=None, grad_out=None):
    dev = indices.device
    h = shared_handle(dev)
    shapes = _shapes(xs)
    n = int(indices.shape[0])
    k = len(xs)
    hs = (C.c_int * k)(*[s[0] for s in shapes])
    ws = (C.c_int * k)(*[s[1] for s in shapes])
    cs = (C.c_int * k)(*[s[2] for s in shapes])
    total = sum(s[2] for s in shapes)
    if grads is None:
        ptrs = (C.c_void_p * k)(*[x.data_ptr() for x in xs])
        out = torch.empty(n, total, device=dev, dtype=torch.float32)
        h._ck(h.lib.strotss_sample(h._h, k, ptrs, hs, ws, cs, _ptr(indices), n, 1 if bilinear else 0, _ptr(out), total, _stream(dev)),
              "strotss_sample")
        return out
    ptrs = (C.c_void_p * k)(*[(g.data_ptr() if g is not None else None) for g in grads])
    h._ck(h.lib.strotss_sample_backward(h._h, k, ptrs, hs, ws, cs, _ptr(indices), n, 1 if bilinear else 0, _ptr(grad_out), total,
                                        _stream(dev)), "strotss_sample_backward")
    return None


class _SampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, indices, bilinear, *xs):
        # which maps need a gradient is a property of the INPUTS: inside forward() autograd is off, so the contiguous copy
        # of a permuted (e.g. NCHW -> NHWC) map would report requires_grad = False and silently drop its gradient
        ctx.needs = list(ctx.needs_input_grad[2:])
        xs = [_check_features("feature map", x).contiguous() for x in xs]
        ctx.bilinear = bilinear
        ctx.shapes = [x.shape for x in xs]
        ctx.save_for_backward(indices)
        return _call_sample([x.detach() for x in xs], indices, bilinear)

    @staticmethod
    def backward(ctx, g):
        (indices,) = ctx.saved_tensors
        g = g.contiguous()
        grads = [torch.zeros(s, device=g.device, dtype=torch.float32) if need else None for s, need in zip(ctx.shapes, ctx.needs)]
        # _call_sample only needs the shapes of xs for the backward
        shapes_only = [torch.empty(s, device="meta") for s in ctx.shapes]
        _call_sample(shapes_only, indices, ctx.bilinear, grads=grads, grad_out=g)
        return (None, None) + tuple(grads)


class Sampling(torch.nn.Module):
    def __init__(self, sample_size: int, generator: Optional[torch.Generator] = None):
        super().__init__()
        self.sample_size = sample_size
        self.generator = generator            # CPU generator for the offsets / shuffle (seed 0 convention: nn/rand.py:12-21)

    def _sample(self, xs: List[torch.Tensor], indices: torch.Tensor, bilinear_sampling: bool) -> torch.Tensor:
        """nn/strotss_utils.py:25-81."""
        return _SampleFn.apply(indices, bool(bilinear_sampling), *xs)

    def _make_indices(self, base_tensor: torch.Tensor, bilinear_sampling: bool, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """nn/strotss_utils.py:83-121 (host-side index logic; tiny)."""
        h, w, _ = _shapes([base_tensor])[0]
        gen = self.generator
        if bilinear_sampling:
            area = math.sqrt((h * w) // (128 ** 2))
            step_x, step_y = max(1, math.floor(area)), max(1, math.ceil(area))
            off_x = int(torch.randint(0, step_x, (), generator=gen))
            off_y = int(torch.randint(0, step_y, (), generator=gen))
            X = torch.arange(h)[off_x::step_x]
            Y = torch.arange(w)[off_y::step_y]
        else:
            X, Y = torch.arange(h), torch.arange(w)
        XX, YY = torch.meshgrid(X, Y, indexing="xy")             # tf.meshgrid default indexing
        ret = torch.stack([XX.reshape(-1), YY.reshape(-1)], dim=1)
        if mask is not None:
            m = mask.detach().float().cpu()
            if m.dim() == 3:                                       # (H, W, 1) like load_mask's output
                m = m.permute(2, 0, 1)[None]
            elif m.dim() == 2:
                m = m[None, None]
            m = torch.nn.functional.interpolate(m, size=(h, w), mode="bilinear", align_corners=False, antialias=False)[0, 0]
            if float(m.max()) < 0.1:
                keep = (m + 1) > 0.5                               # empty-region fallback: everything (:107-108)
            else:
                keep = m > 0.5
            ret = ret[keep[ret[:, 0], ret[:, 1]]]
        perm = torch.randperm(ret.shape[0], generator=gen)         # pairs are shuffled together (:115-119)
        ret = ret[perm][: self.sample_size].to(torch.float32)
        return ret.to(base_tensor.device)

    def forward(self, xs: List[torch.Tensor], ys: Optional[List[torch.Tensor]] = None, mask: Optional[torch.Tensor] = None,
                bilinear_sampling: bool = False) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
        """nn/strotss_utils.py:123-134."""
        indices = self._make_indices(xs[0], bilinear_sampling, mask)
        ret = self._sample(xs, indices, bilinear_sampling)
        if ys:
            return ret, self._sample(ys, indices, bilinear_sampling)
        return ret

    def bilinear(self, xs, ys=None, mask=None):
        return self.forward(xs, ys, mask, bilinear_sampling=True)
