"""Mirror of the one hot-path function of nn/strotss_utils.py."""
from __future__ import annotations

import weakref

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


# --------------------------------------------------------------------------------------------
# Pixel-side step (SURVEY 8f "next #3"): images are (1, h, w, c) float32 CUDA tensors like the reference's
# --------------------------------------------------------------------------------------------
import ctypes as _C
from typing import List

from .runtime import _check_features, _ptr, _stream


def _hwc(x: torch.Tensor):
    if x.dim() != 4 or x.shape[0] != 1:
        raise ValueError(f"expected a (1, h, w, c) image, got {tuple(x.shape)}")
    return int(x.shape[1]), int(x.shape[2]), int(x.shape[3])


def _resize(x: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    x = _check_features("image", x).contiguous()
    h, w, c = _hwc(x)
    out = torch.empty(1, oh, ow, c, device=x.device, dtype=torch.float32)
    hd = shared_handle(x.device)
    hd._ck(hd.lib.strotss_resize_bilinear(hd._h, _ptr(x), h, w, c, _ptr(out), oh, ow, _stream(x.device)), "strotss_resize_bilinear")
    return out


def resize(image: torch.Tensor, max_size) -> torch.Tensor:
    """nn/utils.py:32-37: long side -> max_size (tf.image.resize, bilinear)."""
    if max_size is None:
        return image
    h, w, _ = _hwc(image)
    factor = max(h / max_size, w / max_size)
    return _resize(image, int(h / factor), int(w / factor))


def resize_like(image: torch.Tensor, base: torch.Tensor) -> torch.Tensor:
    """nn/utils.py:40-41."""
    h, w, _ = _hwc(base)
    return _resize(image, h, w)


def make_laplacian(x: torch.Tensor, return_downscale: bool = False):
    """nn/strotss_utils.py:139-146."""
    x = _check_features("image", x).contiguous()
    h, w, c = _hwc(x)
    pyr = torch.empty_like(x)
    down = torch.empty(1, max(h // 2, 1), max(w // 2, 1), c, device=x.device, dtype=torch.float32)
    hd = shared_handle(x.device)
    hd._ck(hd.lib.strotss_make_laplacian(hd._h, _ptr(x), h, w, c, _ptr(pyr), _ptr(down), _stream(x.device)), "strotss_make_laplacian")
    return (pyr, down) if return_downscale else pyr


def make_laplacian_pyramid(x: torch.Tensor, levels: int = 5) -> List[torch.Tensor]:
    """nn/strotss_utils.py:149-156."""
    xs, cur = [], x
    for _ in range(levels):
        pyr, cur = make_laplacian(cur, return_downscale=True)
        xs.append(pyr)
    xs.append(cur)
    return xs


def _level_arrays(shapes):
    n = len(shapes)
    return (_C.c_int * n)(*[s[0] for s in shapes]), (_C.c_int * n)(*[s[1] for s in shapes])


class _Fold(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *xs):
        xs = [_check_features("pyramid level", x).contiguous() for x in xs]
        shapes = [_hwc(x) for x in xs]
        c = shapes[0][2]
        if any(s[2] != c for s in shapes):
            raise ValueError("all pyramid levels must have the same number of channels")
        n = len(xs)
        hs, ws = _level_arrays(shapes)
        out = torch.empty_like(xs[0])
        ptrs = (_C.c_void_p * n)(*[x.data_ptr() for x in xs])
        hd = shared_handle(out.device)
        hd._ck(hd.lib.strotss_pyramid_fold(hd._h, n, ptrs, hs, ws, c, _ptr(out), _stream(out.device)), "strotss_pyramid_fold")
        ctx.shapes = shapes
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        shapes = ctx.shapes
        n, c = len(shapes), shapes[0][2]
        hs, ws = _level_arrays(shapes)
        grads = [torch.empty(1, s[0], s[1], c, device=g.device, dtype=torch.float32) for s in shapes]
        ptrs = (_C.c_void_p * n)(*[t.data_ptr() for t in grads])
        hd = shared_handle(g.device)
        hd._ck(hd.lib.strotss_pyramid_fold_backward(hd._h, n, hs, ws, c, _ptr(g), ptrs, _stream(g.device)),
               "strotss_pyramid_fold_backward")
        return tuple(grads)


def fold_laplacian_pyramid(xs: List[torch.Tensor]) -> torch.Tensor:
    """nn/strotss_utils.py:159-163; differentiable w.r.t. every level (the optimisation variables, run_strotss.py:89)."""
    return _Fold.apply(*xs)


class RMSprop:
    """tf.keras.optimizers.RMSprop(rho, epsilon, learning_rate) as the reference uses it (run_strotss.py:63,85,88,148):
    momentum 0, not centred, one slot per variable created at first use, `lr` settable between scales.  All variables
    are updated by ONE kernel launch."""

    def __init__(self, rho: float = 0.9, epsilon: float = 1e-7, learning_rate: float = 1e-3):
        self.rho, self.epsilon, self.lr = float(rho), float(epsilon), float(learning_rate)
        # id(variable) -> (weak reference to the variable, rms slot).  The reference reuses ONE optimizer across all scales
        # with fresh variables per scale (run_strotss.py:63,89): a freed variable's id can be handed to a new, differently
        # shaped one, so a slot is only reused while its weak reference still points at the very same tensor.
        self._slots = {}

    def _slot(self, v: torch.Tensor) -> torch.Tensor:
        entry = self._slots.get(id(v))
        if entry is not None:
            ref, slot = entry
            if ref() is v and slot.shape == v.shape and slot.device == v.device:
                return slot
        for key in [k for k, (ref, _) in self._slots.items() if ref() is None]:
            del self._slots[key]                      # variables of earlier scales that were collected
        slot = torch.zeros_like(v)
        self._slots[id(v)] = (weakref.ref(v), slot)
        return slot

    def slots(self):
        """The live rms slots (e.g. to zero them when a captured graph is reused for a new image)."""
        return [slot for ref, slot in self._slots.values() if ref() is not None]

    def build(self, var_list):
        """Create the rms slots now (Keras creates them at the first apply_gradients).  Call this before capturing
        apply_gradients into a CUDA graph: a slot created inside the capture would be re-zeroed by every replay."""
        for v in var_list:
            self._slot(v)

    def apply_gradients(self, grads_and_vars):
        pairs = [(g, v) for g, v in grads_and_vars if g is not None]
        if not pairs:
            return
        n = len(pairs)
        slots = []
        for g, v in pairs:
            _check_features("variable", v); _check_features("gradient", g)
            if not v.is_contiguous() or g.shape != v.shape:
                raise ValueError("variables must be contiguous and gradients shaped like them")
            slots.append(self._slot(v))
        gs = [g.contiguous() for g, _ in pairs]
        dev = pairs[0][1].device
        vp = (_C.c_void_p * n)(*[v.data_ptr() for _, v in pairs])
        sp = (_C.c_void_p * n)(*[s.data_ptr() for s in slots])
        gp = (_C.c_void_p * n)(*[g.data_ptr() for g in gs])
        cnt = (_C.c_longlong * n)(*[v.numel() for _, v in pairs])
        hd = shared_handle(dev)
        hd._ck(hd.lib.strotss_rmsprop_step(hd._h, n, vp, sp, gp, cnt, self.lr, self.rho, self.epsilon, _stream(dev)),
               "strotss_rmsprop_step")
