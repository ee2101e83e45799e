"""tf.custom_gradient adapter for the reference's TensorFlow driver.  UNEXECUTED IN THIS ENVIRONMENT:
TensorFlow is not installed in the build image or on the GPU boxes, so this
file is the binding a maintainer of the reference would drop next to nn/losses.py; the adapter that
is actually exercised by the tests is the torch.autograd.Function one (losses.py / modules.py), which
calls the same C ABI with the same argument meaning.

Usage inside run_strotss.py (replacing `from nn.losses import ...`):

    from strotss_tensorflow_b200.tf_adapter import relaxed_emd, moment_matching, self_similarity, StyleLoss, ContentLoss

Mechanics: the forward is a tf.py_function (train_step is a @tf.function graph, run_strotss.py:104,131;
N is dynamic in masked mode, nn/strotss_utils.py:113).  GPU EagerTensors are handed over zero-copy through
DLPack; the raw device pointers go to the C ABI on TensorFlow's compute stream is not exposed, so the
adapter synchronises the device before and after the call (one sync per evaluation; the reference
already syncs three scalars per iteration for tqdm, run_strotss.py:150-152).
"""
from __future__ import annotations

import ctypes as C

from . import _lib

try:                                    # pragma: no cover - TensorFlow is absent in this image
    import tensorflow as tf
    _HAVE_TF = True
except Exception:                        # ModuleNotFoundError here
    tf = None
    _HAVE_TF = False


def _require_tf():
    if not _HAVE_TF:
        raise RuntimeError("tf_adapter needs TensorFlow >= 2.6 (README.md:11 of the reference); it is not installed here. "
                           "Use the torch adapter (strotss_tensorflow_b200.losses) instead.")


class _Handles:                          # pragma: no cover
    lib = None
    by_device = {}

    @classmethod
    def get(cls, device_index: int):
        if cls.lib is None:
            cls.lib = _lib.load()
        h = cls.by_device.get(device_index)
        if h is None:
            h = C.c_void_p()
            _lib.check(cls.lib, h, cls.lib.strotss_create(device_index, C.byref(h)), "strotss_create")
            cls.by_device[device_index] = h
        return h


def _dev_ptr(t):                         # pragma: no cover
    """EagerTensor on GPU -> (raw device pointer, keep-alive capsule) via DLPack."""
    cap = tf.experimental.dlpack.to_dlpack(t)
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    managed = C.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")
    data = C.cast(managed, C.POINTER(C.c_void_p))[0]         # DLTensor.data is the first field
    return C.c_void_p(data), cap


def _reshape_2d(x):                      # pragma: no cover  (nn/losses.py:31-36)
    x = tf.squeeze(x)
    return tf.reshape(x, (-1, tf.shape(x)[-1]))


def _call_self_similarity(x, y):         # pragma: no cover
    lib = _lib.load()
    h = _Handles.get(0)
    n, d = int(x.shape[0]), int(x.shape[1])
    loss = tf.zeros([1], tf.float32)
    grad = tf.zeros_like(x)
    (px, kx), (py, ky), (pl, kl), (pg, kg) = _dev_ptr(x), _dev_ptr(y), _dev_ptr(loss), _dev_ptr(grad)
    tf.test.experimental.sync_devices()
    _lib.check(lib, h, lib.strotss_self_similarity(h, px, d, py, d, n, d, pl, pg, d, None), "strotss_self_similarity")
    tf.test.experimental.sync_devices()
    return loss[0], grad


def self_similarity(x, y):               # pragma: no cover
    """nn/losses.py:55-66 (gradient w.r.t. x, the prediction: run_strotss.py:24)."""
    _require_tf()

    @tf.custom_gradient
    def op(x2, y2):
        loss, grad = tf.py_function(_call_self_similarity, [x2, y2], [tf.float32, tf.float32])
        loss.set_shape([])
        grad.set_shape(x2.shape)

        def bwd(upstream):
            return upstream * grad, None
        return loss, bwd

    return op(_reshape_2d(x), _reshape_2d(y))


def _call_pair(fn_name, x, y, distance_code=None):   # pragma: no cover
    lib = _lib.load()
    h = _Handles.get(0)
    m, d = int(x.shape[0]), int(x.shape[1])
    n = int(y.shape[0])
    loss = tf.zeros([4], tf.float32)
    grad = tf.zeros_like(y)
    (px, kx), (py, ky), (pl, kl), (pg, kg) = _dev_ptr(x), _dev_ptr(y), _dev_ptr(loss), _dev_ptr(grad)
    tf.test.experimental.sync_devices()
    if fn_name == "relaxed_emd":
        code = lib.strotss_relaxed_emd(h, px, d, m, py, d, n, d, int(distance_code), pl, pg, d, None, None, None)
    else:
        code = lib.strotss_moment_matching(h, px, d, m, py, d, n, d, pl, pg, d, None)
    _lib.check(lib, h, code, "strotss_" + fn_name)
    tf.test.experimental.sync_devices()
    return loss[0], grad


def relaxed_emd(x, y, distance: str = "cosine"):     # pragma: no cover
    """nn/losses.py:69-80 (gradient w.r.t. y, the prediction: run_strotss.py:36,39)."""
    _require_tf()
    if distance not in _lib.DIST_CODES:
        raise KeyError(distance)
    code = _lib.DIST_CODES[distance]

    @tf.custom_gradient
    def op(x2, y2):
        loss, grad = tf.py_function(lambda a, b: _call_pair("relaxed_emd", a, b, code), [x2, y2], [tf.float32, tf.float32])
        loss.set_shape([])
        grad.set_shape(y2.shape)
        return loss, (lambda upstream: (None, upstream * grad))

    return op(_reshape_2d(x), _reshape_2d(y))


def moment_matching(x, y):               # pragma: no cover
    """nn/losses.py:39-52 (gradient w.r.t. y)."""
    _require_tf()

    @tf.custom_gradient
    def op(x2, y2):
        loss, grad = tf.py_function(lambda a, b: _call_pair("moment_matching", a, b), [x2, y2], [tf.float32, tf.float32])
        loss.set_shape([])
        grad.set_shape(y2.shape)
        return loss, (lambda upstream: (None, upstream * grad))

    return op(_reshape_2d(x), _reshape_2d(y))


def convert_rgb_to_yuv(x):               # pragma: no cover  (nn/strotss_utils.py:166-167; K=3, stays a TF op)
    _require_tf()
    return tf.image.rgb_to_yuv(x[:, :3])


class ContentLoss:                       # pragma: no cover  (run_strotss.py:21-24)
    def __call__(self, target, prediction):
        return self_similarity(prediction, target)


class StyleLoss:                         # pragma: no cover  (run_strotss.py:27-40)
    def __init__(self, target, alpha: float):
        self.target = target
        self.inv_alpha = 1 / max(alpha, 1)

    def __call__(self, prediction):
        l_m = moment_matching(self.target, prediction)
        l_remd = relaxed_emd(self.target, prediction)
        l_palette = relaxed_emd(convert_rgb_to_yuv(self.target), convert_rgb_to_yuv(prediction), distance="both")
        return l_m + l_remd + (self.inv_alpha * l_palette)


class StrotssLoss:                       # pragma: no cover
    """Fused evaluation for the reference's train_step (run_strotss.py:136-140): one strotss_eval call per iteration
    instead of four ops, with the style-side statistics cached per scale (strotss_set_style_target).

        loss_fn = StrotssLoss(sampling(style_feat), alpha)        # replaces StyleLoss(...) at run_strotss.py:128
        loss, loss_c, loss_s = loss_fn(c_feat, p_feat)            # replaces :138-140

    Masked mode (:97-125): build one StrotssLoss per region exactly as the reference builds one StyleLoss per region, or
    bind strotss_set_style_targets_grouped / strotss_eval_grouped the same way (see modules.MaskedStrotssLoss)."""

    def __init__(self, target, alpha: float, device_index: int = 0):
        _require_tf()
        self.alpha = float(alpha)
        self.lib = _lib.load()
        self.h = C.c_void_p()
        _lib.check(self.lib, self.h, self.lib.strotss_create(device_index, C.byref(self.h)), "strotss_create")
        t = _reshape_2d(target)
        m, d = int(t.shape[0]), int(t.shape[1])
        p, keep = _dev_ptr(t)
        tf.test.experimental.sync_devices()
        _lib.check(self.lib, self.h, self.lib.strotss_set_style_target(self.h, p, m, d, d, None), "strotss_set_style_target")
        tf.test.experimental.sync_devices()

    def _call(self, content, pred):
        n, d = int(pred.shape[0]), int(pred.shape[1])
        scalars = tf.zeros([_lib.NUM_SCALARS], tf.float32)
        grad = tf.zeros_like(pred)
        (pp, k1), (pc, k2), (ps, k3), (pg, k4) = _dev_ptr(pred), _dev_ptr(content), _dev_ptr(scalars), _dev_ptr(grad)
        tf.test.experimental.sync_devices()
        _lib.check(self.lib, self.h, self.lib.strotss_eval(self.h, pp, d, pc, d, n, self.alpha, ps, pg, d, None, None, None),
                   "strotss_eval")
        tf.test.experimental.sync_devices()
        return scalars[_lib.S_TOTAL], scalars[_lib.S_LOSS_C], scalars[_lib.S_LOSS_S], grad

    def __call__(self, content, prediction):
        @tf.custom_gradient
        def op(c2, p2):
            loss, loss_c, loss_s, grad = tf.py_function(self._call, [c2, p2], [tf.float32] * 4)
            for t in (loss, loss_c, loss_s):
                t.set_shape([])
            grad.set_shape(p2.shape)

            def bwd(g_loss, g_c, g_s):          # only `loss` is differentiated by the driver (run_strotss.py:141)
                return None, g_loss * grad
            return (loss, tf.stop_gradient(loss_c), tf.stop_gradient(loss_s)), bwd

        return op(_reshape_2d(content), _reshape_2d(prediction))
