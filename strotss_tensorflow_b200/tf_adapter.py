"""tf.custom_gradient adapter for the reference's TensorFlow driver (the binding a maintainer of the reference drops
next to nn/losses.py).

TensorFlow is not installed in the build image or on the GPU boxes, so this module has never run
against real TensorFlow.  What does run, on the B200 (tests/test_gpu_tf_adapter.py): every line of this file against
`tests/tf_standin.py`, a stand-in `tensorflow` module on CUDA tensors that provides exactly the API used here
(tf.custom_gradient, tf.py_function, tf.experimental.dlpack.{to,from}_dlpack, tf.test.experimental.sync_devices,
tf.GradientTape, ...) with TensorFlow's documented semantics -- the control flow, the DLPack pointer hand-over, the C-ABI
calls and the gradients are therefore exercised; TensorFlow's own behaviour behind those API names is not.

Usage inside run_strotss.py (replacing `from nn.losses import ...`):

    from strotss_tensorflow_b200.tf_adapter import relaxed_emd, moment_matching, self_similarity, StyleLoss, ContentLoss

Mechanics: the forward is a tf.py_function (train_step is a @tf.function graph, run_strotss.py:104,131; N is dynamic in
masked mode, nn/strotss_utils.py:113).  Inputs are GPU EagerTensors handed over zero-copy through DLPack (read-only use).
Outputs are NOT written into TensorFlow tensors (those are immutable): the library allocates them
(strotss_device_alloc), the kernels write there, and the buffer enters TensorFlow through
tf.experimental.dlpack.from_dlpack with a deleter that returns it (strotss_device_free).  TensorFlow does not expose its
compute stream, so the adapter runs on the legacy default stream and synchronises the device before the call (the
producers of the inputs have finished) and after it (the outputs are complete before TensorFlow reads them); the reference
already synchronises three scalars per iteration for tqdm (run_strotss.py:150-152).
"""
from __future__ import annotations

import ctypes as C
import re

from . import _lib

try:
    import tensorflow as tf
    _HAVE_TF = True
except Exception:                        # ModuleNotFoundError in this image
    tf = None
    _HAVE_TF = False


def _require_tf():
    if not _HAVE_TF:
        raise RuntimeError("tf_adapter needs TensorFlow >= 2.6 (README.md:11 of the reference); it is not installed here. "
                           "Use the torch adapter (strotss_tensorflow_b200.losses) instead.")


def _device_index(t) -> int:
    """'/job:localhost/replica:0/task:0/device:GPU:1' -> 1 (the GPU the tensor lives on; --gpu_id, nn/utils.py:73-85)."""
    m = re.search(r"GPU:(\d+)", str(getattr(t, "device", "")))
    if not m:
        raise RuntimeError(f"tensor is on {getattr(t, 'device', '?')!r}: the STROTSS loss path runs on a GPU only (no CPU fallback)")
    return int(m.group(1))


class _Handles:
    """One library handle per GPU for the stateless functions."""
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

    @classmethod
    def close(cls):
        for h in cls.by_device.values():
            cls.lib.strotss_destroy(h)
        cls.by_device = {}


# ---- DLPack -------------------------------------------------------------------------------------------------------------
class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    pass


_DELETER = C.CFUNCTYPE(None, C.c_void_p)
_DLManagedTensor._fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", _DELETER)]
_KDL_CUDA, _KDL_FLOAT = 2, 2
# address of a DLManagedTensor -> everything that must outlive the capsule.  The table and the deleter thunk below are
# referenced by raw pointers inside tensors that TensorFlow may free at any later time, so they must survive a reload of
# this module (importlib.reload re-executes it in the same namespace): they are created once per process.
_live = globals().get("_live", {})

C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
C.pythonapi.PyCapsule_New.restype = C.py_object
C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]


def _dev_ptr(t):
    """GPU EagerTensor -> (raw device pointer, keep-alive capsule) via DLPack; the tensor is only read."""
    cap = tf.experimental.dlpack.to_dlpack(t)
    managed = C.cast(C.pythonapi.PyCapsule_GetPointer(cap, b"dltensor"), C.POINTER(_DLManagedTensor))
    dl = managed.contents.dl_tensor
    return C.c_void_p((dl.data or 0) + dl.byte_offset), cap


def _free_output_impl(managed_addr, _table=_live):
    rec = _table.pop(managed_addr, None)
    if rec is not None:
        rec[0].strotss_device_free(None, rec[2])      # the handle may be gone already: the library ignores it here


_free_output = globals().get("_free_output") or _DELETER(_free_output_impl)


class _Output:
    """fp32 device buffer of `shape`, owned by the library until TensorFlow's DLPack deleter returns it."""

    def __init__(self, lib, handle, device_index: int, shape):
        self.lib, self.handle, self.shape = lib, handle, tuple(int(s) for s in shape)
        n = 1
        for s in self.shape:
            n *= s
        self.ptr = C.c_void_p()
        _lib.check(lib, handle, lib.strotss_device_alloc(handle, max(n, 1) * 4, C.byref(self.ptr)), "strotss_device_alloc")
        self.device_index = device_index

    def to_tf(self):
        """Zero-copy TensorFlow tensor over the buffer (ownership moves to TensorFlow's DLPack consumer)."""
        shape = (C.c_int64 * max(len(self.shape), 1))(*self.shape)
        m = _DLManagedTensor()
        m.dl_tensor.data = self.ptr.value
        m.dl_tensor.device = _DLDevice(_KDL_CUDA, self.device_index)
        m.dl_tensor.ndim = len(self.shape)
        m.dl_tensor.dtype = _DLDataType(_KDL_FLOAT, 32, 1)
        m.dl_tensor.shape = C.cast(shape, C.POINTER(C.c_int64))
        m.dl_tensor.strides = None
        m.dl_tensor.byte_offset = 0
        m.manager_ctx = None
        m.deleter = _free_output
        _live[C.addressof(m)] = (self.lib, self.handle, self.ptr, m, shape)
        cap = C.pythonapi.PyCapsule_New(C.addressof(m), b"dltensor", None)
        return tf.experimental.dlpack.from_dlpack(cap)


def _reshape_2d(x):                      # nn/losses.py:31-36
    x = tf.squeeze(x)
    return tf.reshape(x, (-1, tf.shape(x)[-1]))


def _sync():
    tf.test.experimental.sync_devices()


# ---- the three loss functions of nn/losses.py ------------------------------------------------------------------------------
def _call_self_similarity(x, y):
    dev = _device_index(x)
    lib, h = _lib.load(), _Handles.get(dev)
    n, d = int(x.shape[0]), int(x.shape[1])
    loss, grad = _Output(lib, h, dev, (1,)), _Output(lib, h, dev, (n, d))
    (px, kx), (py, ky) = _dev_ptr(x), _dev_ptr(y)
    _sync()
    code = lib.strotss_self_similarity(h, px, d, py, d, n, d, loss.ptr, grad.ptr, d, None)
    _sync()
    loss_t, grad_t = loss.to_tf(), grad.to_tf()
    _lib.check(lib, h, code, "strotss_self_similarity")
    del kx, ky
    return loss_t[0], grad_t


def self_similarity(x, y):
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


def _call_pair(fn_name, x, y, distance_code=None):
    dev = _device_index(y)
    lib, h = _lib.load(), _Handles.get(dev)
    m, d = int(x.shape[0]), int(x.shape[1])
    n = int(y.shape[0])
    loss, grad = _Output(lib, h, dev, (4,)), _Output(lib, h, dev, (n, d))
    (px, kx), (py, ky) = _dev_ptr(x), _dev_ptr(y)
    _sync()
    if fn_name == "relaxed_emd":
        code = lib.strotss_relaxed_emd(h, px, d, m, py, d, n, d, int(distance_code), loss.ptr, grad.ptr, d, None, None, None)
    else:
        code = lib.strotss_moment_matching(h, px, d, m, py, d, n, d, loss.ptr, grad.ptr, d, None)
    _sync()
    loss_t, grad_t = loss.to_tf(), grad.to_tf()
    _lib.check(lib, h, code, "strotss_" + fn_name)
    del kx, ky
    return loss_t[0], grad_t


def relaxed_emd(x, y, distance: str = "cosine"):
    """nn/losses.py:69-80 (gradient w.r.t. y, the prediction: run_strotss.py:36,39)."""
    _require_tf()
    if distance not in _lib.DIST_CODES:
        raise KeyError(distance)          # dist_metrics[distance], nn/losses.py:74
    code = _lib.DIST_CODES[distance]

    @tf.custom_gradient
    def op(x2, y2):
        loss, grad = tf.py_function(lambda a, b: _call_pair("relaxed_emd", a, b, code), [x2, y2], [tf.float32, tf.float32])
        loss.set_shape([])
        grad.set_shape(y2.shape)
        return loss, (lambda upstream: (None, upstream * grad))

    return op(_reshape_2d(x), _reshape_2d(y))


def moment_matching(x, y):
    """nn/losses.py:39-52 (gradient w.r.t. y)."""
    _require_tf()

    @tf.custom_gradient
    def op(x2, y2):
        loss, grad = tf.py_function(lambda a, b: _call_pair("moment_matching", a, b), [x2, y2], [tf.float32, tf.float32])
        loss.set_shape([])
        grad.set_shape(y2.shape)
        return loss, (lambda upstream: (None, upstream * grad))

    return op(_reshape_2d(x), _reshape_2d(y))


def convert_rgb_to_yuv(x):               # nn/strotss_utils.py:166-167; K = 3, stays a TensorFlow op
    _require_tf()
    return tf.image.rgb_to_yuv(x[:, :3])


class ContentLoss:                       # run_strotss.py:21-24
    def __call__(self, target, prediction):
        return self_similarity(prediction, target)


class StyleLoss:                         # run_strotss.py:27-40
    def __init__(self, target, alpha: float):
        self.target = target
        self.inv_alpha = 1 / max(alpha, 1)

    def __call__(self, prediction):
        l_m = moment_matching(self.target, prediction)
        l_remd = relaxed_emd(self.target, prediction)
        l_palette = relaxed_emd(convert_rgb_to_yuv(self.target), convert_rgb_to_yuv(prediction), distance="both")
        return l_m + l_remd + (self.inv_alpha * l_palette)


class StrotssLoss:
    """Fused evaluation for the reference's train_step (run_strotss.py:136-140): one strotss_eval call per iteration
    instead of four ops, with the style-side statistics cached per scale (strotss_set_style_target).

        loss_fn = StrotssLoss(sampling(style_feat), alpha)        # replaces StyleLoss(...) at run_strotss.py:128
        loss, loss_c, loss_s = loss_fn(c_feat, p_feat)            # replaces :138-140

    Masked mode (:97-125): build one StrotssLoss per region exactly as the reference builds one StyleLoss per region, or
    bind strotss_set_style_targets_grouped / strotss_eval_grouped the same way (see modules.MaskedStrotssLoss)."""

    def __init__(self, target, alpha: float):
        _require_tf()
        self.alpha = float(alpha)
        self.lib = _lib.load()
        t = _reshape_2d(target)
        self.device_index = _device_index(t)          # the handle lives on the GPU of the style features
        self.h = C.c_void_p()
        _lib.check(self.lib, self.h, self.lib.strotss_create(self.device_index, C.byref(self.h)), "strotss_create")
        m, d = int(t.shape[0]), int(t.shape[1])
        p, keep = _dev_ptr(t)
        _sync()
        code = self.lib.strotss_set_style_target(self.h, p, m, d, d, None)
        _sync()
        _lib.check(self.lib, self.h, code, "strotss_set_style_target")
        del keep

    def close(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            _sync()
            self.lib.strotss_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, content, pred):
        n, d = int(pred.shape[0]), int(pred.shape[1])
        scalars = _Output(self.lib, self.h, self.device_index, (_lib.NUM_SCALARS,))
        grad = _Output(self.lib, self.h, self.device_index, (n, d))
        (pp, k1), (pc, k2) = _dev_ptr(pred), _dev_ptr(content)
        _sync()
        code = self.lib.strotss_eval(self.h, pp, d, pc, d, n, self.alpha, scalars.ptr, grad.ptr, d, None, None, None)
        _sync()
        s, g = scalars.to_tf(), grad.to_tf()
        _lib.check(self.lib, self.h, code, "strotss_eval")
        del k1, k2
        return s[_lib.S_TOTAL], s[_lib.S_LOSS_C], s[_lib.S_LOSS_S], g

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
