"""A stand-in `tensorflow` module on torch tensors (CPU or CUDA) -- TEST INFRASTRUCTURE.

TensorFlow is not installable in this image, so `strotss_tensorflow_b200/tf_adapter.py` (the tf.custom_gradient binding of
the C ABI) is executed against this module instead: it provides exactly the API names the adapter and a train_step
(run_strotss.py:131-142) touch, with TensorFlow's documented semantics --

    tf.custom_gradient            f(*x) -> (y, grad_fn); grad_fn(*dy) -> dx (None for non-differentiable inputs)
    tf.py_function                eager call of a Python function on eager tensors, outputs cast to Tout
    tf.experimental.dlpack        to_dlpack / from_dlpack capsules ("dltensor"; the consumer calls the producer's deleter)
    tf.test.experimental.sync_devices, tf.GradientTape, tf.stop_gradient, tf.squeeze / reshape / shape, tf.image.rgb_to_yuv

over torch (autograd.Function, torch.utils.dlpack, torch.cuda.synchronize).  What this executes for real: the adapter's
control flow, its DLPack structs and pointer arithmetic, the C-ABI calls on device memory, and the gradient plumbing.
What it cannot show: that real TensorFlow behaves like its documentation behind the same names.
"""
from __future__ import annotations

import types

import torch
import torch.utils.dlpack

_YUV_KERNEL = [[0.299, -0.14714119, 0.61497538],
               [0.587, -0.28886916, -0.51496512],
               [0.114, 0.43601035, -0.10001026]]


class _Shape(tuple):
    def as_list(self):
        return list(self)

    @property
    def rank(self):
        return len(self)


def _raw(v):
    return v.t if isinstance(v, T) else v


class T:
    """Stand-in for an EagerTensor."""

    def __init__(self, t):
        self.t = t

    @property
    def shape(self):
        return _Shape(self.t.shape)

    @property
    def dtype(self):
        return self.t.dtype

    @property
    def device(self):
        d = self.t.device
        return f"/job:localhost/replica:0/task:0/device:{'GPU' if d.type == 'cuda' else 'CPU'}:{d.index or 0}"

    def set_shape(self, shape):
        want = tuple(shape)
        have = tuple(self.t.shape)
        assert len(want) == len(have) and all(w is None or int(w) == h for w, h in zip(want, have)), (want, have)

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(_raw(i) for i in idx)
        return T(self.t[_raw(idx)])

    def __neg__(self):
        return T(-self.t)

    def __float__(self):
        return float(self.t)

    def __int__(self):
        return int(self.t)


def _binary(name, op):
    setattr(T, f"__{name}__", lambda a, b: T(op(_raw(a), _raw(b))))
    setattr(T, f"__r{name}__", lambda a, b: T(op(_raw(b), _raw(a))))


_binary("add", lambda a, b: a + b)
_binary("sub", lambda a, b: a - b)
_binary("mul", lambda a, b: a * b)
_binary("truediv", lambda a, b: a / b)


def make_module() -> types.ModuleType:
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "0.0-standin"
    tf.Tensor = T
    tf.float32, tf.int32 = torch.float32, torch.int32
    tf.constant = lambda v, dtype=torch.float32, device=None: T(torch.as_tensor(v, dtype=dtype, device=device))
    tf.convert_to_tensor = lambda v, dtype=None: v if isinstance(v, T) else T(torch.as_tensor(v, dtype=dtype))
    tf.zeros = lambda shape, dtype=torch.float32: T(torch.zeros(tuple(int(s) for s in shape), dtype=dtype))
    tf.squeeze = lambda x: T(torch.squeeze(_raw(x)))
    tf.reshape = lambda x, shape: T(torch.reshape(_raw(x), tuple(int(_raw(s)) for s in shape)))
    tf.shape = lambda x: T(torch.tensor(list(_raw(x).shape), dtype=torch.int64))
    tf.stop_gradient = lambda x: T(_raw(x).detach())
    tf.reduce_mean = lambda x: T(torch.mean(_raw(x)))
    tf.add_n = lambda xs: T(sum(_raw(x) for x in xs))
    tf.image = types.SimpleNamespace(
        rgb_to_yuv=lambda x: T(_raw(x) @ torch.tensor(_YUV_KERNEL, dtype=_raw(x).dtype, device=_raw(x).device)))
    tf.function = lambda fn: fn

    def py_function(func, inp, Tout):
        outs = func(*inp)                           # eager tensors in, eager tensors out
        single = not isinstance(Tout, (list, tuple))
        outs = [outs] if single else list(outs)
        touts = [Tout] if single else list(Tout)
        assert len(outs) == len(touts)
        res = [T(_raw(o).to(dt)) for o, dt in zip(outs, touts)]
        return res[0] if single else res
    tf.py_function = py_function

    def custom_gradient(f):
        def wrapper(*args):
            state = {}

            class Fn(torch.autograd.Function):
                @staticmethod
                def forward(ctx, *raw_in):
                    out, grad_fn = f(*[T(r) for r in raw_in])       # autograd is off inside forward, as TF stops recording
                    state["grad_fn"] = grad_fn
                    state["tuple"] = isinstance(out, (tuple, list))
                    outs = tuple(_raw(o) for o in out) if state["tuple"] else (_raw(out),)
                    return tuple(o.clone() for o in outs) if state["tuple"] else outs[0].clone()

                @staticmethod
                def backward(ctx, *dys):
                    dx = state["grad_fn"](*[T(d) for d in dys])
                    dx = dx if isinstance(dx, (tuple, list)) else (dx,)
                    assert len(dx) == len(args)
                    return tuple(None if d is None else _raw(d) for d in dx)

            res = Fn.apply(*[_raw(a) for a in args])
            return tuple(T(r) for r in res) if isinstance(res, tuple) else T(res)
        return wrapper
    tf.custom_gradient = custom_gradient

    tf.experimental = types.SimpleNamespace(dlpack=types.SimpleNamespace(
        to_dlpack=lambda t: torch.utils.dlpack.to_dlpack(_raw(t).detach()),
        from_dlpack=lambda cap: T(torch.utils.dlpack.from_dlpack(cap))))

    def sync_devices():
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    tf.test = types.SimpleNamespace(experimental=types.SimpleNamespace(sync_devices=sync_devices))

    class GradientTape:
        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def watch(self, t):
            pass

        @staticmethod
        def gradient(loss, variables):
            single = not isinstance(variables, (list, tuple))
            vs = [variables] if single else list(variables)
            gs = torch.autograd.grad(_raw(loss), [_raw(v) for v in vs], retain_graph=True, allow_unused=True)
            gs = [None if g is None else T(g) for g in gs]
            return gs[0] if single else gs
    tf.GradientTape = GradientTape
    return tf
