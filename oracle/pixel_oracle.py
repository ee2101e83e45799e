"""TEST INFRASTRUCTURE (not product code): CPU restatement of the pixel-side step of the reference iteration
(SURVEY.md section 8f "next #3"), used only by tests/ to check the CUDA kernels.  Parity: the pyramid / resize helpers are
pinned to the reference's own code (tests/golden/make_reference_golden.py executes nn/strotss_utils.py:139-163 and
nn/utils.py:32-41 over a stand-in for tf.image.resize; this restatement agrees to 1e-12, tests/test_reference_golden.py).
Unpinned: tf.image.resize itself and Keras RMSprop -- TensorFlow cannot be installed here, their semantics are restated
from the documented behaviour (and checked against torch's independent F.interpolate / optim.RMSprop).

  tf.image.resize(x, size, method='bilinear')   TF2: half-pixel centres, antialias=False
       in = (out + 0.5) * in_size/out_size - 0.5 ; lo = max(floor(in), 0) ; hi = min(ceil(in), in_size - 1) ;
       lerp = in - floor(in) ; out = top + (bottom - top) * y_lerp, top = tl + (tr - tl) * x_lerp
  make_laplacian / make_laplacian_pyramid / fold_laplacian_pyramid       nn/strotss_utils.py:139-163
  utils.resize / resize_like                                             nn/utils.py:32-41
  tf.keras.optimizers.RMSprop(rho=0.99, epsilon=1e-8, learning_rate=lr)  run_strotss.py:63,148
       (momentum 0, not centred: rms = rho*rms + (1-rho)*g^2 ; var -= lr * g / (sqrt(rms) + epsilon))
Images are (h, w, c) arrays (the reference's (1, h, w, c) without the batch axis)."""
from __future__ import annotations

import numpy as np


def _axis_taps(in_size: int, out_size: int, dtype):
    """lo, hi, lerp of every output index along one axis."""
    scale = dtype(in_size) / dtype(out_size)
    o = np.arange(out_size, dtype=dtype)
    src = (o + dtype(0.5)) * scale - dtype(0.5)
    f = np.floor(src)
    lo = np.maximum(f.astype(np.int64), 0)
    hi = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
    return lo, hi, (src - f).astype(dtype)


def _axis_weights(in_size: int, out_size: int, dtype):
    """Dense (out_size, in_size) interpolation matrix of one axis (rows sum to 1)."""
    lo, hi, w = _axis_taps(in_size, out_size, dtype)
    R = np.zeros((out_size, in_size), dtype=dtype)
    np.add.at(R, (np.arange(out_size), lo), 1 - w)
    np.add.at(R, (np.arange(out_size), hi), w)
    return R


def resize_bilinear(x, oh: int, ow: int, dtype=np.float64):
    """tf.image.resize(x, (oh, ow)) for x of shape (h, w, c)."""
    x = np.asarray(x, dtype=dtype)
    ylo, yhi, yw = _axis_taps(x.shape[0], oh, dtype)
    xlo, xhi, xw = _axis_taps(x.shape[1], ow, dtype)
    xw = xw[None, :, None]; yw = yw[:, None, None]
    top = x[ylo][:, xlo] + (x[ylo][:, xhi] - x[ylo][:, xlo]) * xw
    bot = x[yhi][:, xlo] + (x[yhi][:, xhi] - x[yhi][:, xlo]) * xw
    return top + (bot - top) * yw


def resize_bilinear_transpose(g, sh: int, sw: int, dtype=np.float64):
    """Gradient of resize_bilinear w.r.t. its (sh, sw, c) input, given g of the output's shape."""
    g = np.asarray(g, dtype=dtype)
    Ry = _axis_weights(sh, g.shape[0], dtype)
    Rx = _axis_weights(sw, g.shape[1], dtype)
    t = np.tensordot(Ry.T, g, axes=(1, 0))                 # (sh, ow, c)
    return np.transpose(np.tensordot(Rx.T, t, axes=(1, 1)), (1, 0, 2))      # (sw, sh, c) -> (sh, sw, c)


def make_laplacian(x, dtype=np.float64):
    """nn/strotss_utils.py:139-146 with return_downscale=True -> (pyr, down)."""
    x = np.asarray(x, dtype=dtype)
    h, w = x.shape[:2]
    hd, wd = max(h // 2, 1), max(w // 2, 1)
    down = resize_bilinear(x, hd, wd, dtype)
    return x - resize_bilinear(down, h, w, dtype), down


def make_laplacian_pyramid(x, levels: int = 5, dtype=np.float64):
    """nn/strotss_utils.py:149-156."""
    xs, cur = [], np.asarray(x, dtype=dtype)
    for _ in range(levels):
        pyr, cur = make_laplacian(cur, dtype)
        xs.append(pyr)
    xs.append(cur)
    return xs


def fold_laplacian_pyramid(xs, dtype=np.float64):
    """nn/strotss_utils.py:159-163."""
    ret = np.asarray(xs[-1], dtype=dtype)
    for x in reversed(xs[:-1]):
        ret = np.asarray(x, dtype=dtype) + resize_bilinear(ret, x.shape[0], x.shape[1], dtype)
    return ret


def fold_laplacian_pyramid_backward(shapes, grad_out, dtype=np.float64):
    """d loss / d xs[k] for every level given d loss / d image (what tape.gradient returns, run_strotss.py:122,141)."""
    grads = [np.asarray(grad_out, dtype=dtype)]
    for k in range(1, len(shapes)):
        grads.append(resize_bilinear_transpose(grads[-1], shapes[k][0], shapes[k][1], dtype))
    return grads


def rmsprop_step(var, rms, grad, lr: float, rho: float = 0.99, eps: float = 1e-8, dtype=np.float64):
    """One Keras RMSprop update (momentum 0, not centred) -> (new var, new rms)."""
    var, rms, grad = (np.asarray(a, dtype=dtype) for a in (var, rms, grad))
    rms = dtype(rho) * rms + dtype(1 - rho) * grad * grad
    return var - dtype(lr) * grad / (np.sqrt(rms) + dtype(eps)), rms
