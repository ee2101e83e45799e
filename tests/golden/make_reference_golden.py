"""Golden vectors produced by the REFERENCE'S OWN loss code (test infrastructure; runs only where /root/reference exists).

The reference (interaction-lab-uh/STROTSS-tensorflow) cannot run here because TensorFlow is not installed.  Its loss path,
however, is plain Python over a dozen TensorFlow ops.  This script executes the reference's source text unmodified --

    nn/losses.py                      imported as a module (mae, cosine_distance, l2_distance, dist_metrics, reshape_2d,
                                      moment_matching, self_similarity, relaxed_emd)
    nn/strotss_utils.py:166-167       convert_rgb_to_yuv      (function source extracted with ast, exec'd)
    run_strotss.py:21-40              ContentLoss, StyleLoss  (class sources extracted with ast, exec'd)
    nn/strotss_utils.py:12-81         _clip_and_cast, Sampling._sample                       (fp32, bit-exact)
    nn/strotss_utils.py:139-163       make_laplacian, make_laplacian_pyramid, fold_laplacian_pyramid       
    nn/utils.py:14-41                 _validate_and_get_shape, resize, resize_like
    run_strotss.py:104-125,131-142    the two nested train_step functions (masked / unmasked), exec'd with stand-ins for
                                      their closure (a toy feature extractor instead of VGG, tf.GradientTape over torch autograd)

-- against a small stand-in for the `tensorflow` module (`TFShim` below) that implements exactly the ops this path calls,
on torch fp64 tensors, with TensorFlow's documented semantics (l2_normalize's epsilon inside the
square root, reduce_min gradient split among ties, maximum's gradient to the first argument on ties, the rgb_to_yuv
kernel).  The two lines that combine the losses (run_strotss.py:92,140) live inside a nested function of the driver and
are restated here.  Gradients come from torch autograd THROUGH the reference's op sequence.

What this pins: the oracle follows the reference's code -- op order, argument order, axes, broadcasting (e.g. the
column-sum division of self_similarity), weights.  What stays unpinned: that the shim's dozen ops equal TensorFlow's
(stated from TensorFlow's documentation, not executed).

tf.image.resize (bilinear, half-pixel centres, no antialiasing) is stood in for by torch's F.interpolate(mode="bilinear",
align_corners=False), which uses the same source coordinate (dst + 0.5) * in / out - 0.5 and the same edge clamping.

    python tests/golden/make_reference_golden.py            # writes tests/golden/ref_*.npz

No reference source is copied into this repository; the script reads it from /root/reference when it runs.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("STROTSS_REFERENCE", "/root/reference")

# the same cases as make_golden.py (inputs are regenerated from the seed by oracle.synth_problem)
CASES = {
    "small_d67": (48, 40, 67, 0.1, 11, 16.0),
    "ragged_d2179": (333, 517, 2179, 0.1, 5, 16.0),
    "default_d2179_eps1": (256, 256, 2179, 1.0, 0, 8.0),
    "near_d2179_eps001": (256, 200, 2179, 0.01, 7, 2.0),
}


# --------------------------------------------------------------------------------------------------- the stand-in
class _Shape(tuple):
    @property
    def dims(self):                       # tf.TensorShape.dims is a list (nn/losses.py:32 compares it with an int)
        return list(self)

    @property
    def rank(self):
        return len(self)

    def as_list(self):
        return list(self)


class T:
    """Stand-in for tf.Tensor: a torch tensor plus the few attributes / operators the reference's loss code touches."""

    def __init__(self, t):
        self.t = t

    @property
    def shape(self):
        return _Shape(self.t.shape)

    @property
    def dtype(self):
        return self.t.dtype

    def __getitem__(self, idx):
        if isinstance(idx, slice):                    # tf.range(h)[offset::step] with a scalar-tensor offset
            idx = slice(*(None if v is None else int(_raw(v)) for v in (idx.start, idx.stop, idx.step)))
        return T(self.t[_raw(idx)])

    def __lt__(self, other):
        return T(self.t < _raw(other))

    def __bool__(self):
        return bool(self.t)

    def __neg__(self):
        return T(-self.t)

    def __pow__(self, p):
        return T(self.t ** p)


def _raw(v):
    return v.t if isinstance(v, T) else v


def _binary(name, op):
    def fwd(a, b):
        return T(op(_raw(a), _raw(b)))

    def rev(a, b):
        return T(op(_raw(b), _raw(a)))
    setattr(T, f"__{name}__", fwd)
    setattr(T, f"__r{name}__", rev)


_binary("add", lambda a, b: a + b)
_binary("sub", lambda a, b: a - b)
_binary("mul", lambda a, b: a * b)
_binary("truediv", lambda a, b: a / b)
_binary("matmul", lambda a, b: a @ b)
_binary("floordiv", lambda a, b: a // b)

_YUV_KERNEL = [[0.299, -0.14714119, 0.61497538],      # tf.image.rgb_to_yuv: images (tensordot) kernel, rows = R, G, B
               [0.587, -0.28886916, -0.51496512],
               [0.114, 0.43601035, -0.10001026]]


def make_tf_shim() -> types.ModuleType:
    tf = types.ModuleType("tensorflow")
    tf.Tensor = T
    tf.float32, tf.float64 = torch.float32, torch.float64

    def reduce(fn):
        def op(x, axis=None, keepdims=False):
            x = _raw(x)
            return T(fn(x) if axis is None else fn(x, dim=axis, keepdim=keepdims))
        return op

    tf.reduce_mean = reduce(torch.mean)
    tf.reduce_sum = reduce(torch.sum)
    tf.reduce_min = reduce(torch.amin)            # amin splits the gradient equally among tied minima, as TensorFlow does
    tf.square = lambda x: T(_raw(x) ** 2)
    tf.abs = lambda x: T(torch.abs(_raw(x)))      # gradient sign(x), 0 at 0
    tf.sqrt = lambda x: T(torch.sqrt(_raw(x)))
    tf.squeeze = lambda x: T(torch.squeeze(_raw(x)))
    tf.reshape = lambda x, shape: T(torch.reshape(_raw(x), tuple(int(_raw(s)) for s in shape)))
    tf.int32 = torch.int32
    tf.split = lambda x, n, axis=0: [T(p) for p in torch.chunk(_raw(x), n, dim=axis)]
    tf.math = types.SimpleNamespace(floor=lambda x: T(torch.floor(_raw(x))))
    tf.clip_by_value = lambda x, lo, hi: T(torch.minimum(torch.maximum(_raw(x), _raw(lo)), _raw(hi)))
    tf.gather = lambda params, indices, axis=0: T(torch.index_select(_raw(params), axis, _raw(indices).long()))
    tf.concat = lambda xs, axis: T(torch.cat([_raw(x) for x in xs], dim=axis))
    tf.range = lambda n: T(torch.arange(int(_raw(n))))
    tf.meshgrid = lambda a, b: [T(m) for m in torch.meshgrid(_raw(a), _raw(b), indexing="xy")]      # TensorFlow's default indexing
    tf.reduce_max = reduce(torch.amax)
    tf.greater = lambda a, b: T(_raw(a) > _raw(b))
    tf.random = types.SimpleNamespace(shuffle=lambda x: x)      # index generation: the candidate list is compared as a set

    def gather_nd(params, indices):
        idx = _raw(indices).long()
        return T(_raw(params)[tuple(idx[:, k] for k in range(idx.shape[1]))])
    tf.gather_nd = gather_nd
    tf.shape = lambda x: T(torch.tensor(list(_raw(x).shape), dtype=torch.int64))      # a 1-D integer tensor, as in TensorFlow

    def cast(x, dtype):
        x = _raw(x)
        return T(x.to(dtype) if torch.is_tensor(x) else torch.tensor(x, dtype=dtype))
    tf.cast = cast

    def maximum(a, b):
        a, b = _raw(a), _raw(b)
        a = a if torch.is_tensor(a) else torch.tensor(a, dtype=b.dtype)
        b = b if torch.is_tensor(b) else torch.tensor(b, dtype=a.dtype)
        return T(torch.where(a >= b, a, b))       # TensorFlow's gradient mask is a >= b: ties go to the first argument
    tf.maximum = maximum

    def matmul(a, b, transpose_a=False, transpose_b=False):
        a, b = _raw(a), _raw(b)
        return T((a.T if transpose_a else a) @ (b.T if transpose_b else b))
    tf.matmul = matmul

    tf.nn = types.SimpleNamespace()

    def l2_normalize(x, axis=None, epsilon=1e-12):
        x = _raw(x)
        sq = torch.sum(x * x, dim=axis, keepdim=True)
        return T(x * torch.rsqrt(torch.clamp_min(sq, epsilon)))      # x * rsqrt(max(sum(x^2), epsilon))
    tf.nn.l2_normalize = l2_normalize

    tf.image = types.SimpleNamespace()
    tf.image.rgb_to_yuv = lambda x: T(_raw(x) @ torch.tensor(_YUV_KERNEL, dtype=_raw(x).dtype))

    def resize(images, size, method="bilinear"):
        assert method == "bilinear"
        x = _raw(images)                                                     # NHWC or HWC
        size = [int(v) for v in (_raw(size).tolist() if torch.is_tensor(_raw(size)) else size)]
        x4 = x if x.dim() == 4 else x[None]
        y = torch.nn.functional.interpolate(x4.permute(0, 3, 1, 2), size=size, mode="bilinear", align_corners=False, antialias=False)
        y = y.permute(0, 2, 3, 1)
        return T(y if x.dim() == 4 else y[0])
    tf.image.resize = resize

    class GradientTape:                            # tape.gradient(loss, variables) over torch autograd
        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        @staticmethod
        def gradient(loss, variables):
            gs = torch.autograd.grad(_raw(loss), [_raw(v) for v in variables], retain_graph=True, allow_unused=True)
            return [None if g is None else T(g) for g in gs]
    tf.GradientTape = GradientTape
    tf.function = lambda fn: fn

    class Module:
        def __init__(self, **kwargs):
            pass

        @staticmethod
        def with_name_scope(fn):
            return fn
    tf.Module = Module
    return tf


# --------------------------------------------------------------------------------------- running the reference's text
def _source_of(path, name):
    text = open(path).read()
    for node in ast.parse(text).body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == name:
            return ast.get_source_segment(text, node)
    raise KeyError(f"{name} not found in {path}")


def load_reference(tf=None):
    """Returns a namespace with the reference's own relaxed_emd, moment_matching, self_similarity, convert_rgb_to_yuv,
    ContentLoss and StyleLoss, bound to the stand-in tensorflow module."""
    tf = tf or make_tf_shim()
    saved = sys.modules.get("tensorflow")
    sys.modules["tensorflow"] = tf
    try:
        spec = importlib.util.spec_from_file_location("_strotss_reference_losses", os.path.join(REFERENCE, "nn", "losses.py"))
        losses = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(losses)
    finally:
        if saved is None:
            del sys.modules["tensorflow"]
        else:
            sys.modules["tensorflow"] = saved
    utils_ns = {"tf": tf}
    exec(_source_of(os.path.join(REFERENCE, "nn", "strotss_utils.py"), "convert_rgb_to_yuv"), utils_ns)
    ns = {"tf": tf, "moment_matching": losses.moment_matching, "relaxed_emd": losses.relaxed_emd,
          "self_similarity": losses.self_similarity, "strotss": types.SimpleNamespace(convert_rgb_to_yuv=utils_ns["convert_rgb_to_yuv"])}
    for cls in ("ContentLoss", "StyleLoss"):
        exec(_source_of(os.path.join(REFERENCE, "run_strotss.py"), cls), ns)
    return types.SimpleNamespace(tf=tf, losses=losses, convert_rgb_to_yuv=utils_ns["convert_rgb_to_yuv"],
                                 ContentLoss=ns["ContentLoss"], StyleLoss=ns["StyleLoss"])


def load_reference_widened(tf=None):
    """The reference's own Sampling (with _clip_and_cast), make_laplacian, make_laplacian_pyramid, fold_laplacian_pyramid
    (nn/strotss_utils.py) and resize, resize_like (nn/utils.py), exec'd from their source text over the stand-in."""
    import math
    from functools import partialmethod
    from typing import List, Optional, Tuple, Union
    tf = tf or make_tf_shim()
    su = os.path.join(REFERENCE, "nn", "strotss_utils.py")
    ut = os.path.join(REFERENCE, "nn", "utils.py")
    ns = {"tf": tf, "math": math, "partialmethod": partialmethod, "List": List, "Optional": Optional, "Tuple": Tuple,
          "Union": Union, "tf_rng": None}
    for name in ("_clip_and_cast", "Sampling", "make_laplacian", "make_laplacian_pyramid", "fold_laplacian_pyramid"):
        exec(_source_of(su, name), ns)
    for name in ("_validate_and_get_shape", "resize", "resize_like"):
        exec(_source_of(ut, name), ns)
    return types.SimpleNamespace(**{k: ns[k] for k in ("Sampling", "make_laplacian", "make_laplacian_pyramid",
                                                        "fold_laplacian_pyramid", "resize", "resize_like")}, tf=tf, ns=ns)


class _FixedRng:
    """Stands in for nn/rand.py's tf_rng inside Sampling._make_indices: returns the given offsets in turn."""

    def __init__(self, values):
        self.values = list(values)

    def uniform(self, shape, minval, maxval, dtype=None):
        v = self.values.pop(0)
        assert minval <= v < maxval
        return T(torch.tensor(v))


INDEX_CASES = {
    # name: (h, w, bilinear, (off_x, off_y), mask kind)
    "nearest_42x64": (42, 64, False, (0, 0), None),
    "bilinear_341x512_off12": (341, 512, True, (1, 2), None),
    "bilinear_341x512_rect": (341, 512, True, (2, 3), "rect"),
    "bilinear_170x256_empty": (170, 256, True, (0, 1), "empty"),
    "nearest_42x64_lowres_mask": (42, 64, False, (0, 0), "lowres"),
}


def index_mask(kind, h, w):
    if kind is None:
        return None
    if kind == "rect":
        m = np.zeros((h, w, 1), np.float32); m[: h // 3, : w // 2] = 1
    elif kind == "empty":
        m = np.zeros((h, w, 1), np.float32)
    else:                                   # a mask at another resolution: resized with tf.image.resize (:105)
        m = np.zeros((2 * h + 1, 3 * w, 1), np.float32); m[h // 2:, w:] = 1
    return m


def evaluate_indices(wid, h, w, bilinear, offsets, mask):
    """All candidate (row, col) pairs of Sampling._make_indices (nn/strotss_utils.py:83-121) in the reference's order
    (tf.random.shuffle stood in for by the identity, sample_size larger than the candidate count)."""
    wid.ns["tf_rng"] = _FixedRng(offsets)
    s = wid.Sampling(10 ** 9)
    base = T(torch.zeros(1, h, w, 3))
    ret = s._make_indices(base, bilinear, None if mask is None else T(torch.tensor(mask)))
    return ret.t.numpy().astype(np.int16)


SAMPLER_SHAPES = [(42, 64, 3), (42, 64, 8), (42, 64, 8), (21, 32, 16), (21, 32, 16), (10, 16, 32), (10, 16, 32), (10, 16, 32),
                  (5, 8, 64), (5, 8, 64)]            # the ten maps of the content image at scale 64, fewer channels
SAMPLER_N = 32


def sampler_inputs(seed=0):
    rng = np.random.default_rng(seed)
    xs = [rng.standard_normal((1,) + s).astype(np.float32) for s in SAMPLER_SHAPES]
    idx = np.stack([rng.uniform(0, 42, SAMPLER_N), rng.uniform(0, 64, SAMPLER_N)], axis=1).astype(np.float32)
    idx[:4] = np.floor(idx[:4])                      # integral positions, the border row and the border column
    idx[4] = [41.0, 63.0]
    idx[5] = [41.7, 63.9]
    return xs, idx


def evaluate_sampler(wid, xs, idx, bilinear):
    s = wid.Sampling(SAMPLER_N)
    out = s._sample([T(torch.tensor(x)) for x in xs], T(torch.tensor(idx)), bilinear)
    return out.t.numpy().copy()


def pyramid_input(seed=1):
    return np.random.default_rng(seed).uniform(0, 1, (1, 21, 30, 3))


def evaluate_pyramid(wid, img):
    x = T(torch.tensor(img, dtype=torch.float64))
    pyr = wid.make_laplacian_pyramid(x, levels=5)
    fold = wid.fold_laplacian_pyramid(pyr)
    lap, down = wid.make_laplacian(x, return_downscale=True)
    small = wid.resize(x, 16)
    like = wid.resize_like(small, x)
    out = {f"pyr{k}": p.t.numpy().copy() for k, p in enumerate(pyr)}
    out.update(fold=fold.t.numpy().copy(), lap=lap.t.numpy().copy(), down=down.t.numpy().copy(), resize16=small.t.numpy().copy(),
               resize_like=like.t.numpy().copy())
    return out


# ------------------------------------------------------------------------------ the driver's train_step functions
def _train_step_sources():
    text = open(os.path.join(REFERENCE, "run_strotss.py")).read()
    nodes = [n for n in ast.walk(ast.parse(text)) if isinstance(n, ast.FunctionDef) and n.name == "train_step"]
    nodes.sort(key=lambda n: n.lineno)
    assert len(nodes) == 2                        # masked (run_strotss.py:104-125) first, unmasked (:131-142) second
    return [textwrap.dedent(ast.get_source_segment(text, n, padded=True)) for n in nodes]


def _toy_vgg(img):
    """Stands in for VGG (nn/model.py): two deterministic 'feature maps' of an NHWC image, differentiable."""
    x = _raw(img)
    m1 = torch.tensor([[0.5, -0.2, 0.1, 0.3], [0.1, 0.4, -0.3, 0.2], [-0.2, 0.1, 0.6, 0.1]], dtype=x.dtype)
    m2 = torch.tensor(np.random.default_rng(7).standard_normal((3, 6)), dtype=x.dtype)
    f1 = torch.relu(x @ m1 + 0.1)
    pooled = torch.nn.functional.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    f2 = torch.relu(pooled @ m2 + 0.2)
    return [T(f1), T(f2)]


def evaluate_train_steps(alpha=16.0, sample_size=40, dtype=torch.float64):
    """Runs the reference's masked and unmasked train_step on a toy problem and records the per-region sampled features, so
    that the composition (per-region StyleLoss targets, alpha / loss_denom weighting, mean over regions) can be checked."""
    ref = load_reference()
    wid = load_reference_widened(ref.tf)
    tf = ref.tf
    rng = np.random.default_rng(11)
    content = T(torch.tensor(rng.uniform(0, 1, (1, 24, 32, 3)), dtype=dtype))
    style = T(torch.tensor(rng.uniform(0, 1, (1, 20, 28, 3)), dtype=dtype))
    start = T(torch.tensor(rng.uniform(0, 1, (1, 24, 32, 3)), dtype=dtype))
    st_variables = [T(v.t.clone().requires_grad_(True)) for v in wid.make_laplacian_pyramid(start, levels=2)]
    cm = np.zeros((2, 24, 32, 1), np.float32); cm[0, :, :16] = 1; cm[1, :, 16:] = 1
    sm = np.zeros((2, 20, 28, 1), np.float32); sm[0, :10] = 1; sm[1, 10:] = 1
    content_masks = [T(torch.tensor(m)) for m in cm]
    style_masks = [T(torch.tensor(m)) for m in sm]
    wid.ns["tf_rng"] = _FixedRng([0] * 64)
    sampling = wid.Sampling(sample_size)
    records = []
    inner = sampling.bilinear

    def recording_bilinear(xs, ys=None, mask=None):
        c, p = inner(xs, ys, mask=mask)
        records.append((c, p))
        return c, p
    sampling.bilinear = recording_bilinear
    vgg = _toy_vgg
    content_feat = [content] + vgg(content)
    style_feat = [style] + vgg(style)
    loss_denom = (2. + alpha + 1. / max(alpha, 1.))                               # run_strotss.py:92
    loss_styles = [ref.StyleLoss(sampling(style_feat, mask=m), alpha=alpha) for m in style_masks]      # :97-101
    loss_style = ref.StyleLoss(sampling(style_feat), alpha=alpha)                  # :128
    ns = dict(tf=tf, strotss=types.SimpleNamespace(fold_laplacian_pyramid=wid.fold_laplacian_pyramid), st_variables=st_variables,
              vgg=vgg, content_masks=content_masks, sampling=sampling, content_feat=content_feat, loss_content=ref.ContentLoss(),
              loss_styles=loss_styles, loss_style=loss_style, alpha=alpha, loss_denom=loss_denom)
    out = {}
    for tag, src in zip(("masked", "plain"), _train_step_sources()):
        del records[:]
        scope = dict(ns)
        exec(src, scope)
        res = scope["train_step"]()
        out[tag + "_loss"] = res["loss"].t.item()
        out[tag + "_loss_c"] = res["loss_c"].t.item()
        out[tag + "_loss_s"] = res["loss_s"].t.item()
        assert all(g is not None for g in res["grads"])                            # the tape reaches the pyramid variables
        gp = torch.autograd.grad(res["loss"].t, [p.t for _, p in records], retain_graph=True)
        for r, ((c, p), g) in enumerate(zip(records, gp)):
            out[f"{tag}_content{r}"] = c.t.detach().numpy().copy()
            out[f"{tag}_pred{r}"] = p.t.detach().numpy().copy()
            out[f"{tag}_grad{r}"] = g.numpy().copy()
        out[tag + "_regions"] = len(records)
    for r, ls in enumerate(loss_styles):
        out[f"masked_style{r}"] = ls.target.t.detach().numpy().copy()
    out["plain_style0"] = loss_style.target.t.detach().numpy().copy()
    out["alpha"] = alpha
    return out


def evaluate(ref, style, content, pred, alpha, dtype=torch.float64):
    """The loss lines of train_step (run_strotss.py:136-141) on given sampled features."""
    st = T(torch.tensor(style, dtype=dtype))
    co = T(torch.tensor(content, dtype=dtype))
    pr_t = torch.tensor(pred, dtype=dtype, requires_grad=True)
    pr = T(pr_t)
    loss_content = ref.ContentLoss()
    loss_style = ref.StyleLoss(st, alpha=alpha)
    loss_denom = (2. + alpha + 1. / max(alpha, 1.))             # run_strotss.py:92
    loss_c = loss_content(co, pr)
    loss_s = loss_style(pr)
    loss = (alpha * loss_c + loss_s) / loss_denom               # run_strotss.py:140
    loss.t.backward()
    with torch.no_grad():
        l_m = ref.losses.moment_matching(st, T(pr_t)).t.item()
        l_remd = ref.losses.relaxed_emd(st, T(pr_t)).t.item()
        l_pal = ref.losses.relaxed_emd(ref.convert_rgb_to_yuv(st), ref.convert_rgb_to_yuv(T(pr_t)), distance="both").t.item()
        l_pal_l2 = ref.losses.relaxed_emd(ref.convert_rgb_to_yuv(st), ref.convert_rgb_to_yuv(T(pr_t)), distance="l2").t.item()
    return dict(total=loss.t.item(), loss_c=loss_c.t.item(), loss_s=loss_s.t.item(), l_m=l_m, l_remd=l_remd,
                l_palette=l_pal, l_palette_l2=l_pal_l2, grad=pr_t.grad.numpy().copy())


def main():
    from oracle import strotss_oracle as O
    ref = load_reference()
    for name, (N, M, D, eps, seed, alpha) in CASES.items():
        st, co, pr = O.synth_problem(N, M, D, eps=eps, seed=seed)
        r = evaluate(ref, st, co, pr, alpha)
        g = r.pop("grad")
        out = dict(N=N, M=M, D=D, eps=eps, seed=seed, alpha=alpha, **r, grad_norm=np.linalg.norm(g), grad_rowsum=g.sum(axis=1),
                   grad_colsum=g.sum(axis=0), grad_head=g[:8, :16].copy(),
                   input_checksum=np.array([st.astype(np.float64).sum(), co.astype(np.float64).sum(), pr.astype(np.float64).sum()]))
        if D <= 128:
            out["grad"] = g
        np.savez_compressed(os.path.join(HERE, "ref_" + name + ".npz"), **out)
        print(name, "total", out["total"], "grad_norm", out["grad_norm"])
    wid = load_reference_widened()
    xs, idx = sampler_inputs()
    np.savez_compressed(os.path.join(HERE, "ref_sampler.npz"), indices=idx, bilinear=evaluate_sampler(wid, xs, idx, True),
                        nearest=evaluate_sampler(wid, xs, idx, False), input_checksum=np.array([float(x.astype(np.float64).sum()) for x in xs]))
    np.savez_compressed(os.path.join(HERE, "ref_pyramid.npz"), **evaluate_pyramid(wid, pyramid_input()))
    np.savez_compressed(os.path.join(HERE, "ref_indices.npz"),
                        **{name: evaluate_indices(wid, h, w, bil, off, index_mask(kind, h, w))
                           for name, (h, w, bil, off, kind) in INDEX_CASES.items()})
    np.savez_compressed(os.path.join(HERE, "ref_train_step.npz"), **evaluate_train_steps())
    print("sampler, pyramid, indices, train_step written")


if __name__ == "__main__":
    main()
