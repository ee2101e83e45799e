"""CPU oracle for the STROTSS per-iteration loss hot path.  TEST INFRASTRUCTURE ONLY.

PARITY STATUS: pinned to the reference's own Python code, NOT to TensorFlow.  The reference
(interaction-lab-uh/STROTSS-tensorflow) ships no tests, golden vectors or fixtures for this path, and
TensorFlow cannot be imported in this environment.  tests/golden/make_reference_golden.py therefore
executes the reference's source text unmodified (nn/losses.py imported; ContentLoss, StyleLoss,
convert_rgb_to_yuv exec'd from run_strotss.py / nn/strotss_utils.py) over a stand-in for the dozen
TensorFlow ops the path calls, and commits the results as tests/golden/ref_*.npz: this restatement
reproduces them to <= 5e-14 (losses, every term, the gradient; tests/test_reference_golden.py), ties
included.  That pins op order, argument order, axes, broadcasting and weights to the reference's
code.  UNPINNED remains that those dozen ops behave as TensorFlow's do (stated from TensorFlow's
documentation, listed below); beyond that the oracle is held by analytic known-answer cases,
invariances and fp64 finite-difference checks in tests/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker.  The product path (strotss_tensorflow_b200) never
imports it and has no CPU fallback.

This is a NumPy restatement, op by op, of
  nn/losses.py:8-9    mae
  nn/losses.py:12-15  cosine_distance
  nn/losses.py:18-24  l2_distance
  nn/losses.py:27-28  dist_metrics
  nn/losses.py:31-36  reshape_2d
  nn/losses.py:39-52  moment_matching
  nn/losses.py:55-66  self_similarity
  nn/losses.py:69-80  relaxed_emd
  nn/strotss_utils.py:166-167  convert_rgb_to_yuv
  run_strotss.py:21-40         ContentLoss / StyleLoss
  run_strotss.py:65,92,140,155 alpha schedule, loss_denom, total loss
with hand-written reverse-mode gradients that encode TensorFlow's op semantics
(third-party, unpinned `tensorflow` in requirements.txt:1; README.md:11 says >= 2.6.0):
  * tf.nn.l2_normalize(x, axis, epsilon=1e-12) = x * rsqrt(max(sum(x^2), epsilon))
  * tf.reduce_min gradient is shared equally among tied minima
  * tf.maximum(a, b) gradient goes entirely to `a` when a >= b
  * tf.abs gradient is sign(x) (0 at 0); tf.sqrt gradient is 0.5 / sqrt(x)
  * tf.image.rgb_to_yuv uses the fixed 3x3 kernel below.
Every function takes a `dtype` so the same code runs as fp64 (truth) or fp32 (what a
TF fp32 run would roughly give).  Matrices are materialised exactly like the reference.
"""
from __future__ import annotations

import numpy as np

# tf.image.rgb_to_yuv kernel (rows = R, G, B inputs; columns = Y, U, V outputs).
RGB_TO_YUV = np.array(
    [[0.299, -0.14714119, 0.61497538],
     [0.587, -0.28886916, -0.51496512],
     [0.114, 0.43601035, -0.10001026]], dtype=np.float64)

L2N_EPS = 1e-12      # tf.nn.l2_normalize default epsilon (nn/losses.py:13-14)
L2D_CLAMP = 1e-06    # nn/losses.py:23
COLSUM_CLAMP = 1e-12  # nn/losses.py:60,63


# --------------------------------------------------------------------------------------
# forward building blocks
# --------------------------------------------------------------------------------------
def reshape_2d(x, channel_axis: int = -1):
    """nn/losses.py:31-36.  The rank test at :32 never fires, so always squeeze+reshape."""
    x = np.asarray(x)
    x = np.squeeze(x)
    if x.ndim == 0:
        x = x.reshape(1, 1)
    if x.ndim == 1:
        # tf.reshape(x, (-1, shape[-1])) of a rank-1 tensor gives one row
        return x.reshape(1, x.shape[0])
    return x.reshape(-1, x.shape[channel_axis])


def mae(x, y):
    """nn/losses.py:8-9."""
    return np.mean(np.abs(x - y))


def l2_normalize(x, dtype=np.float64):
    x = np.asarray(x, dtype=dtype)
    ss = np.sum(x * x, axis=1, keepdims=True)
    r = 1.0 / np.sqrt(np.maximum(ss, dtype(L2N_EPS)))
    return x * r, ss, r


def _l2_normalize_bwd(x, ss, r, g):
    """Reverse of x * rsqrt(max(ss, eps)).  tf.maximum sends the gradient to ss when ss >= eps."""
    gx = g * r
    live = (ss >= L2N_EPS)
    dot = np.sum(g * x, axis=1, keepdims=True)
    gx = gx - np.where(live, x * dot * r ** 3, 0.0)
    return gx


def cosine_distance(x, y, dtype=np.float64):
    """nn/losses.py:12-15."""
    xh, _, _ = l2_normalize(x, dtype)
    yh, _, _ = l2_normalize(y, dtype)
    return 1 - xh @ yh.T


def l2_distance(x, y, dtype=np.float64):
    """nn/losses.py:18-24."""
    x = np.asarray(x, dtype=dtype)
    y = np.asarray(y, dtype=dtype)
    x_sq = np.sum(x ** 2, axis=1).reshape(-1, 1)
    y_sq = np.sum(y ** 2, axis=1).reshape(1, -1)
    matrix = x_sq + y_sq - dtype(2.0) * (x @ y.T)
    matrix = np.maximum(matrix, dtype(L2D_CLAMP)) / dtype(x.shape[1])
    return np.sqrt(matrix)


dist_metrics = {
    'cosine': cosine_distance,
    'l2': l2_distance,
    'both': lambda x, y, dtype=np.float64: cosine_distance(x, y, dtype) + l2_distance(x, y, dtype),
}


def convert_rgb_to_yuv(x, dtype=np.float64):
    """nn/strotss_utils.py:166-167."""
    x = np.asarray(x, dtype=dtype)
    return x[:, :3] @ RGB_TO_YUV.astype(dtype)


# --------------------------------------------------------------------------------------
# distance matrices with their reverse pass w.r.t. the SECOND argument
# --------------------------------------------------------------------------------------
def _cosine_fwd(x, y, dtype):
    xh, xss, xr = l2_normalize(x, dtype)
    yh, yss, yr = l2_normalize(y, dtype)
    C = 1 - xh @ yh.T
    return C, (xh, np.asarray(y, dtype=dtype), yh, yss, yr)


def _cosine_bwd_y(ctx, dC):
    xh, y, yh, yss, yr = ctx
    gyh = -(dC.T @ xh)
    return _l2_normalize_bwd(y, yss, yr, gyh)


def _l2_fwd(x, y, dtype):
    x = np.asarray(x, dtype=dtype)
    y = np.asarray(y, dtype=dtype)
    x_sq = np.sum(x ** 2, axis=1).reshape(-1, 1)
    y_sq = np.sum(y ** 2, axis=1).reshape(1, -1)
    m = x_sq + y_sq - dtype(2.0) * (x @ y.T)
    mc = np.maximum(m, dtype(L2D_CLAMP)) / dtype(x.shape[1])
    out = np.sqrt(mc)
    return out, (x, y, m, out)


def _l2_bwd_y(ctx, dC):
    x, y, m, out = ctx
    D = x.shape[1]
    dm = dC * (0.5 / out) / D
    dm = np.where(m >= L2D_CLAMP, dm, 0.0)      # tf.maximum(matrix, 1e-6): grad to matrix iff >=
    # m_ij = |x_i|^2 + |y_j|^2 - 2 x_i.y_j  ->  dm/dy_j = 2 y_j - 2 x_i
    gy = 2.0 * y * np.sum(dm, axis=0).reshape(-1, 1) - 2.0 * (dm.T @ x)
    return gy


def _dist_fwd(x, y, distance, dtype):
    if distance not in dist_metrics:
        raise KeyError(distance)          # nn/losses.py:74: dict lookup raises KeyError
    if distance == 'cosine':
        C, ctx = _cosine_fwd(x, y, dtype)
        return C, ('cosine', ctx)
    if distance == 'l2':
        C, ctx = _l2_fwd(x, y, dtype)
        return C, ('l2', ctx)
    C1, c1 = _cosine_fwd(x, y, dtype)
    C2, c2 = _l2_fwd(x, y, dtype)
    return C1 + C2, ('both', (c1, c2))


def _dist_bwd_y(tagged, dC):
    tag, ctx = tagged
    if tag == 'cosine':
        return _cosine_bwd_y(ctx, dC)
    if tag == 'l2':
        return _l2_bwd_y(ctx, dC)
    return _cosine_bwd_y(ctx[0], dC) + _l2_bwd_y(ctx[1], dC)


# --------------------------------------------------------------------------------------
# the three losses (value, gradient w.r.t. the prediction operand, diagnostics)
# --------------------------------------------------------------------------------------
def relaxed_emd(x, y, distance: str = 'cosine', dtype=np.float64, want_grad: bool = False):
    """nn/losses.py:69-80.  x = target (rows of C), y = prediction (columns of C).

    Returns loss, or (loss, dL/dy, info) when want_grad.  info carries the observables the
    CUDA path is compared on: R_X, R_Y, the argmins and the fp gap to each runner-up.
    """
    x = reshape_2d(x)
    y = reshape_2d(y)
    C, ctx = _dist_fwd(x, y, distance, dtype)
    row_min = C.min(axis=1)
    col_min = C.min(axis=0)
    R_X = row_min.mean()
    R_Y = col_min.mean()
    loss = np.maximum(R_X, R_Y)
    if not want_grad:
        return loss
    M, N = C.shape
    dC = np.zeros_like(C)
    if R_X >= R_Y:        # tf.maximum: ties go to the first argument
        sel = (C == row_min[:, None])
        dC = sel / sel.sum(axis=1, keepdims=True) / M
    else:
        sel = (C == col_min[None, :])
        dC = sel / sel.sum(axis=0, keepdims=True) / N
    gy = _dist_bwd_y(ctx, dC.astype(dtype))

    def _gap(Cm):
        if Cm.shape[1] < 2:
            return np.full(Cm.shape[0], np.inf)
        part = np.partition(Cm, 1, axis=1)
        return part[:, 1] - part[:, 0]

    info = dict(R_X=R_X, R_Y=R_Y, branch_x=bool(R_X >= R_Y),
                row_argmin=C.argmin(axis=1), col_argmin=C.argmin(axis=0),
                row_gap=_gap(C), col_gap=_gap(C.T), row_min=row_min, col_min=col_min)
    return loss, gy, info


def moment_matching(x, y, dtype=np.float64, want_grad: bool = False):
    """nn/losses.py:39-52.  Gradient w.r.t. y (the prediction; run_strotss.py:35)."""
    x = reshape_2d(x).astype(dtype)
    y = reshape_2d(y).astype(dtype)
    xm = x.mean(axis=0, keepdims=True)
    ym = y.mean(axis=0, keepdims=True)
    cx = x - xm
    cy = y - ym
    xv = cx.T @ cx / dtype(x.shape[0])
    yv = cy.T @ cy / dtype(y.shape[0])
    loss = mae(xv, yv) + mae(xm, ym)
    if not want_grad:
        return loss
    N, D = y.shape
    gV = np.sign(yv - xv) / (D * D)      # d mean|xv - yv| / d yv
    gm = np.sign(ym - xm) / D
    gcy = cy @ (gV + gV.T) / N
    gy = gcy - gcy.mean(axis=0, keepdims=True) + gm / N
    info = dict(l_cov=mae(xv, yv), l_mean=mae(xm, ym))
    return loss, gy.astype(dtype), info


def self_similarity(x, y, dtype=np.float64, want_grad: bool = False):
    """nn/losses.py:55-66.  Gradient w.r.t. x (the prediction; run_strotss.py:24)."""
    x = reshape_2d(x).astype(dtype)
    y = reshape_2d(y).astype(dtype)
    xh, xss, xr = l2_normalize(x, dtype)
    yh, _, _ = l2_normalize(y, dtype)
    Xd = 1 - xh @ xh.T
    s = Xd.sum(axis=0)
    sc = np.maximum(s, dtype(COLSUM_CLAMP))
    Xn = Xd / sc
    Yd = 1 - yh @ yh.T
    t = np.maximum(Yd.sum(axis=0), dtype(COLSUM_CLAMP))
    Yn = Yd / t
    Ny = y.shape[0]
    loss = mae(Xn, Yn) * dtype(Ny)
    if not want_grad:
        return loss
    N = x.shape[0]
    gXn = np.sign(Xn - Yn) * (Ny / (N * N))
    gXd = gXn / sc
    gsc = -(gXn * Xd).sum(axis=0) / sc ** 2
    gs = np.where(s >= COLSUM_CLAMP, gsc, 0.0)
    gXd = gXd + gs[None, :]
    # Xd = 1 - a b^T with a = b = xh (x is normalised twice at nn/losses.py:13-14; same values)
    gxh = -(gXd @ xh) - (gXd.T @ xh)
    gx = _l2_normalize_bwd(x, xss, xr, gxh)
    info = dict(s=s, t=t)
    return loss, gx.astype(dtype), info


# --------------------------------------------------------------------------------------
# wrappers (run_strotss.py:21-40) and the total (run_strotss.py:92,140)
# --------------------------------------------------------------------------------------
def content_loss(target, prediction, dtype=np.float64, want_grad: bool = False):
    """ContentLoss.__call__(target, prediction) -> self_similarity(prediction, target)."""
    return self_similarity(prediction, target, dtype=dtype, want_grad=want_grad)


def style_loss(target, prediction, alpha: float, dtype=np.float64, want_grad: bool = False):
    """StyleLoss(target, alpha)(prediction): l_m + l_remd + l_palette / max(alpha, 1)."""
    inv_alpha = 1.0 / max(alpha, 1.0)
    if not want_grad:
        l_m = moment_matching(target, prediction, dtype)
        l_r = relaxed_emd(target, prediction, 'cosine', dtype)
        l_p = relaxed_emd(convert_rgb_to_yuv(target, dtype), convert_rgb_to_yuv(prediction, dtype), 'both', dtype)
        return l_m + l_r + inv_alpha * l_p
    l_m, g_m, i_m = moment_matching(target, prediction, dtype, True)
    l_r, g_r, i_r = relaxed_emd(target, prediction, 'cosine', dtype, True)
    l_p, g_pyuv, i_p = relaxed_emd(convert_rgb_to_yuv(target, dtype), convert_rgb_to_yuv(prediction, dtype),
                                   'both', dtype, True)
    g = g_m + g_r
    g[:, :3] += inv_alpha * (g_pyuv @ RGB_TO_YUV.astype(dtype).T)
    info = dict(l_m=l_m, l_remd=l_r, l_palette=l_p, moment=i_m, remd=i_r, palette=i_p)
    return l_m + l_r + inv_alpha * l_p, g, info


def loss_denom(alpha: float) -> float:
    """run_strotss.py:92."""
    return 2.0 + alpha + 1.0 / max(alpha, 1.0)


def total_loss(style, content, pred, alpha: float = 16.0, dtype=np.float64, want_grad: bool = False):
    """run_strotss.py:138-140: (alpha * loss_c + loss_s) / loss_denom and d/d pred."""
    den = loss_denom(alpha)
    if not want_grad:
        l_c = content_loss(content, pred, dtype)
        l_s = style_loss(style, pred, alpha, dtype)
        return (alpha * l_c + l_s) / den
    l_c, g_c, i_c = content_loss(content, pred, dtype, True)
    l_s, g_s, i_s = style_loss(style, pred, alpha, dtype, True)
    loss = (alpha * l_c + l_s) / den
    grad = (alpha * g_c + g_s) / den
    info = dict(loss_c=l_c, loss_s=l_s, **i_s)
    return loss, grad, info


def masked_total_loss(styles, contents, preds, alpha: float = 16.0, dtype=np.float64, want_grad: bool = False):
    """The masked train_step (run_strotss.py:112-124): per region r, `(alpha * loss_c_r + loss_s_r) / loss_denom`
    with that region's StyleLoss target (:99-101); the sum is divided by the number of regions (:121).
    Returns loss (and the list of per-region gradients, and info with the averaged loss_c / loss_s of :122-123)."""
    R = len(styles)
    if not (R == len(contents) == len(preds)) or R == 0:
        raise ValueError("one style / content / prediction matrix per region")
    if not want_grad:
        return sum(total_loss(s, c, p, alpha, dtype) for s, c, p in zip(styles, contents, preds)) / R
    loss, grads, lc, ls = 0.0, [], 0.0, 0.0
    for s, c, p in zip(styles, contents, preds):
        l, g, info = total_loss(s, c, p, alpha, dtype, True)
        loss += l / R
        grads.append(g / R)
        lc += info["loss_c"] / R
        ls += info["loss_s"] / R
    return loss, grads, dict(loss_c=lc, loss_s=ls)


# --------------------------------------------------------------------------------------
# synthetic inputs (seed-0 convention of nn/rand.py:12-21)
# --------------------------------------------------------------------------------------
def synth_features(n: int, d: int, rng: np.random.Generator, sigma=None):
    """Channels 0-2: U[0,1) 'RGB'.  Channels 3..: ReLU-like heavy-tailed 'VGG' activations."""
    out = np.empty((n, d), dtype=np.float32)
    k = min(3, d)
    out[:, :k] = rng.random((n, k), dtype=np.float32)
    if d > 3:
        if sigma is None:
            sigma = np.exp(rng.standard_normal(d - 3)).astype(np.float32)
        z = rng.standard_normal((n, d - 3), dtype=np.float32)
        out[:, 3:] = np.maximum(0.0, sigma[None, :] * (z + 0.3))
    return out


def synth_problem(N: int, M: int, D: int = 2179, eps: float = 1.0, seed: int = 0):
    """style (M,D), content (N,D), pred (N,D) = max(0, content + eps*mean|content|*noise)."""
    rng = np.random.default_rng(seed)
    sigma = np.exp(rng.standard_normal(max(D - 3, 0))).astype(np.float32)
    style = synth_features(M, D, rng, sigma)
    content = synth_features(N, D, rng, sigma)
    noise = rng.standard_normal((N, D), dtype=np.float32)
    pred = np.maximum(0.0, content + np.float32(eps) * np.mean(np.abs(content)) * noise).astype(np.float32)
    return style, content, pred


# --------------------------------------------------------------------------------------
# SURVEY section 8(f) "next #1": hypercolumn sampler (nn/strotss_utils.py:25-136)
# --------------------------------------------------------------------------------------
def sampler_scales(shapes):
    """Cumulative index divisors per feature map, nn/strotss_utils.py:31-37.

    shapes: list of (h, w, c).  When a map is lower than its predecessor the indices are divided by
    prev.shape[index] / cur.shape[index], where `index` is fixed at the FIRST such map: the height axis if
    that map's height is a power of two, else the width axis (:35).  Returns the per-map divisor applied
    at that map (1.0 = none); divisions accumulate in float32 in the reference (`indices /= y`)."""
    import math
    divs = []
    index = None
    for i, (h, w, c) in enumerate(shapes):
        d = 1.0
        if i > 0 and h < shapes[i - 1][0]:
            if index is None:
                index = 0 if not (math.log2(h) % 1) else 1      # position in (h, w) == axis 1 or 2 of NHWC
            d = shapes[i - 1][index] / shapes[i][index]
        divs.append(d)
    return divs


def sample_hypercolumns(xs, indices, bilinear_sampling: bool):
    """Sampling._sample (nn/strotss_utils.py:25-81).  xs: list of (1, h, w, c) or (h, w, c) float32 arrays;
    indices: (n, 2) float32 (row coordinate, column coordinate).  Returns (n, sum c) float32."""
    maps = [np.asarray(x, dtype=np.float32).reshape(x.shape[-3], x.shape[-2], x.shape[-1]) for x in xs]
    divs = sampler_scales([m.shape for m in maps])
    idx = np.asarray(indices, dtype=np.float32).copy()
    feats = []
    for cur, d in zip(maps, divs):
        if d != 1.0:
            idx = (idx / np.float32(d)).astype(np.float32)
        h, w, c = cur.shape
        gx, gy = idx[:, 0], idx[:, 1]
        flat = cur.reshape(h * w, c)
        if bilinear_sampling:
            gxf = np.floor(gx); dx = gx - gxf
            gyf = np.floor(gy); dy = gy - gyf
            wa = ((1 - dx) * (1 - dy))[:, None]; wb = ((1 - dx) * dy)[:, None]
            wc = (dx * (1 - dy))[:, None]; wd = (dx * dy)[:, None]
            gxi = np.clip(gxf, 0, h - 1).astype(np.int32); gyi = np.clip(gyf, 0, w - 1).astype(np.int32)
            gxb = np.clip(gxi + 1, 0, h - 1); gyb = np.clip(gyi + 1, 0, w - 1)
            g = (flat[gxi * w + gyi] * wa + flat[gxi * w + gyb] * wb + flat[gxb * w + gyi] * wc + flat[gxb * w + gyb] * wd)
        else:
            gxi = np.clip(gx, 0, h - 1).astype(np.int32); gyi = np.clip(gy, 0, w - 1).astype(np.int32)
            g = flat[gxi * w + gyi]
        feats.append(g.astype(np.float32))
    return np.concatenate(feats, axis=1)


def sampler_grid(h, w, bilinear_sampling: bool, off_x=0, off_y=0):
    """The deterministic part of Sampling._make_indices (nn/strotss_utils.py:87-103): the strided grid of
    candidate (row, col) pairs before masking / shuffling.  off_x, off_y stand for the two tf_rng draws."""
    import math
    if bilinear_sampling:
        area = math.sqrt((h * w) // (128 ** 2))
        step_x, step_y = max(1, math.floor(area)), max(1, math.ceil(area))
        X = np.arange(h)[off_x::step_x]
        Y = np.arange(w)[off_y::step_y]
    else:
        X, Y = np.arange(h), np.arange(w)
    XX, YY = np.meshgrid(X, Y)
    return np.stack([XX.reshape(-1), YY.reshape(-1)], axis=1), (max(1, math.floor(math.sqrt((h * w) // (128 ** 2)))) if bilinear_sampling else 1,
                                                                  max(1, math.ceil(math.sqrt((h * w) // (128 ** 2)))) if bilinear_sampling else 1)
