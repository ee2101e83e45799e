"""CPU pins of the pixel-side oracle (oracle/pixel_oracle.py): the restated tf.image.resize /
Laplacian pyramid / RMSprop semantics against independent implementations (torch) and exact properties."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pixel_oracle as P


def _t(x):          # (h, w, c) -> (1, c, h, w)
    return torch.tensor(x).permute(2, 0, 1)[None]


@pytest.mark.parametrize("shape,out", [((10, 16, 3), (21, 32)), ((21, 32, 3), (42, 64)), ((170, 256, 3), (341, 512)),
                                       ((341, 512, 3), (170, 256)), ((321, 481, 3), (42, 64)), ((7, 5, 2), (7, 5)), ((1, 1, 3), (2, 3)),
                                       ((9, 4, 1), (1, 1))])
def test_resize_matches_half_pixel_bilinear(shape, out):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape)
    got = P.resize_bilinear(x, out[0], out[1], np.float64)
    ref = F.interpolate(_t(x), size=out, mode="bilinear", align_corners=False, antialias=False)[0].permute(1, 2, 0).numpy()
    assert np.allclose(got, ref, rtol=0, atol=1e-12)
    got32 = P.resize_bilinear(x.astype(np.float32), out[0], out[1], np.float32)
    assert np.allclose(got32, ref, rtol=0, atol=1e-4)       # fp32 source coordinates (~1e-5 px at 512 px) times unit-variance noise


def test_resize_weights_are_a_partition_of_unity_and_constant_preserving():
    for i, o in [(10, 21), (341, 170), (5, 5), (1, 4)]:
        R = P._axis_weights(i, o, np.float64)
        assert np.allclose(R.sum(axis=1), 1.0) and (R >= 0).all()
    x = np.full((6, 9, 3), 0.37)
    assert np.allclose(P.resize_bilinear(x, 13, 17), 0.37)


def test_pyramid_folds_back_to_the_image():
    """fold(make_laplacian_pyramid(x)) == x: every level adds back exactly what was subtracted (strotss_utils.py:139-163)."""
    rng = np.random.default_rng(1)
    for shape in [(42, 64, 3), (85, 128, 3), (341, 512, 3), (33, 7, 3)]:
        x = rng.random(shape)
        xs = P.make_laplacian_pyramid(x, 5)
        assert len(xs) == 6 and xs[0].shape == shape
        assert xs[-1].shape[:2] == (max(shape[0] >> 5, 1), max(shape[1] >> 5, 1))
        assert np.allclose(P.fold_laplacian_pyramid(xs), x, rtol=0, atol=1e-13)


def test_fold_backward_matches_autograd():
    rng = np.random.default_rng(2)
    x = rng.random((42, 64, 3))
    xs = P.make_laplacian_pyramid(x, 5)
    ts = [_t(a).clone().requires_grad_(True) for a in xs]
    ret = ts[-1]
    for t in reversed(ts[:-1]):
        ret = t + F.interpolate(ret, size=t.shape[-2:], mode="bilinear", align_corners=False)
    g = rng.standard_normal(x.shape)
    ret.backward(_t(g))
    grads = P.fold_laplacian_pyramid_backward([a.shape for a in xs], g)
    for t, gr in zip(ts, grads):
        assert np.allclose(t.grad[0].permute(1, 2, 0).numpy(), gr, rtol=0, atol=1e-12)


def test_rmsprop_matches_torch_rmsprop():
    """Keras RMSprop (momentum 0, not centred, epsilon outside the sqrt) == torch.optim.RMSprop(alpha=rho, eps)."""
    rng = np.random.default_rng(3)
    v0 = rng.standard_normal((5, 7))
    p = torch.nn.Parameter(torch.tensor(v0))
    opt = torch.optim.RMSprop([p], lr=2e-3, alpha=0.99, eps=1e-8)
    var, rms = v0.copy(), np.zeros_like(v0)
    for step in range(4):
        g = rng.standard_normal(v0.shape)
        p.grad = torch.tensor(g)
        opt.step()
        var, rms = P.rmsprop_step(var, rms, g, 2e-3, 0.99, 1e-8)
        assert np.allclose(p.detach().numpy(), var, rtol=0, atol=1e-14)
