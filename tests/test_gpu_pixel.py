"""GPU parity of the pixel-side step through the C ABI: tf.image.resize / make_laplacian /
fold_laplacian_pyramid (+ backward) / RMSprop kernels against oracle/pixel_oracle.py on seeded inputs.
Tolerances: against the fp32 restatement (same arithmetic, up to fused multiply-adds) 3e-6 absolute on [0, 1] pixels;
against the fp64 restatement 1e-4 (TensorFlow, the oracle and the kernel all compute the source coordinate in fp32:
~1e-5 px at 512 px, times the slope of a white-noise test image); 1e-5 relative on gradients and optimizer state."""
import numpy as np
import pytest
import torch

from oracle import pixel_oracle as P

pytestmark = pytest.mark.gpu

ATOL = 1e-4          # vs fp64
ATOL32 = 3e-6        # vs the fp32 restatement
# content at the four scales of a default run plus awkward shapes
SHAPES = [(42, 64, 3), (85, 128, 3), (170, 256, 3), (341, 512, 3), (33, 7, 3), (2, 3, 1)]


@pytest.fixture(scope="module")
def S(cuda_device):
    import strotss_tensorflow_b200 as S
    return S


def _img(a, dev):            # (h, w, c) -> (1, h, w, c) CUDA
    return torch.tensor(np.ascontiguousarray(a), device=dev, dtype=torch.float32)[None]


@pytest.mark.parametrize("shape,out", [((10, 16, 3), (21, 32)), ((170, 256, 3), (341, 512)), ((341, 512, 3), (170, 256)),
                                       ((321, 481, 3), (42, 64)), ((1600, 1200, 3), (512, 384)), ((7, 5, 2), (7, 5)), ((1, 1, 3), (2, 3))])
def test_resize_bilinear(S, cuda_device, shape, out):
    rng = np.random.default_rng(5)
    x = rng.random(shape).astype(np.float32)
    from strotss_tensorflow_b200 import strotss_utils as U
    got = U._resize(_img(x, cuda_device), out[0], out[1])[0].cpu().numpy()
    ref = P.resize_bilinear(x, out[0], out[1], np.float64)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= ATOL
    assert np.abs(got - P.resize_bilinear(x, out[0], out[1], np.float32)).max() <= ATOL32


def test_resize_long_side_mirrors_utils_resize(S, cuda_device):
    x = _img(np.random.default_rng(6).random((321, 481, 3)), cuda_device)
    for scl, hw in [(64, (42, 64)), (128, (85, 128)), (256, (170, 256)), (512, (341, 512))]:
        assert tuple(S.resize(x, scl).shape[1:3]) == hw                       # nn/utils.py:32-37
    assert S.resize(x, None) is x
    assert tuple(S.resize_like(x, torch.empty(1, 9, 11, 3)).shape) == (1, 9, 11, 3)


@pytest.mark.parametrize("shape", SHAPES)
def test_laplacian_pyramid_and_fold(S, cuda_device, shape):
    rng = np.random.default_rng(7)
    x = rng.random(shape).astype(np.float32)
    xs = S.make_laplacian_pyramid(_img(x, cuda_device), 5)
    ref = P.make_laplacian_pyramid(x, 5, np.float64)
    assert len(xs) == 6
    for a, r in zip(xs, ref):
        assert tuple(a.shape[1:]) == r.shape
        assert np.abs(a[0].cpu().numpy() - r).max() <= ATOL
    img = S.fold_laplacian_pyramid(xs)
    assert np.abs(img[0].cpu().numpy() - x).max() <= ATOL                      # the pyramid folds back to the image
    # fold of arbitrary levels (the optimisation variables drift away from a true pyramid)
    lv = [rng.standard_normal(r.shape).astype(np.float32) for r in ref]
    got = S.fold_laplacian_pyramid([_img(a, cuda_device) for a in lv])[0].cpu().numpy()
    want = P.fold_laplacian_pyramid(lv, np.float64)
    assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())


def test_pyramid_against_reference_code_golden(S, cuda_device):
    """tests/golden/ref_pyramid.npz: make_laplacian_pyramid / fold_laplacian_pyramid / make_laplacian / utils.resize /
    utils.resize_like executed from the reference's own source (fp64) on a 21 x 30 image."""
    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden)
    import make_reference_golden as G
    z = np.load(os.path.join(golden, "ref_pyramid.npz"))
    img = G.pyramid_input()[0]
    x = _img(img, cuda_device)
    xs = S.make_laplacian_pyramid(x, 5)
    assert len(xs) == 6
    for k, a in enumerate(xs):
        assert tuple(a.shape) == z[f"pyr{k}"].shape
        assert np.abs(a[0].cpu().numpy() - z[f"pyr{k}"][0]).max() <= ATOL
    assert np.abs(S.fold_laplacian_pyramid(xs)[0].cpu().numpy() - z["fold"][0]).max() <= ATOL
    lap, down = S.make_laplacian(x, True)
    assert np.abs(lap[0].cpu().numpy() - z["lap"][0]).max() <= ATOL and np.abs(down[0].cpu().numpy() - z["down"][0]).max() <= ATOL
    small = S.resize(x, 16)
    assert tuple(small.shape) == z["resize16"].shape and np.abs(small[0].cpu().numpy() - z["resize16"][0]).max() <= ATOL
    like = S.resize_like(small, x)
    assert np.abs(like[0].cpu().numpy() - z["resize_like"][0]).max() <= ATOL


@pytest.mark.parametrize("shape", [(42, 64, 3), (341, 512, 3), (33, 7, 3)])
def test_fold_backward(S, cuda_device, shape):
    rng = np.random.default_rng(8)
    ref_levels = P.make_laplacian_pyramid(rng.random(shape), 5)
    ts = [_img(a, cuda_device).requires_grad_(True) for a in ref_levels]
    g = rng.standard_normal(shape).astype(np.float32)
    S.fold_laplacian_pyramid(ts).backward(_img(g, cuda_device))
    want = P.fold_laplacian_pyramid_backward([a.shape for a in ref_levels], g, np.float64)
    for t, w in zip(ts, want):
        got = t.grad[0].double().cpu().numpy()
        assert np.abs(got - w).max() <= 1e-5 * max(1.0, np.abs(w).max())
    # run-to-run bit-identical (gather form, no atomics)
    ts2 = [t.detach().clone().requires_grad_(True) for t in ts]
    S.fold_laplacian_pyramid(ts2).backward(_img(g, cuda_device))
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(ts, ts2))


def test_rmsprop_updates_all_variables_in_one_launch(S, cuda_device):
    rng = np.random.default_rng(9)
    shapes = [(1, 42, 64, 3), (1, 21, 32, 3), (1, 10, 16, 3), (1, 5, 8, 3), (1, 2, 4, 3), (1, 1, 2, 3)]
    vs = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    tv = [torch.tensor(v, device=cuda_device) for v in vs]
    opt = S.RMSprop(rho=0.99, epsilon=1e-8, learning_rate=2e-3)               # run_strotss.py:63
    ref_v = [v.astype(np.float64) for v in vs]
    ref_r = [np.zeros_like(v) for v in ref_v]
    h = S.shared_handle(cuda_device)
    for step in range(3):
        if step == 2:
            opt.lr = 1e-3                                                    # set_value(opt.lr, ...) between scales (:85,88)
        gs = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        l0 = h.launch_count
        opt.apply_gradients(zip([torch.tensor(g, device=cuda_device) for g in gs], tv))
        assert h.launch_count - l0 == 1
        for k in range(len(shapes)):
            ref_v[k], ref_r[k] = P.rmsprop_step(ref_v[k], ref_r[k], gs[k], 2e-3 if step < 2 else 1e-3, 0.99, 1e-8)
    for t, r in zip(tv, ref_v):
        assert np.abs(t.double().cpu().numpy() - r).max() <= 1e-5 * max(1.0, np.abs(r).max())


def test_rmsprop_one_optimizer_over_two_scales(S, cuda_device):
    """The reference keeps ONE optimizer for all scales and creates fresh variables per scale (run_strotss.py:63,89): the
    variables of the first scale are freed, their ids may be reused by larger ones, and every new variable must start from
    a zero slot of its own shape (Keras semantics)."""
    opt = S.RMSprop(rho=0.99, epsilon=1e-8, learning_rate=2e-3)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    for scale, shape in enumerate([(1, 8, 12, 3), (1, 16, 24, 3), (1, 32, 48, 3)]):
        import gc
        variables = [torch.rand(shape, generator=g, device=cuda_device) for _ in range(3)]
        start = [v.clone() for v in variables]
        grads = [torch.rand(shape, generator=g, device=cuda_device) - 0.5 for _ in range(3)]
        opt.apply_gradients(zip(grads, variables))
        torch.cuda.synchronize()
        for v, v0, gr in zip(variables, start, grads):
            rms = (1 - 0.99) * gr * gr                                   # zero slot at the first step of every scale
            want = v0 - 2e-3 * gr / (rms.sqrt() + 1e-8)
            assert torch.allclose(v, want, rtol=1e-5, atol=1e-7), f"scale {scale}: stale optimizer slot"
        assert len(opt.slots()) == 3
        del variables, start, grads
        gc.collect()


def test_pixel_side_argument_errors(S, cuda_device):
    with pytest.raises(ValueError):
        S.make_laplacian(torch.zeros(2, 4, 4, 3, device=cuda_device))           # batch must be 1
    with pytest.raises(RuntimeError, match="no CPU"):
        S.make_laplacian(torch.zeros(1, 4, 4, 3))
    with pytest.raises(ValueError):
        S.fold_laplacian_pyramid([torch.zeros(1, 4, 4, 3, device=cuda_device), torch.zeros(1, 2, 2, 1, device=cuda_device)])
