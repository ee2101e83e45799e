"""CPU tests pinning the oracle (the reference ships no tests -- SURVEY.md section 4/8c -- so these are
the pins the repo creates itself): analytic known answers, invariances, fp64 finite differences,
an independent autodiff cross-check and the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import strotss_oracle as O
from oracle import torch_port as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rand(n, d, seed):
    return np.random.default_rng(seed).standard_normal((n, d))


# ---------------------------------------------------------------- known answers
def test_remd_identical_inputs_is_zero():
    x = np.abs(_rand(20, 9, 0)) + 0.1
    assert O.relaxed_emd(x, x, "cosine") == pytest.approx(0.0, abs=1e-12)


def test_palette_identical_inputs_hits_clamp_floor():
    # l2_distance clamps the squared distance at 1e-6 before / D and sqrt (nn/losses.py:23)
    x = np.abs(_rand(15, 3, 1)) + 0.1
    assert O.relaxed_emd(x, x, "both") == pytest.approx(np.sqrt(1e-6 / 3), rel=1e-9)
    assert O.relaxed_emd(x, x, "l2") == pytest.approx(np.sqrt(1e-6 / 3), rel=1e-9)


def test_orthogonal_rows_cost_one_and_selfsim_uniform():
    x = np.eye(6)[:, :6] * np.arange(1, 7)[:, None]
    C = O.cosine_distance(x, x)
    off = C[~np.eye(6, dtype=bool)]
    assert np.allclose(off, 1.0) and np.allclose(np.diag(C), 0.0)
    Xn = C / C.sum(axis=0)
    assert np.allclose(Xn[~np.eye(6, dtype=bool)], 1.0 / 5)


def test_zero_row_normalises_to_zero_distance_one():
    x = np.array([[0.0, 0.0, 0.0], [1.0, 2.0, 3.0]])
    y = np.array([[3.0, 1.0, 2.0]])
    C = O.cosine_distance(x, y)
    assert C[0, 0] == pytest.approx(1.0)


def test_same_inputs_give_zero_losses():
    x = np.abs(_rand(12, 7, 2))
    assert O.self_similarity(x, x) == pytest.approx(0.0, abs=1e-15)
    assert O.moment_matching(x, x) == pytest.approx(0.0, abs=1e-15)


def test_bad_distance_raises_keyerror():
    with pytest.raises(KeyError):
        O.relaxed_emd(np.ones((2, 3)), np.ones((2, 3)), "manhattan")


def test_yuv_matrix_and_wrapper_weights():
    rgb = np.array([[1.0, 1.0, 1.0, 5.0]])
    yuv = O.convert_rgb_to_yuv(rgb)
    assert yuv[0, 0] == pytest.approx(1.0) and abs(yuv[0, 1]) < 1e-6 and abs(yuv[0, 2]) < 1e-6
    st, co, pr = O.synth_problem(10, 9, 12, 0.5, 3)
    for alpha in (16.0, 0.5):
        l_s = O.style_loss(st, pr, alpha)
        parts = (O.moment_matching(st, pr) + O.relaxed_emd(st, pr)
                 + O.relaxed_emd(O.convert_rgb_to_yuv(st), O.convert_rgb_to_yuv(pr), "both") / max(alpha, 1.0))
        assert l_s == pytest.approx(parts, rel=1e-12)
        tot = O.total_loss(st, co, pr, alpha)
        assert tot == pytest.approx((alpha * O.self_similarity(pr, co) + l_s) / (2 + alpha + 1 / max(alpha, 1)), rel=1e-12)
    assert [O.loss_denom(a) for a in (16.0, 8.0, 4.0, 2.0)] == [18.0625, 10.125, 6.25, 4.5]


def test_reshape_2d_matches_reference_quirk():
    assert O.reshape_2d(np.zeros((1, 5, 7))).shape == (5, 7)
    assert O.reshape_2d(np.zeros((2, 3, 4))).shape == (6, 4)
    assert O.reshape_2d(np.zeros((1, 7))).shape == (1, 7)


# ---------------------------------------------------------------- invariances
def test_cosine_losses_invariant_to_row_scaling_and_grad_orthogonal():
    st, co, pr = O.synth_problem(14, 11, 20, 0.7, 4)
    scale = np.random.default_rng(0).uniform(0.5, 3.0, size=(14, 1))
    l0, g0, _ = O.relaxed_emd(st, pr, "cosine", want_grad=True)
    assert O.relaxed_emd(st, pr * scale, "cosine") == pytest.approx(l0, rel=1e-12)
    assert np.abs((g0 * pr).sum(axis=1)).max() < 1e-12
    s0, gs, _ = O.self_similarity(pr, co, want_grad=True)
    assert O.self_similarity(pr * scale, co) == pytest.approx(s0, rel=1e-10)
    assert np.abs((gs * pr).sum(axis=1)).max() < 1e-12


def test_permutation_invariance_and_remd_symmetry():
    st, co, pr = O.synth_problem(13, 10, 16, 0.7, 5)
    perm = np.random.default_rng(1).permutation(13)
    assert O.total_loss(st, co[perm], pr[perm]) == pytest.approx(O.total_loss(st, co, pr), rel=1e-11)
    assert O.relaxed_emd(st, pr) == pytest.approx(O.relaxed_emd(pr, st), rel=1e-13)


def test_tie_rules_follow_tensorflow():
    # two identical prediction rows tie for every target row: reduce_min splits the gradient equally
    x = np.array([[1.0, 0.2, 0.0]])
    y = np.array([[0.5, 1.0, 0.3], [0.5, 1.0, 0.3], [0.0, 0.1, 1.0]])
    _, g, info = O.relaxed_emd(x, y, "cosine", want_grad=True)
    # R_X (1 row) vs R_Y: whichever branch, rows 0 and 1 of y must get identical gradients
    assert np.allclose(g[0], g[1])
    # exact R_X == R_Y tie -> branch X (tf.maximum sends the gradient to its first argument)
    x2 = np.array([[1.0, 0.0], [0.0, 1.0]])
    _, _, info2 = O.relaxed_emd(x2, x2[::-1].copy(), "cosine", want_grad=True)
    assert info2["R_X"] == info2["R_Y"] and info2["branch_x"]


# ---------------------------------------------------------------- gradients
def _fd(f, x, idx, h=1e-6):
    xp = x.copy(); xm = x.copy()
    xp[idx] += h; xm[idx] -= h
    return (f(xp) - f(xm)) / (2 * h)


@pytest.mark.parametrize("alpha", [16.0, 0.5])
def test_total_gradient_matches_finite_differences(alpha):
    st, co, pr = [a.astype(np.float64) for a in O.synth_problem(9, 8, 10, 0.5, 6)]
    pr = pr + 0.05                      # keep away from the ReLU-created exact zeros / kinks
    _, g, _ = O.total_loss(st, co, pr, alpha, np.float64, True)
    rng = np.random.default_rng(2)
    for _ in range(12):
        idx = (rng.integers(0, 9), rng.integers(0, 10))
        fd = _fd(lambda p: O.total_loss(st, co, p, alpha), pr, idx)
        assert g[idx] == pytest.approx(fd, rel=2e-4, abs=1e-9)


@pytest.mark.parametrize("dist", ["cosine", "l2", "both"])
def test_remd_gradient_matches_finite_differences(dist):
    rng = np.random.default_rng(7)
    x = rng.uniform(0.1, 1.0, (7, 3)); y = rng.uniform(0.1, 1.0, (6, 3))
    _, g, _ = O.relaxed_emd(x, y, dist, np.float64, True)
    for i in range(6):
        for c in range(3):
            fd = _fd(lambda p: O.relaxed_emd(x, p, dist), y, (i, c), h=1e-7)
            assert g[i, c] == pytest.approx(fd, rel=1e-4, abs=1e-8)


def test_closed_forms_match_torch_autograd():
    st, co, pr = O.synth_problem(40, 33, 67, 0.3, 8)
    loss, grad, info = O.total_loss(st, co, pr, 4.0, np.float64, True)
    l2, g2, i2 = T.total_loss_and_grad(torch.tensor(st, dtype=torch.float64), torch.tensor(co, dtype=torch.float64),
                                       torch.tensor(pr, dtype=torch.float64), 4.0)
    assert float(l2) == pytest.approx(loss, rel=1e-12)
    assert np.linalg.norm(g2.numpy() - grad) / np.linalg.norm(grad) < 1e-12
    assert float(i2["l_palette"]) == pytest.approx(info["l_palette"], rel=1e-12)


def test_fp32_oracle_noise_floor_is_far_below_tolerance():
    st, co, pr = O.synth_problem(64, 48, 2179, 0.1, 9)
    l64, g64, _ = O.total_loss(st, co, pr, 16.0, np.float64, True)
    l32, g32, _ = O.total_loss(st, co, pr, 16.0, np.float32, True)
    assert abs(l32 - l64) / l64 < 1e-5
    assert np.linalg.norm(g32 - g64) / np.linalg.norm(g64) < 1e-3


# ---------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", ["small_d67", "ragged_d2179", "default_d2179_eps1", "near_d2179_eps001"])
def test_oracle_reproduces_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    N, M, D = int(z["N"]), int(z["M"]), int(z["D"])
    st, co, pr = O.synth_problem(N, M, D, eps=float(z["eps"]), seed=int(z["seed"]))
    chk = np.array([st.astype(np.float64).sum(), co.astype(np.float64).sum(), pr.astype(np.float64).sum()])
    assert np.allclose(chk, z["input_checksum"], rtol=1e-12), "synthetic generator drifted"
    if "style" in z.files:
        assert np.array_equal(st, z["style"]) and np.array_equal(pr, z["pred"])
    loss, grad, info = O.total_loss(st, co, pr, float(z["alpha"]), np.float64, True)
    assert loss == pytest.approx(float(z["total"]), rel=1e-10)
    for key in ("loss_c", "loss_s", "l_m", "l_remd", "l_palette"):
        assert info[key] == pytest.approx(float(z[key]), rel=1e-10)
    assert np.array_equal(info["remd"]["row_argmin"], z["remd_row_argmin"])
    assert np.array_equal(info["remd"]["col_argmin"], z["remd_col_argmin"])
    assert np.linalg.norm(grad) == pytest.approx(float(z["grad_norm"]), rel=1e-9)
    assert np.allclose(grad[:8, :16], z["grad_head"], rtol=1e-8, atol=1e-14)


def test_masked_total_is_mean_of_region_totals():
    """run_strotss.py:112-124: per-region totals are summed and divided by the number of regions."""
    probs = [O.synth_problem(N, M, 35, eps=0.2, seed=70 + r) for r, (N, M) in enumerate([(40, 30), (17, 30), (5, 9)])]
    styles, contents, preds = zip(*probs)
    loss, grads, info = O.masked_total_loss(styles, contents, preds, 4.0, np.float64, True)
    singles = [O.total_loss(s, c, p, 4.0, np.float64, True) for s, c, p in probs]
    assert abs(loss - sum(t[0] for t in singles) / 3) < 1e-15
    for g, t in zip(grads, singles):
        assert np.allclose(g, t[1] / 3, rtol=0, atol=1e-18)
    assert abs(info["loss_c"] - sum(t[2]["loss_c"] for t in singles) / 3) < 1e-15
    # central difference on one entry of the second region
    h = 1e-6
    p2 = [p.astype(np.float64).copy() for p in preds]
    p2[1][3, 7] += h
    up = O.masked_total_loss(styles, contents, p2, 4.0, np.float64)
    p2[1][3, 7] -= 2 * h
    dn = O.masked_total_loss(styles, contents, p2, 4.0, np.float64)
    assert abs((up - dn) / (2 * h) - grads[1][3, 7]) < 1e-6 * max(1.0, abs(grads[1][3, 7]))
