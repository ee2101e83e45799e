"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the
CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): <= 1e-3 relative on every loss, <= 1e-2 relative on the
gradient norm, exact argmin agreement wherever the fp64 gap to the runner-up exceeds GAP_THR
(bf16 operands cannot resolve nearer ties; see DESIGN.md "precision").  The direction of the
gradient is additionally checked by cosine similarity.
"""
import os

import numpy as np
import pytest
import torch

from oracle import strotss_oracle as O

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-3
GRADNORM_RTOL = 1e-2
GAP_THR = 4e-3          # cosine-distance gap below which bf16 operands may flip an argmin
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def S(cuda_device):
    import strotss_tensorflow_b200 as S
    return S


@pytest.fixture(scope="module")
def handle(S, cuda_device):
    return S.shared_handle(cuda_device)


def _t(a, dev):
    return torch.tensor(np.ascontiguousarray(a), device=dev, dtype=torch.float32)


def _gcheck(g, ref, cos_min=0.995):
    g = g.detach().double().cpu().numpy()
    nr = np.linalg.norm(ref)
    assert abs(np.linalg.norm(g) - nr) / nr <= GRADNORM_RTOL
    cos = float((g * ref).sum() / (np.linalg.norm(g) * nr))
    assert cos >= cos_min, f"gradient direction off: cos={cos}"
    return cos


# ------------------------------------------------------------------------------ GEMM core
@pytest.mark.parametrize("m,n,k,tile", [(128, 256, 64, 256), (128, 128, 64, 128), (300, 500, 2179, 256),
                                        (300, 500, 2179, 128), (77, 33, 100, 128), (1, 1, 3, 256), (513, 257, 2240, 256)])
def test_tcgen05_gemm_core(handle, cuda_device, m, n, k, tile):
    g = torch.Generator().manual_seed(m * 7 + n)
    A = torch.randn(m, k, generator=g).to(cuda_device)
    B = torch.randn(n, k, generator=g).to(cuda_device)
    C = handle.debug_gemm(A, B, 0.5, tile)
    ref = 0.5 * (A.bfloat16().double() @ B.bfloat16().double().T)
    assert (C.double() - ref).abs().max().item() <= 1e-5 * k ** 0.5 * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (300, 500, 2048), (1000, 2179, 200), (64, 16, 64), (129, 257, 130)])
def test_tcgen05_gemm_transposed_a(handle, cuda_device, m, n, k):
    """MN-major A operand (used by stage 2b of the symmetric self-similarity) incl. the accumulate epilogue."""
    g = torch.Generator().manual_seed(m + 3 * n + k)
    At = torch.randn(k, m, generator=g).to(cuda_device)
    B = torch.randn(n, k, generator=g).to(cuda_device)
    ref = At.bfloat16().double().T @ B.bfloat16().double().T
    C = handle.debug_gemm_ta(At, B, 1.0)
    tol = 1e-5 * k ** 0.5 * ref.abs().max().item() + 1e-6
    assert (C.double() - ref).abs().max().item() <= tol
    C2 = handle.debug_gemm_ta(At, B, 0.5, C=C.clone())
    assert (C2.double() - 1.5 * ref).abs().max().item() <= 2 * tol


@pytest.mark.parametrize("m,n,k", [(256, 256, 64), (300, 500, 2048), (2179, 700, 1000), (64, 16, 64), (129, 257, 130)])
def test_tcgen05_pair_gemm_mn_major_a(handle, cuda_device, m, n, k):
    """CTA-pair kernel with an MN-major A operand: how stage 2 of the self-similarity reads the row-major x^ as x^T."""
    g = torch.Generator().manual_seed(2 * m + 5 * n + k)
    At = torch.randn(k, m, generator=g).to(cuda_device)
    B = torch.randn(n, k, generator=g).to(cuda_device)
    ref = At.bfloat16().double().T @ B.bfloat16().double().T
    C = handle.debug_gemm_ta(At, B, 1.0, variant=1)
    tol = 1e-5 * k ** 0.5 * ref.abs().max().item() + 1e-6
    assert (C.double() - ref).abs().max().item() <= tol


@pytest.mark.parametrize("m,k", [(256, 64), (300, 2048), (2179, 1000), (64, 130), (513, 16384)])
def test_tcgen05_pair_gemm_gram_both_mn_major(handle, cuda_device, m, k):
    """Both operands MN-major from ONE row-major matrix: the covariance cen^T cen without a transposed copy."""
    g = torch.Generator().manual_seed(3 * m + k)
    At = torch.randn(k, m, generator=g).to(cuda_device)
    ref = At.bfloat16().double().T @ At.bfloat16().double()
    C = handle.debug_gemm_ta(At, torch.zeros(m, 1, device=cuda_device), 0.5, variant=2)
    tol = 1e-5 * k ** 0.5 * ref.abs().max().item() + 1e-6
    assert (C.double() - 0.5 * ref).abs().max().item() <= tol


# ------------------------------------------------------------------------------ relaxed EMD
@pytest.mark.parametrize("N,M,seed", [(700, 517, 0), (333, 1024, 1), (128, 256, 2), (1, 300, 3), (130, 1, 4),
                                      (2300, 300, 5), (2561, 2100, 6)])       # > 2048: CTA-pair kernels, ragged tile couples
def test_relaxed_emd_cosine(handle, cuda_device, N, M, seed):
    st, co, pr = O.synth_problem(N, M, 2179, eps=1.0, seed=seed)
    out, grad, ra, ca = handle.relaxed_emd(_t(st, cuda_device), _t(pr, cuda_device), "cosine", True, True)
    l64, g64, info = O.relaxed_emd(st, pr, "cosine", np.float64, True)
    out = out.cpu().numpy()
    assert abs(out[0] - l64) / l64 <= LOSS_RTOL
    assert abs(out[1] - info["R_X"]) / info["R_X"] <= LOSS_RTOL and abs(out[2] - info["R_Y"]) / info["R_Y"] <= LOSS_RTOL
    ra = ra.cpu().numpy(); ca = ca.cpu().numpy()
    clear_r = info["row_gap"] > GAP_THR
    clear_c = info["col_gap"] > GAP_THR
    assert np.array_equal(ra[clear_r], info["row_argmin"][clear_r])
    assert np.array_equal(ca[clear_c], info["col_argmin"][clear_c])
    # near-ties may pick the runner-up, never anything worse than the gap threshold
    C = O.cosine_distance(st, pr)
    assert np.all(C[np.arange(M), ra] - C.min(axis=1) <= GAP_THR)
    assert np.all(C[ca, np.arange(N)] - C.min(axis=0) <= GAP_THR)
    if abs(info["R_X"] - info["R_Y"]) > 1e-4:
        assert bool(out[3]) == info["branch_x"]
        _gcheck(grad, g64, cos_min=0.98)
        # gradient of a cosine loss is orthogonal to its input row
        assert float((grad * _t(pr, cuda_device)).sum(dim=1).abs().max()) <= 1e-4 * float(grad.abs().max()) * float(np.abs(pr).max()) * 50


def test_relaxed_emd_is_symmetric_and_zero_on_identical(handle, cuda_device):
    st, co, pr = O.synth_problem(400, 300, 2179, eps=1.0, seed=5)
    a = handle.relaxed_emd(_t(st, cuda_device), _t(pr, cuda_device), "cosine", False)[0].cpu().numpy()
    b = handle.relaxed_emd(_t(pr, cuda_device), _t(st, cuda_device), "cosine", False)[0].cpu().numpy()
    assert a[0] == b[0] and a[1] == b[2] and a[2] == b[1]
    z = handle.relaxed_emd(_t(st, cuda_device), _t(st, cuda_device), "cosine", False, True)
    assert abs(z[0][0].item()) <= 1e-2          # bf16 dot of a unit row with itself: 1 - O(2^-9)
    assert np.array_equal(z[2].cpu().numpy(), np.arange(300))


@pytest.mark.parametrize("dist", ["both", "l2", "cosine"])
def test_relaxed_emd_three_channel(handle, cuda_device, dist):
    rng = np.random.default_rng(3)
    a = O.convert_rgb_to_yuv(rng.uniform(0.02, 1.0, (777, 3)), np.float32).astype(np.float32)
    b = O.convert_rgb_to_yuv(rng.uniform(0.02, 1.0, (1000, 3)), np.float32).astype(np.float32)
    out, grad, ra, ca = handle.relaxed_emd(_t(a, cuda_device), _t(b, cuda_device), dist, True, True)
    l64, g64, info = O.relaxed_emd(a, b, dist, np.float64, True)
    out = out.cpu().numpy()
    assert abs(out[0] - l64) / l64 <= 1e-4
    assert bool(out[3]) == info["branch_x"]
    clear_r = info["row_gap"] > 1e-5
    clear_c = info["col_gap"] > 1e-5
    assert np.array_equal(ra.cpu().numpy()[clear_r], info["row_argmin"][clear_r])
    assert np.array_equal(ca.cpu().numpy()[clear_c], info["col_argmin"][clear_c])
    _gcheck(grad, g64, cos_min=0.999)


def _palette_case(name, rng):
    """Colour sets that stress the candidate search of the palette term (pal_min2_kernel): ties, clamps, clipping."""
    M, N = 2300, 2600
    a = rng.uniform(0.02, 1.0, (M, 3)); b = rng.uniform(0.02, 1.0, (N, 3))
    if name == "uniform":
        pass
    elif name == "duplicates":             # 60 % of the targets share one colour (flat background)
        a[: int(0.6 * M)] = a[0]
    elif name == "clustered":              # tight cluster plus a few outliers that stretch the bounding box
        a = 0.5 + 0.002 * rng.standard_normal((M, 3)); a[:5] = rng.uniform(0, 1, (5, 3))
        b = 0.5 + 0.01 * rng.standard_normal((N, 3))
    elif name == "disjoint":               # predictions far outside the targets' bounding box
        b = rng.uniform(1.5, 3.0, (N, 3)); b[:100] = rng.uniform(-2.0, -1.0, (100, 3))
    elif name == "scaled":                 # 0..255 colours, negative values
        a = a * 255.0 - 40.0; b = b * 255.0 - 40.0
    elif name == "flat_axis":              # grey images: U = V = 0 for every colour
        a = np.repeat(rng.uniform(0.02, 1.0, (M, 1)), 3, axis=1); b = np.repeat(rng.uniform(0.02, 1.0, (N, 1)), 3, axis=1)
    elif name == "quantised":              # 8-bit style colours: many exact ties
        a = np.round(a * 15) / 15; b = np.round(b * 15) / 15
    return a, b


@pytest.mark.parametrize("dist", ["both", "l2"])
@pytest.mark.parametrize("case", ["uniform", "duplicates", "clustered", "disjoint", "scaled", "flat_axis", "quantised"])
def test_palette_search_is_exact(handle, cuda_device, case, dist):
    """The one-pass search (sqrt.approx, pre-scaled records, redux.sync column minima) must return the minimum cost of
    every row / column (fp64 oracle) and the oracle's argmin wherever it is unique by a clear margin; on exact ties the
    LOWEST index."""
    rng = np.random.default_rng(11)
    a, b = _palette_case(case, rng)
    if dist == "both" or case != "scaled":
        a = O.convert_rgb_to_yuv(a, np.float32)
        b = O.convert_rgb_to_yuv(b, np.float32)
    a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
    out, grad, ra, ca = handle.relaxed_emd(_t(a, cuda_device), _t(b, cuda_device), dist, True, True)
    l64, g64, info = O.relaxed_emd(a, b, dist, np.float64, True)
    out = out.cpu().numpy()
    # |a - b|^2 by expansion (nn/losses.py:19-22, the reference's own formula) cancels in fp32 when colours nearly coincide:
    # the tight cluster is held to the north-star's 1e-3, everything else to 1e-4
    tol = LOSS_RTOL if case == "clustered" else 1e-4
    assert abs(out[0] - l64) / l64 <= tol
    assert abs(out[1] - info["R_X"]) <= tol * info["R_X"] + 1e-7 and abs(out[2] - info["R_Y"]) <= tol * info["R_Y"] + 1e-7
    C = O._dist_fwd(a, b, dist, np.float64)[0]
    ra, ca = ra.cpu().numpy(), ca.cpu().numpy()
    scale = max(1.0, float(np.abs(C).max()))
    # the chosen candidates attain the minimum to the fp32 accuracy of the cost: ~ulp(|a|^2) / (2 sqrt(m)), i.e. up to ~1e-4
    # next to the clamp floor sqrt(1e-6 / 3) where near-identical colours sit (their computed costs tie exactly)
    att = 2e-4 if case in ("clustered", "flat_axis") else 2e-5
    assert np.all(C[np.arange(C.shape[0]), ra] - C.min(axis=1) <= att * scale)
    assert np.all(C[ca, np.arange(C.shape[1])] - C.min(axis=0) <= att * scale)
    # ... and are the oracle's own choice wherever that is unique by a clear margin
    clear_r = info["row_gap"] > 10 * att * scale
    clear_c = info["col_gap"] > 10 * att * scale
    assert np.array_equal(ra[clear_r], info["row_argmin"][clear_r])
    assert np.array_equal(ca[clear_c], info["col_argmin"][clear_c])
    if case == "quantised":
        # exact duplicates of a colour: the lowest index wins, as in the exhaustive kernel
        _, first = np.unique(a, axis=0, return_index=True)
        firsts = set(first.tolist())
        assert all(int(i) in firsts for i in ca)


def test_palette_identical_inputs_hit_clamp_floor(handle, cuda_device):
    rng = np.random.default_rng(4)
    a = rng.uniform(0.1, 1.0, (300, 3)).astype(np.float32)
    out = handle.relaxed_emd(_t(a, cuda_device), _t(a, cuda_device), "both", False)[0].cpu().numpy()
    assert out[0] == pytest.approx(np.sqrt(1e-6 / 3), rel=1e-3)      # nn/losses.py:23 clamp floor


def test_distance_errors(S, handle, cuda_device):
    x = torch.rand(8, 16, device=cuda_device)
    with pytest.raises(KeyError):
        S.relaxed_emd(x, x, distance="manhattan")
    with pytest.raises(NotImplementedError):
        S.relaxed_emd(x, x, distance="l2")
    with pytest.raises(ValueError):
        handle.relaxed_emd(torch.rand(0, 16, device=cuda_device), x, "cosine", False)


def test_convert_rgb_to_yuv(S, cuda_device):
    x = torch.rand(257, 11, device=cuda_device)
    got = S.convert_rgb_to_yuv(x).cpu().numpy()
    assert np.abs(got - O.convert_rgb_to_yuv(x.cpu().numpy(), np.float64)).max() <= 1e-6


# ------------------------------------------------------------------------------ moment matching
@pytest.mark.parametrize("N,M,D,seed", [(600, 500, 2179, 2), (130, 1024, 2179, 6), (1000, 64, 515, 7)])
def test_moment_matching(handle, cuda_device, N, M, D, seed):
    st, co, pr = O.synth_problem(N, M, D, eps=1.0, seed=seed)
    out, grad = handle.moment_matching(_t(st, cuda_device), _t(pr, cuda_device), True)
    l64, g64, info = O.moment_matching(st, pr, np.float64, True)
    out = out.cpu().numpy()
    assert abs(out[0] - l64) / l64 <= LOSS_RTOL
    assert abs(out[1] - info["l_cov"]) / info["l_cov"] <= LOSS_RTOL
    assert abs(out[2] - info["l_mean"]) / info["l_mean"] <= LOSS_RTOL
    _gcheck(grad, g64, cos_min=0.999)
    same = handle.moment_matching(_t(st, cuda_device), _t(st, cuda_device), False)[0].cpu().numpy()
    assert same[0] <= 1e-6 * l64


# ------------------------------------------------------------------------------ self-similarity
@pytest.mark.parametrize("N,eps,seed", [(640, 1.0, 3), (640, 0.1, 3), (640, 0.01, 3), (130, 0.1, 8), (1025, 0.1, 9)])
def test_self_similarity(handle, cuda_device, N, eps, seed):
    st, co, pr = O.synth_problem(N, 8, 2179, eps=eps, seed=seed)
    out, grad = handle.self_similarity(_t(pr, cuda_device), _t(co, cuda_device), True)
    l64, g64, _ = O.self_similarity(pr, co, np.float64, True)
    assert abs(out.item() - l64) / l64 <= LOSS_RTOL
    # sign flips of near-zero L1 terms perturb the direction by ~1e-2 (DESIGN.md "precision")
    _gcheck(grad, g64, cos_min=0.999)
    assert float((grad * _t(pr, cuda_device)).sum(dim=1).abs().max()) <= 1e-3 * float(grad.norm())
    same = handle.self_similarity(_t(co, cuda_device), _t(co, cuda_device), False)[0].item()
    assert abs(same) <= 1e-6 * max(l64, 1e-4) * 100


@pytest.mark.parametrize("N,D,eps,seed", [(4500, 515, 0.1, 31), (2300, 2179, 0.1, 32), (4096, 259, 1.0, 33), (2049, 131, 0.1, 34),
                                          (4700, 67, 0.1, 35)])
def test_self_similarity_multi_panel(handle, cuda_device, N, D, eps, seed):
    """N > 2048: several row panels; exercises the symmetric path (mirror accounting, transposed stage 2b)."""
    st, co, pr = O.synth_problem(N, 8, D, eps=eps, seed=seed)
    out, grad = handle.self_similarity(_t(pr, cuda_device), _t(co, cuda_device), True)
    l64, g64, _ = O.self_similarity(pr, co, np.float64, True)
    assert abs(out.item() - l64) / l64 <= LOSS_RTOL
    _gcheck(grad, g64, cos_min=0.999)
    # every row block must carry the right gradient, not just the total
    g = grad.double().cpu().numpy()
    for lo in range(0, N, 1024):
        ref = g64[lo:lo + 1024]
        got = g[lo:lo + 1024]
        assert abs(np.linalg.norm(got) - np.linalg.norm(ref)) / np.linalg.norm(ref) <= GRADNORM_RTOL
        assert float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref))) >= 0.999


# ------------------------------------------------------------------------------ fused evaluation
@pytest.mark.parametrize("name", ["small_d67", "ragged_d2179", "default_d2179_eps1", "near_d2179_eps001"])
def test_total_against_golden(S, cuda_device, name):
    from strotss_tensorflow_b200 import _lib
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    N, M, D, alpha = int(z["N"]), int(z["M"]), int(z["D"]), float(z["alpha"])
    st, co, pr = O.synth_problem(N, M, D, eps=float(z["eps"]), seed=int(z["seed"]))
    mod = S.StrotssLoss(_t(st, cuda_device), alpha)
    sc, grad, ra, ca = mod.handle.eval(_t(pr, cuda_device), _t(co, cuda_device), alpha, True, True)
    s = sc.cpu().numpy()
    for slot, key in [(_lib.S_TOTAL, "total"), (_lib.S_LOSS_C, "loss_c"), (_lib.S_LOSS_S, "loss_s"), (_lib.S_L_M, "l_m"),
                      (_lib.S_L_REMD, "l_remd"), (_lib.S_L_PALETTE, "l_palette")]:
        assert abs(s[slot] - float(z[key])) / abs(float(z[key])) <= LOSS_RTOL, key
    clear = z["remd_row_gap"] > GAP_THR
    assert np.array_equal(ra.cpu().numpy()[clear], z["remd_row_argmin"][clear])
    clear = z["remd_col_gap"] > GAP_THR
    assert np.array_equal(ca.cpu().numpy()[clear], z["remd_col_argmin"][clear])
    g = grad.double().cpu().numpy()
    assert abs(np.linalg.norm(g) - float(z["grad_norm"])) / float(z["grad_norm"]) <= GRADNORM_RTOL
    if "grad" in z.files:
        ref = z["grad"]
        assert float((g * ref).sum() / (np.linalg.norm(g) * np.linalg.norm(ref))) >= 0.99


@pytest.mark.parametrize("name", ["small_d67", "ragged_d2179", "default_d2179_eps1", "near_d2179_eps001"])
def test_total_against_reference_code_golden(S, cuda_device, name):
    """The same four problems against tests/golden/ref_*.npz: numbers produced by the REFERENCE'S OWN loss code
    (nn/losses.py, StyleLoss / ContentLoss of run_strotss.py executed over a stand-in for the TensorFlow ops they call;
    tests/golden/make_reference_golden.py), gradient by autograd through the reference's op sequence."""
    from strotss_tensorflow_b200 import _lib
    z = np.load(os.path.join(GOLDEN, "ref_" + name + ".npz"))
    N, M, D, alpha = int(z["N"]), int(z["M"]), int(z["D"]), float(z["alpha"])
    st, co, pr = O.synth_problem(N, M, D, eps=float(z["eps"]), seed=int(z["seed"]))
    mod = S.StrotssLoss(_t(st, cuda_device), alpha)
    sc, grad, _, _ = mod.handle.eval(_t(pr, cuda_device), _t(co, cuda_device), alpha, True, True)
    s = sc.cpu().numpy()
    for slot, key in [(_lib.S_TOTAL, "total"), (_lib.S_LOSS_C, "loss_c"), (_lib.S_LOSS_S, "loss_s"), (_lib.S_L_M, "l_m"),
                      (_lib.S_L_REMD, "l_remd"), (_lib.S_L_PALETTE, "l_palette")]:
        assert abs(s[slot] - float(z[key])) / abs(float(z[key])) <= LOSS_RTOL, key
    g = grad.double().cpu().numpy()
    assert abs(np.linalg.norm(g) - float(z["grad_norm"])) / float(z["grad_norm"]) <= GRADNORM_RTOL
    if "grad" in z.files:
        ref = z["grad"]
        assert float((g * ref).sum() / (np.linalg.norm(g) * np.linalg.norm(ref))) >= 0.99


@pytest.mark.parametrize("alpha", [16.0, 0.5])
def test_modules_match_reference_wrappers(S, cuda_device, alpha):
    """ContentLoss / StyleLoss keep run_strotss.py:21-40 semantics; autograd delivers d loss / d pred."""
    st, co, pr = O.synth_problem(384, 300, 2179, eps=0.1, seed=12)
    style = S.StyleLoss(_t(st, cuda_device), alpha)
    content = S.ContentLoss()
    pred = _t(pr, cuda_device).requires_grad_(True)
    loss_c = content(_t(co, cuda_device), pred)
    loss_s = style(pred)
    denom = 2.0 + alpha + 1.0 / max(alpha, 1.0)
    loss = (alpha * loss_c + loss_s) / denom
    loss.backward()
    ref, gref, info = O.total_loss(st, co, pr, alpha, np.float64, True)
    assert abs(loss.item() - ref) / ref <= LOSS_RTOL
    assert abs(loss_c.item() - info["loss_c"]) / info["loss_c"] <= LOSS_RTOL
    assert abs(loss_s.item() - info["loss_s"]) / info["loss_s"] <= LOSS_RTOL
    _gcheck(pred.grad, gref, cos_min=0.99)
    fused = S.StrotssLoss(_t(st, cuda_device), alpha)
    p2 = _t(pr, cuda_device).requires_grad_(True)
    l2 = fused(_t(co, cuda_device), p2)
    l2.backward()
    assert abs(l2.item() - loss.item()) <= 1e-5 * abs(loss.item())
    # the fused call rounds delta = x^ - y^ in a different kernel than the stand-alone self_similarity: a few
    # near-zero L1 terms change sign, which moves the gradient by less than either path's distance to the oracle
    assert float((p2.grad - pred.grad).norm() / pred.grad.norm()) <= 1e-2
    _gcheck(p2.grad, gref, cos_min=0.99)


def test_row_strides_are_honoured(S, cuda_device):
    """Inputs may be column slices of wider matrices (row stride > D): the C ABI takes `ld` and must give the results of
    the packed copies.  Packed, 16-byte aligned rows take the streaming row pass (bulk copies), strided ones the general
    row pass: same formulas, different summation order of the column sums, so the comparison is to fp32 rounding, the
    argmins exactly."""
    st, co, pr = O.synth_problem(333, 200, 67, eps=0.1, seed=77)
    def wide(a, pad):
        buf = torch.full((a.shape[0], a.shape[1] + pad), 7.5, device=cuda_device, dtype=torch.float32)
        buf[:, :a.shape[1]] = _t(a, cuda_device)
        return buf[:, :a.shape[1]]
    h1, h2 = S.Handle(cuda_device), S.Handle(cuda_device)
    h1.set_style_target(_t(st, cuda_device))
    sw = wide(st, 5)
    assert not sw.is_contiguous()
    h2.set_style_target(sw)
    s1, g1, ra1, ca1 = h1.eval(_t(pr, cuda_device), _t(co, cuda_device), 4.0, True, True)
    s2, g2, ra2, ca2 = h2.eval(wide(pr, 13), wide(co, 3), 4.0, True, True)
    assert torch.allclose(s1[:12], s2[:12], rtol=2e-5, atol=1e-8) and torch.equal(ra1, ra2) and torch.equal(ca1, ca2)
    assert float((g1 - g2).norm() / g1.norm()) <= 2e-3            # sign flips of L1 terms that are ~0 to within fp32 rounding
    # two packed evaluations (the same path twice) are bit-identical in every scalar
    s3, g3, _, _ = h1.eval(_t(pr, cuda_device), _t(co, cuda_device), 4.0, True, True)
    assert torch.equal(s1[:12], s3[:12])


def test_evaluation_is_cuda_graph_capturable(S, cuda_device):
    """After one warm-up call at a given size no entry point allocates or synchronises, and the branch streams fork from /
    join to the caller's stream: an evaluation can be captured into the caller's CUDA graph and replayed on new data."""
    st, co, pr = O.synth_problem(384, 300, 2179, eps=0.1, seed=91)
    h = S.Handle(cuda_device)
    h.set_style_target(_t(st, cuda_device))
    pred, content = _t(pr, cuda_device), _t(co, cuda_device)
    want, gwant, _, _ = h.eval(pred, content, 8.0, True)                   # warm-up: workspace growth
    want, gwant = want.clone(), gwant.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        sc, grad, _, _ = h.eval(pred, content, 8.0, True)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(sc[:12], want[:12])
    assert torch.allclose(grad, gwant, rtol=1e-5, atol=1e-10)
    # new data in the captured buffers
    _, co2, pr2 = O.synth_problem(384, 300, 2179, eps=0.3, seed=92)
    pred.copy_(_t(pr2, cuda_device)); content.copy_(_t(co2, cuda_device))
    graph.replay()
    torch.cuda.synchronize()
    ref = O.total_loss(st, co2, pr2, 8.0, np.float64)
    assert abs(sc[0].item() - ref) / ref <= LOSS_RTOL


def test_per_function_entry_points_are_cuda_graph_capturable(S, cuda_device):
    """strotss_relaxed_emd / strotss_moment_matching / strotss_self_similarity stage nothing through the host (the scalar
    slots they copy out travel as kernel arguments), so a replayed capture gives the answers of a direct call -- also after
    the capturing call's stack frame is long gone and on new data."""
    st, co, pr = O.synth_problem(300, 260, 2179, eps=0.2, seed=93)
    h = S.Handle(cuda_device)
    x, y, c = _t(st, cuda_device), _t(pr, cuda_device), _t(co, cuda_device)
    for fn in (lambda: h.relaxed_emd(x, y, "cosine", True)[:2], lambda: h.relaxed_emd(x[:, :3].contiguous(), y[:, :3].contiguous(), "both", True)[:2],
               lambda: h.moment_matching(x, y, True), lambda: h.self_similarity(y, c, True)):
        want, gwant = fn()                                                 # warm-up: workspace growth
        want, gwant = want.clone(), gwant.clone()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out, grad = fn()
        _ = [np.zeros(64) for _ in range(8)]                                # churn the host stack / heap between capture and replay
        out.zero_(); grad.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want)
        assert torch.allclose(grad, gwant, rtol=1e-5, atol=1e-10)


def test_wide_features_take_the_general_preparation_path(S, cuda_device):
    """D > 2560 exceeds the fused row pass (10 columns per thread) and must fall back, not fail."""
    st, co, pr = O.synth_problem(200, 180, 2700, eps=0.1, seed=51)
    mod = S.StrotssLoss(_t(st, cuda_device), 4.0)
    sc, grad, _, _ = mod.handle.eval(_t(pr, cuda_device), _t(co, cuda_device), 4.0, True)
    ref, gref, _ = O.total_loss(st, co, pr, 4.0, np.float64, True)
    assert abs(sc[0].item() - ref) / ref <= LOSS_RTOL
    _gcheck(grad, gref, cos_min=0.99)


def test_region_sizes_of_masked_mode(S, cuda_device):
    """Config 3 (masked transfer): R independent ragged problems (N_r, M_r), averaged (run_strotss.py:114-121)."""
    total, ref_total = 0.0, 0.0
    for r, (N, M) in enumerate([(1024, 1024), (700, 1024), (333, 517)]):
        st, co, pr = O.synth_problem(N, M, 2179, eps=0.1, seed=20 + r)
        mod = S.StrotssLoss(_t(st, cuda_device), 16.0)
        sc, _, _, _ = mod.handle.eval(_t(pr, cuda_device), _t(co, cuda_device), 16.0, False)
        total += sc[0].item() / 3
        ref_total += O.total_loss(st, co, pr, 16.0, np.float64) / 3
    assert abs(total - ref_total) / ref_total <= LOSS_RTOL


def test_pipelined_host_evaluation_matches_blocking_call(S, cuda_device):
    """strotss_eval_host_submit/_wait (two evaluations in flight) returns what strotss_eval_host returns."""
    from strotss_tensorflow_b200 import _lib
    probs = [O.synth_problem(300 + 40 * k, 260, 2179, eps=0.1, seed=80 + k) for k in range(5)]
    h = S.Handle(cuda_device)
    h.set_style_target(_t(probs[0][0], cuda_device))
    want = []
    for _, co, pr in probs:
        g = torch.empty(pr.shape, dtype=torch.float32).pin_memory()
        sc = torch.empty(_lib.NUM_SCALARS, dtype=torch.float32)
        h.eval_host(torch.tensor(pr).pin_memory(), torch.tensor(co).pin_memory(), 16.0, g, sc)
        want.append((sc.clone(), g.clone()))
    ins = [(torch.tensor(pr).pin_memory(), torch.tensor(co).pin_memory()) for _, co, pr in probs]
    outs = [(torch.empty(_lib.NUM_SCALARS, dtype=torch.float32), torch.zeros(pr.shape, dtype=torch.float32).pin_memory()) for _, _, pr in probs]
    tickets = []
    for k, ((p, c), (sc, g)) in enumerate(zip(ins, outs)):
        tickets.append(h.eval_host_submit(p, c, 16.0, g, sc))
        if k >= 1:
            h.eval_host_wait(tickets[k - 1])
    h.eval_host_wait(tickets[-1])
    for (sc, g), (wsc, wg) in zip(outs, want):
        assert torch.equal(sc[:6], wsc[:6])
        # the relaxed-EMD scatter branch uses float atomics: allow their reordering noise only
        assert torch.allclose(g, wg, rtol=1e-4, atol=1e-9)
    with pytest.raises(S.runtime._lib.StrotssError):
        h.eval_host_wait(tickets[0])                     # already collected


def test_grouped_masked_evaluation(S, cuda_device):
    """SURVEY 8(f) next #2 / BASELINE config 3: strotss_eval_grouped == the masked train_step loop
    (run_strotss.py:112-124): mean over regions of the per-region totals, gradients carry the 1/R."""
    from strotss_tensorflow_b200 import _lib
    alpha = 8.0
    sizes = [(1024, 1024), (700, 1024), (333, 517), (5, 40)]       # N_r = 1 is degenerate: Xd = [~0] / clamp 1e-12
    probs = [O.synth_problem(N, M, 2179, eps=0.1, seed=40 + r) for r, (N, M) in enumerate(sizes)]
    styles, contents, preds = zip(*probs)
    mod = S.MaskedStrotssLoss([_t(s, cuda_device) for s in styles], alpha)
    pts = [_t(p, cuda_device).requires_grad_(True) for p in preds]
    loss = mod([_t(c, cuda_device) for c in contents], pts)
    loss.backward()
    ref, grefs, info = O.masked_total_loss(styles, contents, preds, alpha, np.float64, True)
    assert abs(loss.item() - ref) / ref <= LOSS_RTOL
    s = mod.last_scalars.cpu().numpy()
    assert abs(s[_lib.S_LOSS_C] - info["loss_c"]) / info["loss_c"] <= LOSS_RTOL
    assert abs(s[_lib.S_LOSS_S] - info["loss_s"]) / info["loss_s"] <= LOSS_RTOL
    for r, (p, g) in enumerate(zip(pts, grefs)):
        _gcheck(p.grad, g, cos_min=0.99)
    # each region agrees with the single-region entry point (same kernels, different stream / workspace)
    reg = mod.last_region_scalars.cpu().numpy()
    for r in range(len(sizes)):
        one = S.StrotssLoss(_t(styles[r], cuda_device), alpha)
        sc, gr, _, _ = one.handle.eval(_t(preds[r], cuda_device), _t(contents[r], cuda_device), alpha, True)
        assert abs(reg[r, _lib.S_TOTAL] - sc[0].item()) <= 1e-6 * abs(sc[0].item())
        assert torch.allclose(pts[r].grad * len(sizes), gr, rtol=1e-5, atol=1e-9)
    # the number of prediction samples per region is dynamic (nn/strotss_utils.py:113): re-evaluate with other sizes
    probs2 = [O.synth_problem(N, M, 2179, eps=0.1, seed=60 + r) for r, (N, M) in enumerate([(900, 1024), (1024, 1024), (64, 517), (300, 40)])]
    l2 = mod([_t(p[1], cuda_device) for p in probs2], [_t(p[2], cuda_device) for p in probs2])
    ref2 = O.masked_total_loss(styles, [p[1] for p in probs2], [p[2] for p in probs2], alpha, np.float64)
    assert abs(l2.item() - ref2) / ref2 <= LOSS_RTOL


def test_grouped_evaluation_threads_and_capture(S, cuda_device):
    """Repeated grouped calls (launching threads after the warm-up call) and a captured grouped call (everything on the
    capturing thread) return what the first, single-threaded call returned."""
    alpha = 4.0
    probs = [O.synth_problem(N, M, 259, eps=0.1, seed=120 + r) for r, (N, M) in enumerate([(300, 200), (129, 64), (40, 70)])]
    h = S.Handle(cuda_device)
    h.set_style_targets_grouped([_t(p[0], cuda_device) for p in probs])
    preds = [_t(p[2], cuda_device) for p in probs]; contents = [_t(p[1], cuda_device) for p in probs]
    s0, r0, g0 = h.eval_grouped(preds, contents, alpha, True)                # warm-up: calling thread only
    for _ in range(5):
        s1, r1, g1 = h.eval_grouped(preds, contents, alpha, True)            # regions 1.. on launching threads
        assert torch.equal(s0[:6], s1[:6]) and torch.equal(r0[:, :6], r1[:, :6])
        assert all(torch.allclose(a, b, rtol=1e-5, atol=1e-10) for a, b in zip(g0, g1))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        s2, r2, g2 = h.eval_grouped(preds, contents, alpha, True)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(s0[:6], s2[:6])
    assert all(torch.allclose(a, b, rtol=1e-5, atol=1e-10) for a, b in zip(g0, g2))
    ref = O.masked_total_loss([p[0] for p in probs], [p[1] for p in probs], [p[2] for p in probs], alpha, np.float64)
    assert abs(s2[0].item() - ref) / ref <= LOSS_RTOL


def test_grouped_argument_errors(S, cuda_device):
    st, co, pr = O.synth_problem(64, 48, 67, eps=0.1, seed=1)
    mod = S.MaskedStrotssLoss([_t(st, cuda_device), _t(st[:20], cuda_device)], 4.0)
    with pytest.raises(ValueError):
        mod([_t(co, cuda_device)], [_t(pr, cuda_device)])                       # wrong number of regions
    with pytest.raises(ValueError):
        mod.handle.eval_grouped([_t(pr, cuda_device), _t(pr[:0], cuda_device)], [_t(co, cuda_device), _t(co[:0], cuda_device)], 4.0)


# ------------------------------------------------------------------------------ hypercolumn sampler (8f next #1)
_SAMPLER_SHAPES = [(85, 128, 3), (85, 128, 64), (85, 128, 64), (42, 64, 128), (42, 64, 128), (21, 32, 256), (21, 32, 256),
                   (21, 32, 256), (10, 16, 512), (10, 16, 512)]                 # content at scale 128


@pytest.mark.parametrize("bilinear", [True, False])
def test_sampler_forward_is_bit_exact(S, cuda_device, bilinear):
    rng = np.random.default_rng(5)
    xs = [rng.standard_normal((1,) + s).astype(np.float32) for s in _SAMPLER_SHAPES]
    idx = np.stack([rng.uniform(-1, 86, 700), rng.uniform(-1, 129, 700)], axis=1).astype(np.float32)   # incl. out-of-range
    ref = O.sample_hypercolumns(xs, idx, bilinear)
    samp = S.Sampling(1024)
    got = samp._sample([_t(x, cuda_device) for x in xs], _t(idx, cuda_device), bilinear).cpu().numpy()
    assert got.shape == (700, 2179)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("mode", ["bilinear", "nearest"])
def test_sampler_against_reference_code_golden(S, cuda_device, mode):
    """tests/golden/ref_sampler.npz: Sampling._sample executed from the reference's own source (fp32) on ten small maps --
    the kernel must reproduce it bit for bit."""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_reference_golden as G
    z = np.load(os.path.join(GOLDEN, "ref_sampler.npz"))
    xs, idx = G.sampler_inputs()
    assert np.array_equal(idx, z["indices"])
    got = S.Sampling(G.SAMPLER_N)._sample([_t(x, cuda_device) for x in xs], _t(idx, cuda_device), mode == "bilinear").cpu().numpy()
    assert got.shape == z[mode].shape
    assert np.array_equal(got, z[mode])


def test_sampler_backward_matches_autograd(S, cuda_device):
    rng = np.random.default_rng(6)
    shapes = [(20, 24, 3), (20, 24, 16), (10, 12, 32), (5, 6, 40)]
    xs_np = [rng.standard_normal((1,) + s).astype(np.float32) for s in shapes]
    idx = np.stack([rng.uniform(0, 20, 300), rng.uniform(0, 24, 300)], axis=1).astype(np.float32)
    gout = rng.standard_normal((300, sum(s[2] for s in shapes))).astype(np.float32)
    xs = [_t(x, cuda_device).requires_grad_(True) for x in xs_np]
    out = S.Sampling(300)._sample(xs, _t(idx, cuda_device), True)
    out.backward(_t(gout, cuda_device))
    # reference: the same gather written with torch indexing on the CPU in fp64
    xr = [torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in xs_np]
    divs = O.sampler_scales(shapes)
    cur = torch.tensor(idx, dtype=torch.float32)
    feats = []
    for x, d, (h, w, c) in zip(xr, divs, shapes):
        if d != 1.0:
            cur = cur / np.float32(d)
        gx, gy = cur[:, 0], cur[:, 1]
        gxf, gyf = gx.floor(), gy.floor()
        dx, dy = (gx - gxf).double()[:, None], (gy - gyf).double()[:, None]
        xi = gxf.clamp(0, h - 1).long(); yi = gyf.clamp(0, w - 1).long()
        xb = (xi + 1).clamp(0, h - 1); yb = (yi + 1).clamp(0, w - 1)
        flat = x.reshape(h * w, c)
        feats.append(flat[xi * w + yi] * (1 - dx) * (1 - dy) + flat[xi * w + yb] * (1 - dx) * dy
                     + flat[xb * w + yi] * dx * (1 - dy) + flat[xb * w + yb] * dx * dy)
    torch.cat(feats, dim=1).backward(torch.tensor(gout, dtype=torch.float64))
    for a, b in zip(xs, xr):
        assert float((a.grad.double().cpu() - b.grad).abs().max()) <= 1e-4 * float(b.grad.abs().max())


def test_sampler_gradient_reaches_non_contiguous_maps(S, cuda_device):
    """A permuted (NCHW -> NHWC view) feature map is copied inside the sampler; its gradient must still arrive."""
    g = torch.Generator(device=cuda_device).manual_seed(7)
    nchw = [torch.rand(1, c, hh, ww, generator=g, device=cuda_device, requires_grad=True) for (hh, ww, c) in [(24, 32, 3), (12, 16, 8)]]
    views = [t.permute(0, 2, 3, 1) for t in nchw]
    assert not views[1].is_contiguous()
    samp = S.Sampling(64, torch.Generator().manual_seed(0))
    idx = samp._make_indices(views[0], True)
    out = samp._sample(views, idx, True)
    wgt = torch.rand(out.shape, generator=g, device=cuda_device)
    (out * wgt).sum().backward()
    ref_in = [t.detach().clone().requires_grad_(True) for t in nchw]
    ref = samp._sample([t.permute(0, 2, 3, 1).contiguous() for t in ref_in], idx, True)
    (ref * wgt).sum().backward()
    for a, b in zip(nchw, ref_in):
        assert a.grad is not None and float(a.grad.abs().sum()) > 0
        assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7)


def test_sampler_feeds_the_loss_path(S, cuda_device):
    """train_step shape of the path (run_strotss.py:134-141): sample content/pred maps at the same points,
    evaluate the fused loss, and carry the gradient back to the prediction's feature maps."""
    g = torch.Generator(device=cuda_device).manual_seed(0)
    content_maps = [torch.rand((1,) + s, generator=g, device=cuda_device) for s in _SAMPLER_SHAPES]
    pred_maps = [(m + 0.05 * torch.rand(m.shape, generator=g, device=cuda_device)).requires_grad_(True) for m in content_maps]
    style_maps = [torch.rand((1,) + s, generator=g, device=cuda_device) for s in _SAMPLER_SHAPES]
    samp = S.Sampling(1024, torch.Generator().manual_seed(0))
    loss_fn = S.StrotssLoss(samp(style_maps), 16.0)
    c_feat, p_feat = samp.bilinear(content_maps, pred_maps)
    assert c_feat.shape == (1024, 2179) and p_feat.shape == (1024, 2179)
    loss = loss_fn(c_feat, p_feat)
    loss.backward()
    assert torch.isfinite(loss) and all(m.grad is not None and bool(torch.isfinite(m.grad).all()) for m in pred_maps)
    assert float(pred_maps[0].grad.abs().sum()) > 0 and float(pred_maps[-1].grad.abs().sum()) > 0
    # parity of this evaluation against the oracle on the very same sampled features
    st = loss_fn.style_features.cpu().numpy()
    want = O.total_loss(st, c_feat.detach().cpu().numpy(), p_feat.detach().cpu().numpy(), 16.0)
    assert abs(loss.item() - want) / want <= LOSS_RTOL


# ------------------------------------------------------------------------------ alternative kernel paths
_ALT_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from oracle import strotss_oracle as O
import strotss_tensorflow_b200 as S
dev = torch.device('cuda', 0)
st, co, pr = O.synth_problem(2300, 2100, 2179, eps=0.1, seed=41)   # sizes at which tile couples pay (two rounds of plain tiles)
mod = S.StrotssLoss(torch.tensor(st, device=dev), 16.0)
sc, grad, _, _ = mod.handle.eval(torch.tensor(pr, device=dev), torch.tensor(co, device=dev), 16.0, True)
ref, gref, info = O.total_loss(st, co, pr, 16.0, np.float64, True)
g = grad.double().cpu().numpy()
print('REL', abs(sc[0].item() - ref) / ref, abs(np.linalg.norm(g) - np.linalg.norm(gref)) / np.linalg.norm(gref),
      float((g * gref).sum() / (np.linalg.norm(g) * np.linalg.norm(gref))))
"""


@pytest.mark.parametrize("env", [{"STROTSS_NO_PAIR": "1"}, {"STROTSS_SS1_GENERIC": "1"},
                                 {"STROTSS_NO_PAIR": "1", "STROTSS_SS1_GENERIC": "1"},
                                 {"STROTSS_BRANCHES": "0"}, {"STROTSS_PAL_TWO_PASS": "1"},
                                 {"STROTSS_BRANCHES": "0", "STROTSS_OVERLAP": "1", "STROTSS_PANEL": "1024"}, {"STROTSS_NO_TRAP": "1", "STROTSS_PANEL": "2048"},
                                 {"STROTSS_PANEL": "1024"}, {"STROTSS_FINALIZE_GENERIC": "1", "STROTSS_NO_KTAIL": "1", "STROTSS_PREP_V1": "1"},
                                 {"STROTSS_WIDE": "0"}, {"STROTSS_WIDE": "0", "STROTSS_PANEL": "1024"},
                                 {"STROTSS_SS1_MERGED": "0", "STROTSS_REMD_SKEW": "-1"}, {"STROTSS_SS1_TAIL": "0", "STROTSS_REMD_SKEW": "0"},
                                 {"STROTSS_SS1_TAIL": "40", "STROTSS_REMD_SKEW": "40"}, {"STROTSS_PREP_V2": "1"},
                                 {"STROTSS_PREP_V2": "1", "STROTSS_V_FP32": "1", "STROTSS_WIDE": "0"}, {"STROTSS_WIDE": "0", "STROTSS_NO_TRAP": "1"}, {"STROTSS_PDL": "1"},
                                 {"STROTSS_PDL": "1", "STROTSS_BRANCHES": "0", "STROTSS_PREP_V2": "1"}])
def test_alternative_kernel_paths(cuda_device, env):
    """The single-CTA GEMM kernels (STROTSS_NO_PAIR), the generic stage-1 epilogue (STROTSS_SS1_GENERIC), the
    single-stream launch order (STROTSS_BRANCHES=0), the two-pass palette search and the two-stream stage-1/stage-2
    pipelining (STROTSS_OVERLAP=1) and the rectangular panel walk (STROTSS_NO_TRAP) stay selectable for A/B measurements; they must give the same answers.  The
    switches are read once per process."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", _ALT_SCRIPT % root], env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    rel = [float(v) for v in out.stdout.strip().splitlines()[-1].split()[1:]]
    assert rel[0] <= LOSS_RTOL and rel[1] <= GRADNORM_RTOL and rel[2] >= 0.999


# ------------------------------------------------------------------------------ full-size properties
def test_full_size_properties(S, cuda_device):
    """N = M = 16384, D = 2179 (BASELINE.json configs[3]) through size-independent properties: the oracle
    cannot run this size in seconds, so check invariances the math guarantees."""
    import bench
    from strotss_tensorflow_b200 import _lib
    N = M = 16384
    style, content, pred = bench.synth_torch(N, M, 2179, 0.1, 0, cuda_device)
    h = S.Handle(cuda_device)
    h.set_style_target(style)
    sc, grad, ra, ca = h.eval(pred, content, 16.0, True, True)
    s = sc.cpu().numpy()
    assert np.all(np.isfinite(s)) and bool(torch.isfinite(grad).all())
    assert s[_lib.S_TOTAL] == pytest.approx((16.0 * s[_lib.S_LOSS_C] + s[_lib.S_LOSS_S]) / 18.0625, rel=1e-5)
    assert s[_lib.S_L_REMD] == pytest.approx(max(s[_lib.S_REMD_RX], s[_lib.S_REMD_RY]), rel=1e-6)
    ra = ra.long(); ca = ca.long()
    assert int(ra.min()) >= 0 and int(ra.max()) < N and int(ca.min()) >= 0 and int(ca.max()) < M
    # the reported minima are attained at the reported indices (fp32 recomputation of the chosen pairs)
    xs = torch.nn.functional.normalize(style, dim=1); yp = torch.nn.functional.normalize(pred, dim=1)
    rx = (1 - (xs * yp[ra]).sum(dim=1)).mean().item()
    ry = (1 - (yp * xs[ca]).sum(dim=1)).mean().item()
    assert rx == pytest.approx(s[_lib.S_REMD_RX], rel=1e-3) and ry == pytest.approx(s[_lib.S_REMD_RY], rel=1e-3)
    # permuting the samples (content and prediction together) leaves every loss unchanged and permutes the gradient
    perm = torch.randperm(N, device=cuda_device, generator=torch.Generator(device=cuda_device).manual_seed(1))
    sc2, grad2, _, _ = h.eval(pred[perm].contiguous(), content[perm].contiguous(), 16.0, True)
    s2 = sc2.cpu().numpy()
    assert np.allclose(s2[:6], s[:6], rtol=2e-4)
    assert float((grad2 - grad[perm]).norm() / grad.norm()) <= 3e-2
    # run-to-run: losses are bit-identical (fixed-order reductions)
    sc3, _, _, _ = h.eval(pred, content, 16.0, True)
    assert np.array_equal(sc3.cpu().numpy()[:12], s[:12])
    # self-similarity of a set with itself vanishes, so loss_c -> 0 and only the style terms remain
    sc4, _, _, _ = h.eval(content, content, 16.0, False)
    assert abs(sc4[_lib.S_LOSS_C].item()) <= 1e-6


@pytest.mark.parametrize("eps", [1.0, 0.1, 0.01])
def test_full_size_against_the_fp64_restatement(S, cuda_device, eps):
    """N = M = 16384, D = 2179 (the bench workload, BASELINE.json configs[3]) against the fp64 restatement of the
    reference op sequence (oracle/torch_port.py: materialised matrices + autograd), evaluated on the same device in
    float64 -- the one place where the full size can be checked value by value rather than through invariants.
    eps = 1 is the bench's decorrelated setting; 0.1 and 0.01 are the near-content regime the optimiser lives in and
    the precision stress of the bf16 operands."""
    import bench
    from oracle import torch_port as T
    from strotss_tensorflow_b200 import _lib
    N = M = 16384
    style, content, pred = bench.synth_torch(N, M, 2179, eps, 0, cuda_device)
    h = S.Handle(cuda_device)
    h.set_style_target(style)
    sc, grad, _, _ = h.eval(pred, content, 16.0, True)
    s = sc.double().cpu().numpy()
    g = grad.double()
    del h
    ref, gref, info = T.total_loss_and_grad(style.double(), content.double(), pred.double(), 16.0)
    assert abs(s[_lib.S_TOTAL] - ref.item()) / ref.item() <= LOSS_RTOL
    for slot, key in [(_lib.S_LOSS_C, "loss_c"), (_lib.S_LOSS_S, "loss_s"), (_lib.S_L_M, "l_m"), (_lib.S_L_REMD, "l_remd"),
                      (_lib.S_L_PALETTE, "l_palette")]:
        assert abs(s[slot] - info[key].item()) / info[key].item() <= LOSS_RTOL, key
    gn, rn = g.norm().item(), gref.norm().item()
    assert abs(gn - rn) / rn <= GRADNORM_RTOL
    assert (g * gref).sum().item() / (gn * rn) >= 0.99
    del gref, g
    torch.cuda.empty_cache()
