"""strotss_tensorflow_b200/tf_adapter.py (the tf.custom_gradient binding of the C ABI) EXECUTED on the B200 against
tests/tf_standin.py -- TensorFlow itself is not installable here -- and checked against

  * the oracle on a seeded problem (every function: value and gradient through tf.GradientTape), and
  * tests/golden/ref_train_step.npz: the scalars and feature gradients the REFERENCE'S OWN train_step functions
    (run_strotss.py:104-125, 131-142, exec'd by tests/golden/make_reference_golden.py) produced on the recorded sampled
    features.  The loss lines of train_step are restated here (:114-124, :138-141) over the adapter's StyleLoss /
    ContentLoss / StrotssLoss; reference source cannot travel to the GPU box.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import strotss_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_RTOL, GRADNORM_RTOL = 1e-3, 1e-2
# ref_train_step.npz holds 13-channel toy features (RGB + two maps of a toy extractor) that are nearly collinear: their cosine
# distances are ~1e-2, so the 2^-9 rounding of the bf16 operands is ~1e-2 RELATIVE on the relaxed-EMD term (on the 2179-channel
# hypercolumns of the other fixtures the same rounding stays below 1e-4).  The style term of this fixture is therefore held
# to 2e-2; the total and the content term (which dominate the step and its gradient) to the usual 1e-3.
TOY_STYLE_RTOL = 2e-2


@pytest.fixture(scope="module")
def tfa(cuda_device):
    import tf_standin
    saved = sys.modules.get("tensorflow")
    sys.modules["tensorflow"] = tf_standin.make_module()
    try:
        import strotss_tensorflow_b200.tf_adapter as mod
        mod = importlib.reload(mod)                 # bind the adapter to the stand-in
        assert mod._HAVE_TF
        yield mod
        mod._Handles.close()
    finally:
        if saved is None:
            sys.modules.pop("tensorflow", None)
        else:
            sys.modules["tensorflow"] = saved
        import strotss_tensorflow_b200.tf_adapter as mod2
        importlib.reload(mod2)


def _var(tfa, a, dev):
    return tfa.tf.Tensor(torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev).requires_grad_(True))


def _const(tfa, a, dev):
    return tfa.tf.Tensor(torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev))


def _close(got, want, rtol=LOSS_RTOL):
    assert abs(float(got) - float(want)) <= rtol * abs(float(want)), (float(got), float(want))


def _gclose(g, ref, cos_min=0.995):
    g = g.t.double().cpu().numpy()
    nr = np.linalg.norm(ref)
    assert abs(np.linalg.norm(g) - nr) / nr <= GRADNORM_RTOL
    assert float((g * ref).sum() / (np.linalg.norm(g) * nr)) >= cos_min


def test_functions_against_the_oracle(tfa, cuda_device):
    tf = tfa.tf
    st, co, pr = O.synth_problem(300, 260, 2179, eps=0.1, seed=3)
    x, c = _const(tfa, st, cuda_device), _const(tfa, co, cuda_device)
    for name, call, ref in [
        ("relaxed_emd", lambda y: tfa.relaxed_emd(x, y), lambda: O.relaxed_emd(st, pr, "cosine", np.float64, True)),
        ("moment_matching", lambda y: tfa.moment_matching(x, y), lambda: O.moment_matching(st, pr, np.float64, True)),
        ("self_similarity", lambda y: tfa.self_similarity(y, c), lambda: O.self_similarity(pr, co, np.float64, True)),
    ]:
        y = _var(tfa, pr, cuda_device)
        with tf.GradientTape() as tape:
            loss = 3.0 * call(y)                     # an upstream factor: the custom gradient must scale with it
        g = tape.gradient(loss, y)
        want, gwant = ref()[:2]
        _close(loss, 3.0 * want)
        _gclose(g, 3.0 * np.asarray(gwant), cos_min=0.99)
    # the palette call of StyleLoss (run_strotss.py:37-39): 'both' on 3 YUV channels, gradient back through rgb_to_yuv
    y = _var(tfa, pr, cuda_device)
    with tf.GradientTape() as tape:
        lp = tfa.relaxed_emd(tfa.convert_rgb_to_yuv(x), tfa.convert_rgb_to_yuv(y), distance="both")
    gp = tape.gradient(lp, y)
    want = O.relaxed_emd(O.convert_rgb_to_yuv(st, np.float64), O.convert_rgb_to_yuv(pr, np.float64), "both", np.float64)
    _close(lp, want)
    assert float(gp.t[:, 3:].abs().max()) == 0.0 and float(gp.t[:, :3].abs().sum()) > 0
    with pytest.raises(KeyError):
        tfa.relaxed_emd(x, y, distance="sinkhorn")      # dist_metrics[distance], nn/losses.py:74


def _train_step_plain(tfa, z, dev, fused):
    """run_strotss.py:131-142 from the sampled features on (the recorded c_feat / p_feat stand for :135-136)."""
    tf = tfa.tf
    alpha = float(z["alpha"])
    loss_denom = (2. + alpha + 1. / max(alpha, 1.))                    # :92
    c_feat = _const(tfa, z["plain_content0"], dev)
    p_feat = _var(tfa, z["plain_pred0"], dev)
    style = _const(tfa, z["plain_style0"], dev)
    with tf.GradientTape() as tape:
        if fused:
            loss_fn = tfa.StrotssLoss(style, alpha)
            loss, loss_c, loss_s = loss_fn(c_feat, p_feat)
        else:
            loss_content, loss_style = tfa.ContentLoss(), tfa.StyleLoss(style, alpha=alpha)
            loss_c = loss_content(c_feat, p_feat)
            loss_s = loss_style(p_feat)
            loss = (alpha * loss_c + loss_s) / loss_denom              # :140
    grads = tape.gradient(loss, p_feat)                                # :141 (w.r.t. the features here)
    return loss, loss_c, loss_s, grads


@pytest.mark.parametrize("fused", [False, True])
def test_train_step_against_reference_code_golden(tfa, cuda_device, fused):
    z = np.load(os.path.join(GOLDEN, "ref_train_step.npz"))
    loss, loss_c, loss_s, grads = _train_step_plain(tfa, z, cuda_device, fused)
    _close(loss, z["plain_loss"]); _close(loss_c, z["plain_loss_c"]); _close(loss_s, z["plain_loss_s"], TOY_STYLE_RTOL)
    _gclose(grads, z["plain_grad0"], cos_min=0.99)


def test_masked_train_step_against_reference_code_golden(tfa, cuda_device):
    """run_strotss.py:104-125: one StyleLoss per region, mean over regions."""
    tf = tfa.tf
    z = np.load(os.path.join(GOLDEN, "ref_train_step.npz"))
    alpha = float(z["alpha"])
    loss_denom = (2. + alpha + 1. / max(alpha, 1.))
    R = int(z["masked_regions"])
    loss_content = tfa.ContentLoss()
    loss_styles = [tfa.StyleLoss(_const(tfa, z[f"masked_style{r}"], cuda_device), alpha=alpha) for r in range(R)]
    p_feats = [_var(tfa, z[f"masked_pred{r}"], cuda_device) for r in range(R)]
    with tf.GradientTape() as tape:
        loss, loss_c, loss_s = [], [], []
        for r in range(R):
            c_feat = _const(tfa, z[f"masked_content{r}"], cuda_device)
            loss_c.append(loss_content(c_feat, p_feats[r]))            # :116-117
            loss_s.append(loss_styles[r](p_feats[r]))
            loss.append((alpha * loss_c[-1] + loss_s[-1]) / loss_denom)
        loss, loss_c, loss_s = [tf.reduce_mean(tf.add_n(v) / float(R)) for v in (loss, loss_c, loss_s)]
    grads = tape.gradient(loss, p_feats)
    _close(loss, z["masked_loss"]); _close(loss_c, z["masked_loss_c"]); _close(loss_s, z["masked_loss_s"], TOY_STYLE_RTOL)
    for r in range(R):
        _gclose(grads[r], z[f"masked_grad{r}"], cos_min=0.99)


def test_handle_follows_the_tensor_device_and_is_released(tfa, cuda_device):
    st, co, pr = O.synth_problem(64, 48, 67, eps=0.1, seed=5)
    fn = tfa.StrotssLoss(_const(tfa, st, cuda_device), 4.0)
    assert fn.device_index == cuda_device.index
    loss, _, _ = fn(_const(tfa, co, cuda_device), _var(tfa, pr, cuda_device))
    _close(loss, O.total_loss(st, co, pr, 4.0))
    fn.close()
    assert fn.h is None
    with pytest.raises(RuntimeError):
        tfa._device_index(tfa.tf.Tensor(torch.zeros(2, 2)))              # a CPU tensor: no fallback
    n_live = len(tfa._live)
    import gc
    del loss
    gc.collect()
    assert len(tfa._live) <= n_live                                     # DLPack deleters return the output buffers
