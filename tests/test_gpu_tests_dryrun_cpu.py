"""CPU dry run of the GPU tests that compare the CUDA path with the reference-code fixtures: the test bodies are executed
against a stand-in of the package API whose functions are backed by the oracle, so that their own logic (fixture keys,
shapes, tolerances) is exercised where no GPU exists.  Says nothing about the kernels."""
import types

import numpy as np
import torch

from oracle import pixel_oracle as P
from oracle import strotss_oracle as O

import test_gpu_parity as TP
import test_gpu_pixel as TX


def _t32(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)


class _Sampling:
    def __init__(self, n):
        pass

    def _sample(self, xs, idx, bilinear):
        return _t32(O.sample_hypercolumns([x.numpy() for x in xs], idx.numpy(), bilinear))


class _Handle:
    def __init__(self, st):
        self.st = st

    def eval(self, pr, co, alpha, want_grad, want_arg=False):
        from strotss_tensorflow_b200 import _lib
        loss, grad, info = O.total_loss(self.st.numpy(), co.numpy(), pr.numpy(), alpha, np.float32, True)
        sc = np.zeros(_lib.NUM_SCALARS, np.float32)
        for slot, v in [(_lib.S_TOTAL, loss), (_lib.S_LOSS_C, info["loss_c"]), (_lib.S_LOSS_S, info["loss_s"]), (_lib.S_L_M, info["l_m"]),
                        (_lib.S_L_REMD, info["l_remd"]), (_lib.S_L_PALETTE, info["l_palette"])]:
            sc[slot] = v
        return torch.tensor(sc), _t32(grad), None, None


def _np64(x):
    return x[0].numpy().astype(np.float64)


def _resize_long_side(x, m):
    h, w = x.shape[1], x.shape[2]
    f = max(h / m, w / m)
    return _t32(P.resize_bilinear(_np64(x), int(h / f), int(w / f)))[None]


FAKE = types.SimpleNamespace(
    Sampling=_Sampling,
    StrotssLoss=lambda st, alpha: types.SimpleNamespace(handle=_Handle(st)),
    make_laplacian_pyramid=lambda x, n: [_t32(a)[None] for a in P.make_laplacian_pyramid(_np64(x), n)],
    fold_laplacian_pyramid=lambda xs: _t32(P.fold_laplacian_pyramid([_np64(a) for a in xs]))[None],
    make_laplacian=lambda x, ret: tuple(_t32(a)[None] for a in P.make_laplacian(_np64(x))),
    resize=_resize_long_side,
    resize_like=lambda x, b: _t32(P.resize_bilinear(_np64(x), b.shape[1], b.shape[2]))[None])
CPU = torch.device("cpu")


def test_dry_run_of_the_reference_golden_gpu_tests():
    for mode in ("bilinear", "nearest"):
        TP.test_sampler_against_reference_code_golden(FAKE, CPU, mode)
    for name in ("small_d67", "near_d2179_eps001"):
        TP.test_total_against_reference_code_golden(FAKE, CPU, name)
    TX.test_pyramid_against_reference_code_golden(FAKE, CPU)
