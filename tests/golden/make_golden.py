"""Generate tests/golden/*.npz from the oracle (fp64).  The reference has no golden vectors and cannot run
here (no TensorFlow), so these fixtures pin the ORACLE's outputs on seeded synthetic inputs; they guard
against regressions of the oracle and give the GPU tests fixed targets (incl. argmin rows and gaps).
The companion make_reference_golden.py produces ref_*.npz for the same problems from the reference's own
loss code run over a stand-in for its TensorFlow ops; tests/test_reference_golden.py holds the two together.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import strotss_oracle as O  # noqa: E402

CASES = {
    # name: (N, M, D, eps, seed, alpha)
    "small_d67": (48, 40, 67, 0.1, 11, 16.0),        # inputs stored in the file
    "ragged_d2179": (333, 517, 2179, 0.1, 5, 16.0),  # mask-config region sizes; inputs by seed
    "default_d2179_eps1": (256, 256, 2179, 1.0, 0, 8.0),
    "near_d2179_eps001": (256, 200, 2179, 0.01, 7, 2.0),
}


def main():
    for name, (N, M, D, eps, seed, alpha) in CASES.items():
        st, co, pr = O.synth_problem(N, M, D, eps=eps, seed=seed)
        loss, grad, info = O.total_loss(st, co, pr, alpha, np.float64, True)
        out = dict(N=N, M=M, D=D, eps=eps, seed=seed, alpha=alpha,
                   total=loss, loss_c=info["loss_c"], loss_s=info["loss_s"], l_m=info["l_m"], l_remd=info["l_remd"],
                   l_palette=info["l_palette"], remd_RX=info["remd"]["R_X"], remd_RY=info["remd"]["R_Y"],
                   pal_RX=info["palette"]["R_X"], pal_RY=info["palette"]["R_Y"],
                   remd_row_argmin=info["remd"]["row_argmin"].astype(np.int32),
                   remd_col_argmin=info["remd"]["col_argmin"].astype(np.int32),
                   remd_row_gap=info["remd"]["row_gap"].astype(np.float32),
                   remd_col_gap=info["remd"]["col_gap"].astype(np.float32),
                   grad_norm=np.linalg.norm(grad), grad_rowsum=grad.sum(axis=1).astype(np.float64),
                   grad_colsum=grad.sum(axis=0).astype(np.float64),
                   grad_head=grad[:8, :16].astype(np.float64),
                   input_checksum=np.array([st.astype(np.float64).sum(), co.astype(np.float64).sum(), pr.astype(np.float64).sum()]))
        if D <= 128:
            out.update(style=st, content=co, pred=pr, grad=grad.astype(np.float64))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "total", loss, "grad_norm", out["grad_norm"])


if __name__ == "__main__":
    main()
