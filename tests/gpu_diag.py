"""Stage-by-stage GPU diagnostic (developer tool, not a pytest file): runs every entry point against
the oracle and prints error statistics.  Usage on the GPU box:
    timeout 900 python tests/gpu_diag.py [--sizes 1024,4096] > gpurun_out/diag.log 2>&1
"""
import argparse
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import strotss_oracle as O  # noqa: E402
import strotss_tensorflow_b200 as S  # noqa: E402
from strotss_tensorflow_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)


def rel(a, b):
    return float(abs(a - b) / max(abs(b), 1e-30))


def gstats(g, ref):
    g = np.asarray(g, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    nr = np.linalg.norm(ref)
    return dict(norm_rel=abs(np.linalg.norm(g) - nr) / max(nr, 1e-30), diff_rel=np.linalg.norm(g - ref) / max(nr, 1e-30),
                cos=float((g * ref).sum() / max(np.linalg.norm(g) * nr, 1e-30)))


def stage(name):
    def deco(fn):
        def run(*a, **k):
            print(f"\n=== {name} ===", flush=True)
            t = time.time()
            try:
                fn(*a, **k)
                torch.cuda.synchronize()
                print(f"--- {name}: ok ({time.time() - t:.1f}s)", flush=True)
                return True
            except Exception:
                traceback.print_exc()
                print(f"--- {name}: FAILED", flush=True)
                return False
        return run
    return deco


@stage("debug_gemm")
def t_gemm(h):
    g = torch.Generator(device="cpu").manual_seed(0)
    for (m, n, k, tn) in [(128, 256, 64, 256), (128, 128, 64, 128), (128, 256, 256, 256), (300, 500, 2179, 256),
                          (300, 500, 2179, 128), (1024, 1024, 2179, 256), (77, 33, 100, 128)]:
        A = torch.randn(m, k, generator=g).to(dev)
        B = torch.randn(n, k, generator=g).to(dev)
        C = h.debug_gemm(A, B, 0.5, tn)
        ref = 0.5 * (A.bfloat16().double() @ B.bfloat16().double().T)
        err = (C.double() - ref).abs().max().item()
        print(f"  m={m} n={n} k={k} tile_n={tn}: max abs err {err:.3e} (ref max {ref.abs().max().item():.3e})", flush=True)


@stage("relaxed_emd cosine")
def t_remd(h, N, M, D):
    st, co, pr = O.synth_problem(N, M, D, eps=1.0, seed=0)
    x = torch.tensor(st, device=dev)
    y = torch.tensor(pr, device=dev)
    out, grad, ra, ca = h.relaxed_emd(x, y, "cosine", True, True)
    l64, g64, info = O.relaxed_emd(st, pr, "cosine", np.float64, True)
    out = out.cpu().numpy()
    print(f"  loss {out[0]:.7f} ref {l64:.7f} rel {rel(out[0], l64):.2e}; R_X {out[1]:.6f}/{info['R_X']:.6f} R_Y {out[2]:.6f}/{info['R_Y']:.6f} branch {out[3]}/{info['branch_x']}")
    ra = ra.cpu().numpy(); ca = ca.cpu().numpy()
    ragree = (ra == info["row_argmin"]); cagree = (ca == info["col_argmin"])
    thr = 4e-3
    print(f"  row argmin agree {ragree.mean():.4f} (gap>{thr}: {ragree[info['row_gap'] > thr].mean():.4f}, n={int((info['row_gap'] > thr).sum())}); "
          f"col argmin agree {cagree.mean():.4f} (gap>{thr}: {cagree[info['col_gap'] > thr].mean():.4f})")
    bad = np.where(~ragree)[0][:5]
    for b in bad:
        print(f"    row {b}: got {ra[b]} ref {info['row_argmin'][b]} gap {info['row_gap'][b]:.2e}")
    print("  grad:", gstats(grad.cpu().numpy(), g64))


@stage("relaxed_emd D=3 (palette)")
def t_pal(h, N, M):
    st, co, pr = O.synth_problem(N, M, 16, eps=1.0, seed=1)
    a = O.convert_rgb_to_yuv(st, np.float32).astype(np.float32)
    b = O.convert_rgb_to_yuv(pr, np.float32).astype(np.float32)
    for dist in ["both", "cosine", "l2"]:
        out, grad, ra, ca = h.relaxed_emd(torch.tensor(a, device=dev), torch.tensor(b, device=dev), dist, True, True)
        l64, g64, info = O.relaxed_emd(a, b, dist, np.float64, True)
        out = out.cpu().numpy()
        print(f"  {dist}: loss {out[0]:.7f} ref {l64:.7f} rel {rel(out[0], l64):.2e} branch {out[3]}/{info['branch_x']} "
              f"argmin agree {np.mean(ra.cpu().numpy() == info['row_argmin']):.4f}/{np.mean(ca.cpu().numpy() == info['col_argmin']):.4f} grad {gstats(grad.cpu().numpy(), g64)}")
    yuv = h.convert_rgb_to_yuv(torch.tensor(st, device=dev)).cpu().numpy()
    print("  yuv max err", np.abs(yuv - O.convert_rgb_to_yuv(st, np.float64)).max())


@stage("moment_matching")
def t_mom(h, N, M, D):
    st, co, pr = O.synth_problem(N, M, D, eps=1.0, seed=2)
    out, grad = h.moment_matching(torch.tensor(st, device=dev), torch.tensor(pr, device=dev), True)
    l64, g64, info = O.moment_matching(st, pr, np.float64, True)
    out = out.cpu().numpy()
    print(f"  loss {out[0]:.7f} ref {l64:.7f} rel {rel(out[0], l64):.2e}; l_cov {out[1]:.7f}/{info['l_cov']:.7f} l_mean {out[2]:.7f}/{info['l_mean']:.7f}")
    print("  grad:", gstats(grad.cpu().numpy(), g64))
    out2, _ = h.moment_matching(torch.tensor(st, device=dev), torch.tensor(st, device=dev), False)
    print("  moment_matching(x,x) =", out2.cpu().numpy())


@stage("self_similarity")
def t_ss(h, N, D):
    for eps in [1.0, 0.1, 0.01]:
        st, co, pr = O.synth_problem(N, 8, D, eps=eps, seed=3)
        out, grad = h.self_similarity(torch.tensor(pr, device=dev), torch.tensor(co, device=dev), True)
        l64, g64, _ = O.self_similarity(pr, co, np.float64, True)
        l32 = O.self_similarity(pr, co, np.float32)
        print(f"  eps={eps}: loss {out.item():.7e} ref {l64:.7e} rel {rel(out.item(), l64):.2e} (oracle fp32 rel {rel(l32, l64):.2e}) grad {gstats(grad.cpu().numpy(), g64)}")
    out2, _ = h.self_similarity(torch.tensor(co, device=dev), torch.tensor(co, device=dev), False)
    print("  self_similarity(x,x) =", out2.item())


@stage("total eval")
def t_total(N, M, D, check=True, iters=5):
    st, co, pr = O.synth_problem(N, M, D, eps=0.1, seed=0)
    mod = S.StrotssLoss(torch.tensor(st, device=dev), 16.0)
    p = torch.tensor(pr, device=dev); c = torch.tensor(co, device=dev)
    sc, grad, ra, ca = mod.handle.eval(p, c, 16.0, True, True)
    torch.cuda.synchronize()
    s = sc.cpu().numpy()
    print("  scalars:", np.array2string(s[:14], precision=6))
    if check:
        l64, g64, info = O.total_loss(st, co, pr, 16.0, np.float64, True)
        print(f"  total {s[0]:.7f} ref {l64:.7f} rel {rel(s[0], l64):.2e}; loss_c rel {rel(s[1], info['loss_c']):.2e} loss_s rel {rel(s[2], info['loss_s']):.2e} "
              f"l_m {rel(s[3], info['l_m']):.2e} l_remd {rel(s[4], info['l_remd']):.2e} l_pal {rel(s[5], info['l_palette']):.2e}")
        print("  grad:", gstats(grad.cpu().numpy(), g64))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        mod.handle.eval(p, c, 16.0, True, False)
    torch.cuda.synchronize()
    l0 = mod.handle.launch_count
    e0.record()
    for _ in range(iters):
        mod.handle.eval(p, c, 16.0, True, False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    D_ = D
    falg = 2.0 * M * N * D_ + 2 * (2.0 * N * N * D_) + 2.0 * N * N * D_ + 2 * (2.0 * N * D_ * D_)
    print(f"  N={N} M={M}: {ms:.3f} ms/eval, {1000 / ms:.1f} evals/s, F_alg {falg:.3e} -> {falg / ms / 1e9:.1f} TFLOP/s; launches/eval {(mod.handle.launch_count - l0) / iters:.0f}; workspace {mod.handle.workspace_bytes / 1e6:.0f} MB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024")
    ap.add_argument("--big", default="")
    args = ap.parse_args()
    print(torch.cuda.get_device_name(0), _lib.load().strotss_version())
    h = S.shared_handle(dev)
    ok = t_gemm(h)
    if not ok:
        print("GEMM core failed; stopping")
        return
    t_remd(h, 700, 517, 2179)
    t_pal(h, 1000, 777)
    t_mom(h, 600, 500, 2179)
    t_ss(h, 640, 2179)
    for n in [int(s) for s in args.sizes.split(",") if s]:
        t_total(n, n, 2179, check=(n <= 2048))
    for n in [int(s) for s in args.big.split(",") if s]:
        t_total(n, n, 2179, check=False, iters=3)


if __name__ == "__main__":
    main()
