"""SURVEY.md section 8(d) sweep: loss+grad evaluations at N = M in {1024, 2048, 4096, 8192, 16384}, D = 2179, one GPU, inputs
resident in HBM, CUDA events around `--steps` back-to-back evaluations after 3 warm-ups.  One JSON line per size."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--sizes", type=int, nargs="*", default=[1024, 2048, 4096, 8192, 16384])
    args = ap.parse_args()
    import torch
    import strotss_tensorflow_b200 as S
    from strotss_tensorflow_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("size_sweep.py: no CUDA device; the B200 path has no CPU fallback")
    dev = torch.device("cuda", 0)
    pk = bench.peaks()
    h = S.Handle(dev)
    for n in args.sizes:
        style, content, pred = bench.synth_torch(n, n, bench.D_FEAT, 1.0, 0, dev)
        h.set_style_target(style)
        ms, out = bench.timed_events(torch, lambda: h.eval(pred, content, bench.ALPHA, True, False), args.steps, 3)
        flops = bench.f_alg(n, n)
        tf = flops / (ms * 1e-3) / 1e12
        print(json.dumps({"N": n, "M": n, "D": bench.D_FEAT, "ms_per_eval": round(ms, 4), "evals_per_s": round(1e3 / ms, 1),
                          "f_alg": flops, "tflops_alg": round(tf, 1), "frac_of_burst_peak": round(tf / pk["tf_burst"], 3),
                          "loss": float(out[0][_lib.S_TOTAL].item()), "steps": args.steps,
                          "workspace_mb": round(h.workspace_bytes / 1e6, 1),
                          "l2": "no flush; the three inputs exceed the 126 MB L2 from N = 8192 on"}), flush=True)
        del style, content, pred


if __name__ == "__main__":
    main()
