#!/usr/bin/env python
"""Benchmark of the STROTSS loss hot path (BASELINE.json metric: loss+grad evals/sec at
N=M=16384, D=2179) on B200, with the CPU restatement of the reference timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload large|default]

One "step" = one evaluation: total loss (self-similarity + moment matching + relaxed EMD + palette)
and its gradient w.r.t. the (N, 2179) prediction hypercolumns, style-side preparation excluded
(constant per scale, run_strotss.py:100,128).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

D_FEAT = 2179
ALPHA = 16.0
WORKLOADS = {
    # name: (N, M)
    "large": (16384, 16384),     # BASELINE.json configs[3]: large-sample loss microbench (the metric's config)
    "default": (1024, 1024),     # reference default sample count (run_strotss.py:68)
    "masked": (1024, 1024),      # BASELINE.json configs[2]: R = 3 regions (N_r, M_r) of SURVEY 8d, grouped evaluation
}
MASKED_REGIONS = [(1024, 1024), (700, 1024), (333, 517)]


def f_alg(N, M, D=D_FEAT):
    """Algorithmic FLOPs per evaluation: no recompute / padding credit."""
    return 2.0 * M * N * D + 2 * (2.0 * N * N * D) + 2.0 * N * N * D + 2 * (2.0 * N * D * D)


def f_ref(N, M, D=D_FEAT):
    """Dense FLOPs the reference's framework autodiff executes (5 fwd + 5 bwd GEMMs)."""
    return 2 * (2.0 * M * N * D) + 4 * (2.0 * N * N * D) + 4 * (2.0 * N * D * D)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


def synth_torch(N, M, D, eps, seed, device):
    """Synthetic hypercolumns of SURVEY.md section 8d (RGB in [0,1), ReLU-like heavy-tailed VGG channels)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    sigma = torch.exp(torch.randn(D - 3, generator=g, device=device))

    def feats(n):
        out = torch.empty(n, D, device=device, dtype=torch.float32)
        out[:, :3] = torch.rand(n, 3, generator=g, device=device)
        out[:, 3:] = torch.clamp_min(sigma[None, :] * (torch.randn(n, D - 3, generator=g, device=device) + 0.3), 0.0)
        return out

    style = feats(M)
    content = feats(N)
    noise = torch.randn(N, D, generator=g, device=device)
    pred = torch.clamp_min(content + eps * content.abs().mean() * noise, 0.0)
    return style, content, pred


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return None
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU baseline: the torch-CPU port of the reference op sequence (oracle/torch_port.py)
# ------------------------------------------------------------------------------------------
def cpu_eval_seconds(n_sample, reps, warmup, threads):
    import torch
    from oracle import torch_port as T
    torch.set_num_threads(threads)
    st, co, pr = synth_torch(n_sample, n_sample, D_FEAT, 1.0, 0, torch.device("cpu"))
    for _ in range(warmup):
        T.total_loss_and_grad(st, co, pr, ALPHA)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        T.total_loss_and_grad(st, co, pr, ALPHA)
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args, N, M):
    """`--impl reference`: the reference's own CPU path.  TensorFlow is not installable here, so this
    times the CPU restatement of the reference op sequence (materialised matrices + framework
    autodiff), all host threads, each step a bounded sample scaled by the reference FLOP ratio."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    n_s = min(N, 2048)
    scale = f_ref(N, M) / f_ref(n_s, n_s)
    times = cpu_eval_seconds(n_s, args.steps, args.warmup, threads)
    t_step = sum(times) / len(times)
    value = 1.0 / (t_step * scale)
    sample = (f"torch-CPU port of nn/losses.py (fp32, materialised matrices, autograd), N=M={n_s} full evaluation per step, "
              f"{threads} threads; evals/s scaled to N=M={N} by the reference FLOP ratio {scale:.2f}")
    line = {
        "impl": "reference", "metric": f"loss+grad evals/sec at N=M={N}, D={D_FEAT}", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"large-sample loss microbench N=M={N} D={D_FEAT} alpha={ALPHA} eps={args.eps}"
                   if N > 1024 else f"default sample count N=M={N} D={D_FEAT} alpha={ALPHA} eps={args.eps}", "sampled_as": f"N=M={n_s}"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_masked(args, dev, S, _lib):
    """BASELINE configs[2] (masked region-guided transfer): one step = the loss + gradient of one masked train_step
    (run_strotss.py:112-124) over R = 3 ragged regions, grouped (one C-ABI call, regions on concurrent streams)
    against the same regions evaluated one after the other through strotss_eval."""
    import torch
    probs = [synth_torch(N, M, D_FEAT, args.eps, 100 + r, dev) for r, (N, M) in enumerate(MASKED_REGIONS)]
    styles = [p[0] for p in probs]; contents = [p[1] for p in probs]; preds = [p[2] for p in probs]
    h = S.Handle(dev)
    h.set_style_targets_grouped(styles)
    singles = []
    for st in styles:
        hs = S.Handle(dev); hs.set_style_target(st); singles.append(hs)

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, out

    l0 = h.launch_count
    ms_g, out = timed(lambda: h.eval_grouped(preds, contents, ALPHA, True))
    launches = (h.launch_count - l0) // (args.steps + max(args.warmup, 3))
    ms_s, _ = timed(lambda: [hs.eval(p, c, ALPHA, True) for hs, p, c in zip(singles, preds, contents)])
    line = {"metric": "masked train_step loss+grad evals/sec, R=3 regions", "value": 1000.0 / ms_g, "unit": "evals/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_g, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"masked transfer, regions (N_r, M_r) = {MASKED_REGIONS}, D={D_FEAT}, alpha={ALPHA}",
                       "l2": "inputs fit in L2 (launch-bound regime)"},
            "gpu_launches_per_step": int(launches), "sequential_ms_per_step": ms_s,
            "loss": float(out[0][_lib.S_TOTAL].item())}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="large", choices=sorted(WORKLOADS))
    ap.add_argument("--eps", type=float, default=1.0, help="pred = content + eps*noise (SURVEY 8d)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--samples", type=int, default=0,
                    help="override the sample count N = M of the workload (SURVEY 8d sweep: 1024 ... 16384)")
    ap.add_argument("--graph", action="store_true",
                    help="replay the evaluation from a CUDA graph captured around the C-ABI call (launch-bound small workloads)")
    ap.add_argument("--mode", default="replicas", choices=["rowshard", "replicas"],
                    help="N>1: 'replicas' = one problem per GPU, no collective (throughput mode, weak scaling; the headline); "
                         "'rowshard' = ONE evaluation sharded by prediction rows over all GPUs (latency mode, strong scaling). "
                         "In replicas mode the row-sharded latency is measured as well and reported under 'rowshard'.")
    args = ap.parse_args()
    N, M = WORKLOADS[args.workload]
    if args.samples > 0:
        N = M = args.samples

    if args.impl == "reference":
        run_reference(args, N, M)
        return

    import torch
    import torch.distributed as dist
    import strotss_tensorflow_b200 as S
    from strotss_tensorflow_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "masked":
        run_masked(args, dev, S, _lib)
        return

    # rowshard: every rank holds the same (replicated) inputs and computes its rows of ONE evaluation;
    # replicas: every rank evaluates its own (seeded per rank) problem of the full size.
    rowshard = world > 1 and args.mode == "rowshard"
    style, content, pred = synth_torch(N, M, D_FEAT, args.eps, 0 if rowshard else rank, dev)
    h = S.Handle(dev)
    if rowshard:
        from strotss_tensorflow_b200 import distributed as Dm
        Dm.attach(h)
    h.set_style_target(style)
    scalars = None
    for _ in range(max(args.warmup, 3)):
        scalars, grad, _, _ = h.eval(pred, content, ALPHA, True, False)
    barrier()

    # ---- device-resident timed region ------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    graph = None
    if args.graph:
        # stream capture of the library's launch sequence (branch streams fork from / join to the capture stream)
        l0 = h.launch_count
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            scalars, grad, _, _ = h.eval(pred, content, ALPHA, True, False)
        per_eval = h.launch_count - l0
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
    else:
        h.profile_enable(True)
        h.profile_read()
    l0 = h.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            scalars, grad, _, _ = h.eval(pred, content, ALPHA, True, False)
    e1.record()
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launches = (per_eval * args.steps) if graph is not None else (h.launch_count - l0)
    phases = h.profile_read() if graph is None else {}
    h.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    total = float(scalars[_lib.S_TOTAL].item())

    # ---- end to end through the host-buffer entry points ---------------------------------
    # every step copies that step's inputs from pinned host memory and reads its scalars + gradient back to the host;
    # (a) serial: strotss_eval_host, one blocking call per step; (b) pipelined: strotss_eval_host_submit/_wait with two
    # evaluations in flight, so the PCIe copies of neighbouring steps overlap the kernels (independent evaluations --
    # BASELINE "throughput mode").  (b) is the headline e2e; (a) is reported beside it.
    ph = torch.empty(N, D_FEAT, dtype=torch.float32).pin_memory(); ph.copy_(pred)
    ch = torch.empty(N, D_FEAT, dtype=torch.float32).pin_memory(); ch.copy_(content)
    gh = [torch.empty(N, D_FEAT, dtype=torch.float32).pin_memory() for _ in range(2)]
    sh = [torch.empty(_lib.NUM_SCALARS, dtype=torch.float32) for _ in range(2)]
    for _ in range(2):
        h.eval_host(ph, ch, ALPHA, gh[0], sh[0])
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.eval_host(ph, ch, ALPHA, gh[0], sh[0])
    torch.cuda.synchronize()
    e2e_serial_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    barrier()

    def pipelined(nsteps):
        prev = None
        for k in range(nsteps):
            t = h.eval_host_submit(ph, ch, ALPHA, gh[k & 1], sh[k & 1])
            if prev is not None:
                h.eval_host_wait(prev)
            prev = t
        h.eval_host_wait(prev)

    pipelined(3)
    e2e_steps = max(4, args.steps)
    barrier()
    t0 = time.perf_counter()
    pipelined(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    e2e_loss = float(sh[(e2e_steps - 1) & 1][_lib.S_TOTAL])
    barrier()

    # ---- N>1, replicas mode: also time ONE evaluation row-sharded over all GPUs (same inputs everywhere) ----
    shard_info = None
    if world > 1 and not rowshard:
        from strotss_tensorflow_b200 import distributed as Dm
        style0, content0, pred0 = synth_torch(N, M, D_FEAT, args.eps, 0, dev)
        hs = S.Handle(dev)
        Dm.attach(hs)
        hs.set_style_target(style0)
        for _ in range(3):
            hs.eval(pred0, content0, ALPHA, True, False)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            sc_s, _, _, _ = hs.eval(pred0, content0, ALPHA, True, False)
        s1.record()
        barrier()
        ms_shard = max_over_ranks(s0.elapsed_time(s1) / args.steps)
        r0s, r1s = hs.shard_rows(N)
        shard_info = {"value": 1000.0 / ms_shard, "unit": "evals/s", "ms_per_step": ms_shard, "scaling": "strong",
                      "rows_per_rank": r1s - r0s, "loss": float(sc_s[_lib.S_TOTAL].item()),
                      "collectives_per_eval": f"1 allreduce-max of 2x{M} packed u64 minima + 1 allreduce-sum of {16 + D_FEAT} floats (NCCL)"}
        del hs

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    jobs = 1 if rowshard else world                     # evaluations completed per step across the job
    value = jobs * 1000.0 / ms_step
    # dominant kernel: self-similarity stage 1 (one launch per 2048-row panel).  Algorithmic work of the
    # launches of one step = the two Gram products Xd, Yd restricted to this rank's rows (2*(2*N^2*D)
    # per evaluation); the kernel executes 3 bf16 K-passes (delta form) over the tiles it visits -- all of them
    # when row-sharded, the upper block triangle (36/64 at 8 panels) on a single GPU where symmetry is exploited.
    ss1_ms, ss1_n = phases.get("ss_stage1_gemm", (0.0, 0))
    own_rows = h.shard_rows(N)[1] - h.shard_rows(N)[0] if rowshard else N
    alg_flops_step = 2 * (2.0 * own_rows * N * D_FEAT)
    # tiles visited / all tiles: everything when row-sharded or for a single panel; on one GPU the exact upper block
    # triangle of 256 x 256 tiles (rectangular panels, 36/64 at N = 16384, with STROTSS_NO_TRAP=1)
    nt = -(-N // 256)
    if rowshard or N <= 2048:
        visited = 1.0
    elif os.environ.get("STROTSS_NO_TRAP"):
        ph = int(os.environ.get("STROTSS_PANEL", "4096"))
        visited = sum(min(ph, N - p * ph) * (N - p * ph) for p in range(-(-own_rows // ph))) / (float(N) * N)
    else:
        visited = nt * (nt + 1) / 2.0 / (nt * nt)
    ach = alg_flops_step / (ss1_ms / args.steps * 1e-3) / 1e12 if ss1_n else None
    roof = {"bound": "tensor", "kernel": "ss1_pair_merged_kernel (self-similarity stage 1, cta_group::2; all launches of a step)",
            "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": (ach / pk["tf_sust"]) if ach else None,
            "frac_of_burst_peak": (ach / pk["tf_burst"]) if ach else None,
            # ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the launches of one step of the DEFAULT build at this
            # workload (4096-row panels, triangle walk, merged K loop: profiles/r01_v21_ss1_merged_remd_skew_ncu_summary.txt;
            # the kernel with three separate K loops moved 1.808e9); other switch settings: not captured
            "traffic": 1.628e9 if (world == 1 and N == 16384 and not os.environ.get("STROTSS_NO_TRAP")
                                   and not os.environ.get("STROTSS_PANEL") and not os.environ.get("STROTSS_SS1_MERGED")) else None,
            "traffic_unit": "bytes per step (all launches of the kernel)",
            # operands: A 4 panels x 3 x 4096 x 2240 bf16 = 220 MB, B 550 MB; P panels 2080 tiles x 128 KB = 272 MB
            "algorithmic_bytes": 1.042e9 if (world == 1 and N == 16384) else None,
            "peak_source": pk["source"] + " bf16 sustained (kernel timed inside a long step)",
            "launches_per_step": ss1_n / args.steps if ss1_n else None,
            "achieved_executed": (ach * 1.5 * visited) if ach else None,
            "executed_over_algorithmic": 1.5 * visited,
            "note": "algorithmic = 2 Gram GEMMs (Xd, Yd); executed = 3 bf16 K-passes over the visited tiles "
                    "(symmetry: only the upper block triangle, 2080 of 4096 tiles at N = 16384, is visited on a single GPU)"}
    line = {
        "metric": f"loss+grad evals/sec at N=M={N}, D={D_FEAT}", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong" if rowshard else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"large-sample loss microbench N=M={N} D={D_FEAT} alpha={ALPHA} eps={args.eps}"
                   if args.workload == "large" else f"default sample count N=M={N} D={D_FEAT} alpha={ALPHA} eps={args.eps}",
                   "parallelism": "single GPU" if world == 1 else (
                       f"one evaluation row-sharded over {world} GPUs by prediction rows; per evaluation one NCCL allreduce-max of "
                       f"2x{M} packed u64 minima + one allreduce-sum of {16 + D_FEAT} floats" if rowshard
                       else f"{world} independent replicas (one problem per GPU, no collective)"),
                   "l2": "inputs (3 x %.0f MB fp32) exceed the 126 MB L2; no flush" % (N * D_FEAT * 4 / 1e6)
                   if N * D_FEAT * 4 * 3 > 126e6 else "inputs fit in L2 (launch-bound regime)",
                   "precision": "bf16 operands (delta-form self-similarity), fp32 accumulate/reductions",
                   "launch": "CUDA graph replay of one captured strotss_eval" if args.graph else "direct C-ABI calls"},
        "clocks": clocks,
        "e2e": {"value": jobs * 1000.0 / e2e_ms, "unit": "evals/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 2 * N * D_FEAT * 4, "d2h_bytes_per_step": own_rows * D_FEAT * 4 + _lib.NUM_SCALARS * 4,
                "mode": "strotss_eval_host_submit/_wait, 2 evaluations in flight (host<->device copies of neighbouring steps "
                        "overlap the kernels); every step copies its inputs from pinned host memory and reads scalars + gradient back",
                "serial_value": jobs * 1000.0 / e2e_serial_ms, "serial_ms_per_step": e2e_serial_ms,
                "serial_mode": "strotss_eval_host, one blocking call per step", "loss": e2e_loss},
        "gpu_launches": int(launches),
        "roofline": roof,
        "whole_eval": {"f_alg": f_alg(N, M), "tflops_alg": f_alg(N, M) / (ms_step * 1e-3) / 1e12,
                       "frac_of_burst_peak": f_alg(N, M) / (ms_step * 1e-3) / 1e12 / pk["tf_burst"],
                       "frac_of_sustained_peak": f_alg(N, M) / (ms_step * 1e-3) / 1e12 / pk["tf_sust"]},
        "phases_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in phases.items()},
        "loss": total,
    }
    if shard_info is not None:
        line["rowshard"] = shard_info
    if world == 1 and not args.no_cpu_baseline:
        threads = len(os.sched_getaffinity(0))
        n_s = min(N, 2048)
        t_probe = cpu_eval_seconds(n_s, 1, 1, threads)[0]
        if N >= 4096 and t_probe * (f_ref(4096, 4096) / f_ref(n_s, n_s)) < 12.0:
            n_s = 4096
        t_one = t_probe * (f_ref(n_s, n_s) / f_ref(min(N, 2048), min(N, 2048)))
        reps = max(2, min(24, int(10.0 / max(t_one, 1e-3))))          # ~10 s of all-core CPU work
        times = cpu_eval_seconds(n_s, reps, 0, threads)
        t = sum(times) / len(times)
        scale = f_ref(N, M) / f_ref(n_s, n_s)
        # the reference pins TensorFlow to one intra-op / one inter-op thread (nn/rand.py:16-17): time that setting too,
        # on a smaller sample (N = M = 1024, one evaluation)
        n_1 = min(N, 1024)
        t_1s = cpu_eval_seconds(n_1, 5, 1, 1)
        t_1 = sum(t_1s) / len(t_1s)
        one_thread = 1.0 / (t_1 * f_ref(N, M) / f_ref(n_1, n_1))
        line["cpu_baseline"] = {
            "value": 1.0 / (t * scale), "unit": "evals/s", "cores": threads, "kind": "port",
            "value_1_thread": one_thread,
            "sample_1_thread": f"same port pinned to 1 thread (the reference's own setting, nn/rand.py:16-17), 5 evaluations at "
                               f"N=M={n_1} ({t_1:.2f} s), scaled by the reference FLOP ratio",
            "sample": f"torch-CPU port of the reference op sequence (fp32, materialised matrices, autograd), {reps} evaluations at "
                      f"N=M={n_s} ({t:.2f} s each), scaled to N=M={N} by the reference FLOP ratio {scale:.2f}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
