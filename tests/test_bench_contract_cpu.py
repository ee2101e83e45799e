"""CPU checks of bench.py's contract: the parts of the measurement harness that do not need a GPU.

* the reference arm (`--impl reference`) runs here, prints ONE JSON line with the keys the driver parses, is silent on
  ranks other than 0, and times full evaluations at the workload's own size (`same_config`);
* our arm refuses to run without a CUDA device (no CPU fallback: nothing may be measured through the oracle);
* the roofline inputs: algorithmic FLOPs per evaluation follow SURVEY.md section 8(d), the DRAM traffic is read from a
  committed ncu export under profiles/ (not a literal), the peaks come from MEASURED_PEAKS.json or the stated fallback.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _run(argv, env_extra=None, timeout=300):
    env = dict(os.environ)
    env.pop("RANK", None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    if env_extra:
        env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + argv, cwd=ROOT, env=env, capture_output=True,
                          text=True, timeout=timeout)


@pytest.fixture(scope="module")
def reference_line():
    r = _run(["--impl", "reference", "--workload", "default", "--steps", "3", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "the reference arm prints exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys(reference_line):
    d = reference_line
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None                       # BASELINE.md holds no published number for this metric
    assert d["gpu_launches"] == 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["steps"] >= 3 and d["warmup"] >= 1
    assert d["value"] == pytest.approx(1e3 / d["ms_per_step"], rel=1e-6)
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] == d["value"] and c["cores"] == len(os.sched_getaffinity(0))
    assert "FULL evaluations at N=M=1024" in c["sample"]


def test_reference_arm_times_the_workload_itself(reference_line):
    cfg = reference_line["config"]
    assert cfg["same_config"] is True and "N=M=1024" in cfg["workload"] and "full size" in cfg["sampled_as"]
    assert "N=M=1024" in reference_line["metric"]
    # the loss of the synthetic problem is a property of the workload, not of the machine
    assert 0.01 < reference_line["loss"] < 0.2


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--workload", "default", "--steps", "3", "--warmup", "1", "--gpus", "2"], {"RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_our_arm_refuses_to_run_without_a_gpu():
    r = _run(["--steps", "1", "--warmup", "1"], timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_algorithmic_flops_follow_the_survey():
    N = M = 16384
    D = bench.D_FEAT
    assert D == 2179
    unit = 2.0 * N * N * D
    # cost matrix + two Gram matrices + their gradient product + covariance forward / backward  
    assert bench.f_alg(N, M) == pytest.approx(unit + 2 * unit + unit + 2 * (2.0 * N * D * D), rel=1e-12)
    assert bench.f_alg(N, M) == pytest.approx(4.99e12, rel=2e-3)
    assert bench.f_ref(N, M) > bench.f_alg(N, M)          # framework autodiff executes more than the algorithmic count


def test_roofline_traffic_is_read_from_a_committed_ncu_export():
    got = bench.ncu_traffic("ss1_pair_merged_kernel")
    assert got is not None
    total, launches, path = got
    assert path.startswith("profiles/") and path.endswith("_ncu_raw.csv") and os.path.exists(os.path.join(ROOT, path))
    assert launches == 4                                   # the four row panels of one evaluation at N = 16384
    operands = 3 * 16384 * 2240 * 2                        # x^, y^, delta read once (bf16, K padded to 2240)
    assert operands < total < 12 * operands                # 1.62 GB in the round-2 capture: 7.4 x the operand bytes
    assert bench.ncu_traffic("no_such_kernel") is None


def test_peaks_come_from_the_measured_file_or_the_stated_fallback():
    p = bench.peaks()
    assert set(p) >= {"hbm", "tf_burst", "tf_sust", "source"}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        assert p["source"].startswith("measured")
    else:
        assert p["source"].startswith("fallback")
    assert 4000 < p["hbm"] < 8000 and 1000 < p["tf_sust"] <= p["tf_burst"] < 2300


def test_size_sweep_extra_over_a_stand_in_handle(monkeypatch):
    """The sweep that rides in the default line's extras, executed over a stand-in handle (logic only): one entry per size with the
    SURVEY 8d FLOP count, and an error entry instead of an exception when the handle fails."""
    import types

    import torch

    calls = []

    class _Handle:
        def __init__(self, dev):
            self.n = None

        def set_style_target(self, style):
            self.n = style.shape[0]

        def eval(self, pred, content, alpha, want_grad, want_arg):
            assert pred.shape == content.shape == (self.n, 8) and alpha == bench.ALPHA and want_grad and not want_arg
            calls.append(self.n)
            return None

    monkeypatch.setattr(bench, "synth_torch", lambda N, M, D, eps, seed, dev: tuple(torch.zeros(n, 8) for n in (M, N, N)))
    monkeypatch.setattr(bench, "timed_events", lambda torch_, fn, steps, warmup: ([fn() for _ in range(steps + warmup)], 2.0)[1:] + (None,))
    got = bench.size_sweep_extra("cpu", types.SimpleNamespace(Handle=_Handle), torch, 1.0, sizes=(256, 512), steps=2)
    assert list(got) == ["N=M=256", "N=M=512"] and calls == [256] * 5 + [512] * 5
    assert got["N=M=512"]["ms_per_step"] == 2.0
    assert got["N=M=512"]["tflops_alg"] == pytest.approx(bench.f_alg(512, 512) / 2e-3 / 1e12, abs=0.06)

    class _Broken(_Handle):
        def set_style_target(self, style):
            raise RuntimeError("boom")

    assert bench.size_sweep_extra("cpu", types.SimpleNamespace(Handle=_Broken), torch, 1.0) == {"error": "boom"}
