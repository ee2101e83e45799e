"""CPU check of what the built library contains (B200_PROFILING.md, "What proves a Blackwell-native kernel"): every GEMM
kernel of libstrotss_b200.so issues tcgen05.mma (SASS UTCHMMA) fed by TMA (UTMALDG) and reads its accumulators back
with tcgen05.ld (LDTM); the CTA-pair kernels use the cta_group::2 forms; nothing uses the legacy mma.sync path (HMMA).
cuobjdump runs without a GPU."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sass():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from strotss_tensorflow_b200 import build
    lib = build.build()
    out = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    funcs, cur = {}, None
    for ln in out.stdout.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(ln)
    assert "sm_100a" in out.stdout or "sm_100" in out.stdout
    return {k: "\n".join(v) for k, v in funcs.items()}


def _kernels(sass, pattern):
    found = {k: v for k, v in sass.items() if re.search(pattern, k)}
    assert found, f"no kernel matches {pattern}"
    return found


def test_every_gemm_kernel_is_tcgen05_fed_by_tma(sass):
    gemm = _kernels(sass, r"gemm_kernel|gemm2_kernel|gemm2w_kernel|gemm2s_kernel|ss1_kernel|ss1_pair")
    assert len(gemm) >= 12
    for name, text in gemm.items():
        assert "UTCHMMA" in text, f"{name}: no tcgen05.mma"
        assert "UTMALDG" in text, f"{name}: operands are not staged by TMA"
        assert "LDTM" in text, f"{name}: accumulators are not read with tcgen05.ld"
        assert "UTCBAR" in text, f"{name}: no tcgen05.commit"


def test_pair_kernels_use_cta_group_2(sass):
    pair = _kernels(sass, r"gemm2_kernel|gemm2w_kernel|gemm2s_kernel|ss1_pair")
    for name, text in pair.items():
        assert "UTCHMMA.2CTA" in text and "UTMALDG.2D.2CTA" in text and "UTCBAR.2CTA.MULTICAST" in text, name
    single = _kernels(sass, r"2sb11gemm_kernel|2sb10ss1_kernel")
    for name, text in single.items():
        assert "UTCHMMA.2CTA" not in text, name


def test_the_operand_sharing_kernels_issue_two_accumulators_per_stage(sass):
    """ss1_pair_merged_kernel / gemm2s_kernel / gemm2w_kernel: one A tile, two B tiles -> more MMA issue sites than the plain
    pair kernel of the same epilogue family has."""
    n = lambda pat: min(t.count("UTCHMMA.2CTA") for t in _kernels(sass, pat).values())   # noqa: E731
    assert n(r"ss1_pair_merged") > n(r"ss1_pair_kernel")
    assert n(r"gemm2s_kernel") > n(r"gemm2_kernel")


def test_no_legacy_tensor_path_anywhere(sass):
    for name, text in sass.items():
        assert not re.search(r"\bHMMA\b|\bHGMMA\b|\bIMMA\b", text), f"{name} uses a legacy tensor-core path"
