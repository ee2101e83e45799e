"""CPU tests of the C-ABI boundary: the in-tree library loads without a GPU, exports every symbol
declared in include/strotss_b200.h, and the host-side mirror keeps the reference's error behaviour.
No compute is launched here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from strotss_tensorflow_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "strotss_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(strotss_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_exported_and_bound(lib):
    from strotss_tensorflow_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(declared)


def test_version_and_null_handle(lib):
    assert b"sm_100a" in lib.strotss_version()
    assert lib.strotss_last_error(None) == b"null handle"
    assert lib.strotss_workspace_bytes(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_create_fails_loudly_without_gpu(lib):
    h = ctypes.c_void_p()
    code = lib.strotss_create(0, ctypes.byref(h))
    assert code != 0
    assert h.value, "handle is returned so the error text can be read"
    assert len(lib.strotss_last_error(h)) > 0
    lib.strotss_destroy(h)


def test_plain_c_program_links_against_the_boundary(lib, tmp_path):
    """gcc-compiled C consumer of include/strotss_b200.h + libstrotss_b200.so (tests/c_abi/abi_smoke.c)."""
    import shutil
    import subprocess
    from strotss_tensorflow_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"),
           "-L", libdir, "-lstrotss_b200", "-Wl,-rpath," + libdir, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "ok" in run.stdout


def test_no_cpu_fallback_in_host_mirror():
    import strotss_tensorflow_b200 as S
    x = torch.rand(8, 5)
    with pytest.raises(RuntimeError, match="no CPU"):
        S.relaxed_emd(x, x)
    with pytest.raises(RuntimeError, match="no CPU"):
        S.self_similarity(x, x)
    with pytest.raises(RuntimeError, match="no CPU"):
        S.moment_matching(x, x)


def test_unknown_distance_is_keyerror_like_reference():
    import strotss_tensorflow_b200 as S
    x = torch.rand(4, 3)
    with pytest.raises(KeyError):
        S.relaxed_emd(x, x, distance="manhattan")
    assert set(S.dist_metrics) == {"cosine", "l2", "both"}


def test_reshape_2d_mirror():
    import strotss_tensorflow_b200 as S
    assert tuple(S.reshape_2d(torch.zeros(1, 5, 7)).shape) == (5, 7)
    assert tuple(S.reshape_2d(torch.zeros(2, 3, 4)).shape) == (6, 4)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "strotss_tensorflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} references the oracle"


def test_tf_adapter_calls_match_the_abi_signatures():
    """tf_adapter.py cannot be executed here (no TensorFlow); at least every C-ABI call it makes must exist in the
    header and pass as many arguments as the entry point takes."""
    import ast
    from strotss_tensorflow_b200 import _lib
    src = open(os.path.join(ROOT, "strotss_tensorflow_b200", "tf_adapter.py")).read()
    calls = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute)
             and n.func.attr.startswith("strotss_")]
    assert len(calls) >= 6
    for c in calls:
        name = c.func.attr
        assert name in _lib.SIGNATURES, f"{name} is not declared in include/strotss_b200.h"
        assert not c.keywords and len(c.args) == len(_lib.SIGNATURES[name][1]), \
            f"{name}: {len(c.args)} arguments passed, the entry point takes {len(_lib.SIGNATURES[name][1])}"


def test_ctypes_signatures_have_the_arity_of_the_header_prototypes():
    from strotss_tensorflow_b200 import _lib
    text = open(os.path.join(ROOT, "include", "strotss_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(strotss_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text)
    assert len(protos) == len(_lib.SIGNATURES)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), f"{name}: header has {n} parameters, ctypes binding {len(_lib.SIGNATURES[name][1])}"
