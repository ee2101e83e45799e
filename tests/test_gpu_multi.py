"""Row-sharded evaluation on >= 2 GPUs against the single-GPU evaluation (run with -m gpu on a multi-GPU box; skipped on
one GPU).  Launches tests/multi_gpu_check.py under torch.distributed.run: every rank evaluates the same problems twice --
sharded over all ranks and alone -- and compares scalars, argmins, its own gradient rows and the all-gathered gradient."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_matches_single_gpu(cuda_device):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the driver's single-GPU box runs the same check inside bench.py --gpus N)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-3000:] + out.stderr[-3000:])
    assert "multi-GPU parity: OK" in out.stdout
