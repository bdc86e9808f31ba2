"""Multi-GPU data paths on real devices (needs >= 2 GPUs; skipped otherwise): row-sliced upload + NVLink all-gather,
sharded integration, device-to-device shard gather — the gathered volume must equal the CPU oracle bit for bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_rows_allgather_and_device_gather_match_the_oracle(world, engine_lib, oracle_lib):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29530 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"MGPU_OK world={world}" in r.stdout
