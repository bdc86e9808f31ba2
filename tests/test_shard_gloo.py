"""N > 1 path on CPU: two gloo ranks each own half of the coarse cells; the union of
their volumes must equal the single-volume oracle bit for bit, with no data-path collective."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_shards_reproduce_the_whole_volume():
    from tests.emu import emu_py
    from oracle import oracle_py
    emu_py.build(); oracle_py.build()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SHARD_OK world=2" in r.stdout
