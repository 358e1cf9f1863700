"""GPU: the sharded path (b200_set_shard: targets split over ranks, results all-gathered) gives
bit-identical particle state to the single-rank path.  Two ranks share cuda:0 and stage the
collective through the host (gloo) so the test runs on a one-GPU box; with >= 2 GPUs the same
runner is also launched over NCCL."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
N = 30000


def _launch(world, out, backend, min_work="1", compact="1", types="0", n=N):
    runner = os.path.join(HERE, "multi_runner.py")
    if world == 1:
        cmd = [sys.executable, runner, out, str(n), backend]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", "29541", runner, out, str(n), backend]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, B200_SHARD_MIN_WORK=min_work, B200_COMPACT=compact, B200_TYPES=types))
    assert r.returncode == 0, (r.stdout[-1500:] + "\n".join(l for l in r.stderr.splitlines() if "Error" in l or "assert" in l or "File" in l)[-3000:])
    return np.load(out)


def _same(a, b):
    for k in a.files:
        assert np.array_equal(a[k], b[k]), f"{k} differs between 1 rank and 2 ranks"


def test_two_ranks_equal_one_rank(tmp_path):
    one = _launch(1, str(tmp_path / "one.npz"), "gloo")
    assert one["sct0"][2] + one["sct1"][2] > 0, "no scatterings - fixture too quiet"
    two = _launch(2, str(tmp_path / "two.npz"), "gloo")
    _same(one, two)


def test_two_ranks_three_types(tmp_path):
    """three particle types (one tree per type, query groups and 32-aligned leaf blocks per tree) sharded over two ranks"""
    one = _launch(1, str(tmp_path / "one.npz"), "gloo", types="1")
    assert one["sct0"][2] + one["sct1"][2] > 0, "no scatterings - fixture too quiet"
    two = _launch(2, str(tmp_path / "two.npz"), "gloo", types="1")
    _same(one, two)


def test_two_ranks_dense_exchange(tmp_path):
    """the 32-byte SlotRec exchange (fallback of the compact one)"""
    one = _launch(1, str(tmp_path / "one.npz"), "gloo")
    two = _launch(2, str(tmp_path / "two.npz"), "gloo", compact="0")
    _same(one, two)


def test_two_ranks_default_threshold(tmp_path):
    """with the default "shard_min_work" a problem this small is done completely by every rank (no exchanges
    in the compute path); the sharded host-array upload / download still run"""
    one = _launch(1, str(tmp_path / "one.npz"), "gloo")
    two = _launch(2, str(tmp_path / "two.npz"), "gloo", min_work=str(1 << 18))
    _same(one, two)


def test_two_gpus_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    one = _launch(1, str(tmp_path / "one.npz"), "nccl")
    two = _launch(2, str(tmp_path / "two.npz"), "nccl")
    _same(one, two)


@pytest.mark.parametrize("world", [2, 4])
def test_nccl_default_threshold_large(tmp_path, world):
    """NCCL over NVLink with everything at its defaults (shard_min_work = 2^18, compact exchange, SIDM chain on its own
    stream next to the walk = shard_overlap): the all-active passes of 3e5 particles are sharded, the repair passes
    (< 2^18 particles) run completely on every rank.  N-GPU state = 1-GPU state, bit for bit."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    big = 300000
    one = _launch(1, str(tmp_path / "one.npz"), "nccl", min_work=str(1 << 18), n=big)
    assert one["sct0"][2] + one["sct1"][2] > 0 and one["sct0"][4] > 0, "no scatterings / no repair pass - fixture too quiet"
    many = _launch(world, str(tmp_path / "many.npz"), "nccl", min_work=str(1 << 18), n=big)
    _same(one, many)
