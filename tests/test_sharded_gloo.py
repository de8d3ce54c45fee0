"""CPU, world_size 2 over gloo: the multi-GPU plumbing of ``mlvectordb_b200.sharded`` (row
partitioning, candidate all-gather layout, merge contract).  See ``tests/_gloo_worker.py``."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_gloo_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total_rows,k", [(1001, 10), (7, 5), (1, 3)])
def test_world2_gloo_sharded_search(total_rows, k):
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), "2", str(port), str(total_rows), "16", str(k), "cosine"],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT) for r in range(2)]
    for r, p in enumerate(procs):
        try:
            out, err = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, err[-3000:]
        assert f"rank {r} ok" in out


def test_world2_gloo_sharded_protocol_index():
    """``ShardedGpuIndex`` (the reference Index protocol, one process per GPU) call for call against a single-process
    ``GpuIndex`` -- see tests/_gloo_index_worker.py."""
    port = _free_port()
    worker = os.path.join(ROOT, "tests", "_gloo_index_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, cwd=ROOT) for r in range(2)]
    for r, p in enumerate(procs):
        try:
            out, err = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, err[-3000:]
        assert f"rank {r} ok" in out


def test_shard_range_partitions_rows():
    from mlvectordb_b200.sharded import shard_range
    for n in (0, 1, 7, 8, 9, 1000, 10_000_000):
        for g in (1, 2, 4, 8):
            blocks = [shard_range(n, r, g) for r in range(g)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(hi - lo for lo, hi in blocks) == -(-n // g)
