"""Multi-GPU parity: row shards on 2 (and 4/8 when present) GPUs, one process per GPU, vs the
unsharded search on one GPU -- bit-identical rows, scores and counts.  Covers the fused
peer-memory exchange (csrc/exchange.cuh) and the NCCL all-gather + merge path.  Needs >= 2 GPUs
(`gpurun --gpus 2`); on a single-GPU box there is nothing to shard over and the test is skipped."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_nccl_worker.py")


def _n_gpus():
    from mlvectordb_b200 import _capi
    return _capi.lib().mlv_device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_search_equals_unsharded(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, cwd=ROOT) for r in range(world)]
    for r, p in enumerate(procs):
        try:
            out, err = p.communicate(timeout=420)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, f"rank {r}:\n{err[-3000:]}"
        assert f"rank {r} ok" in out
