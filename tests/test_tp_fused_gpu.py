"""Fused tensor-parallel exchange on real GPUs (`gpurun --gpus 2 -- python -m pytest tests -m gpu -k tp_fused`; with one GPU
the same test runs at world_size 1).

Spawns tests/tp_fused_worker.py under torchrun; the worker checks every rank's gathered chain output (eager and CUDA-graph
replays) against the unsharded single-GPU chain."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("M,counters", [(1, 0), (3, 0), (2, 1)])
def test_tp_fused_exchange_matches_single_gpu(M, counters):
    n = torch.cuda.device_count()
    if n < 1:
        pytest.skip("needs a GPU")
    # one GPU: world_size 1 still runs the tensor-parallel kernel instantiation, the word-image hand-over between launches,
    # publish / wait and graph replay -- against this GPU's own symmetric-memory buffers instead of a peer's
    world = 8 if n >= 8 else 4 if n >= 4 else 2 if n >= 2 else 1
    env = dict(os.environ, TP_TEST_M=str(M), TP_TEST_COUNTERS=str(counters))
    port = 29700 + (os.getpid() % 200) + M + 7 * counters
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "tp_fused_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-4000:])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
