"""Tensor-parallel column sharding over a real process group (gloo, world_size 2, CPU).

Each rank takes its row slice of the packed weight, computes its [M, N/tp] output slice with the
ORACLE (this is a host-logic test: no GPU here), all-gathers, and the result must equal the oracle's
full-width output bit for bit -- i.e. sharding + one all-gather is exact."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, bits, M, out_q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flute_b200 import parallel, utils
        from oracle import c_oracle
        N, K, group = (2048 if bits == 3 else 1024), 256, 64
        g = torch.Generator().manual_seed(7)
        W = torch.randint(0, 2 ** bits, (K, N), generator=g, dtype=torch.int64).to(torch.uint8)
        S = torch.randn((N, K // group), generator=g).to(torch.float16)
        A = (torch.randn((M, K), generator=g) / 100).to(torch.float16)
        table = torch.randn(2 ** bits, generator=g).to(torch.float16)
        table2 = utils.make_qmap2_from_qmap(table)
        Q = utils.pack_tile_p(W, bits, 32)
        b16 = lambda t: t.contiguous().view(torch.int16).numpy().view(np.uint16)
        Qr, Sr = parallel.shard_packed_linear(Q, S, bits, rank, world, 32)
        D_local = c_oracle.qgemm(b16(A), Qr.numpy(), b16(Sr), table2.numpy(), bits, group, False)
        D_local = torch.from_numpy(D_local.view(np.int16)).view(torch.float16)
        D = parallel.all_gather_columns(D_local)
        D_full = torch.from_numpy(c_oracle.qgemm(b16(A), Q.numpy(), b16(S), table2.numpy(), bits, group, False).view(np.int16)).view(torch.float16)
        ok = bool(torch.equal(D.view(torch.int16), D_full.view(torch.int16))) and D.shape == (M, N)
        out_q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bits,M", [(4, 1), (4, 3), (3, 2), (2, 1)])
def test_tp2_allgather_equals_full(bits, M):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + bits * 7 + M
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bits, M, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=5) for _ in range(2))
    assert results == {0: True, 1: True}
