"""Worker of tests/test_tp_fused_gpu.py -- launched by torchrun, one process per GPU (world_size 2..8).

Column-shards a small chain of quantised linears, runs it with the exchange fused into the GEMM
(flute_b200.parallel.FusedGather -> flute_b200_qgemm_tp: NVLink peer stores of {value, sequence} words; arrival
counters only for the step's final plain image), eagerly and from a
CUDA graph, and checks every rank's gathered result against the same chain computed on that GPU alone with the
unsharded weights (flute_b200_qgemm), whose parity with the oracle the single-GPU tests establish.  Sharding is exact
(tests/test_tp_gloo.py), whole-tile outputs are deterministic, so the two must agree to fp32-reduction-order noise."""
import datetime
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from flute_b200 import _lib, parallel, utils
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    M, bits, group, layers = (int(os.environ.get("TP_TEST_M", "1")), 4, 64, 3)
    # N_total multiples of world * 128; tiles per rank: 1 partial (1024/8 = 128 columns) .. several
    shapes = [("up", 4096, 1024), ("down", 1024, 4096)]
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(99)          # identical on every rank
    table = torch.randn(16, generator=g).to(dtype)
    full = []
    for _ in range(layers):
        lin = {}
        for name, N, K in shapes:
            W = torch.randint(0, 16, (K, N), generator=g, dtype=torch.int64).to(torch.uint8)
            S = (torch.randn((N, K // group), generator=g) / K ** 0.5).to(dtype)
            lin[name] = (utils.pack_tile_p(W, bits, 32), S, N, K)
        full.append(lin)
    x0 = (torch.randn((M, shapes[0][2]), generator=g)).to(dtype).to(dev)
    table_d, table2_d = table.to(dev), utils.make_qmap2_from_qmap(table).to(dev)
    ws = utils.get_workspace_streamk(dev)
    st = lambda: torch.cuda.current_stream().cuda_stream

    # single-GPU reference chain with the unsharded weights
    full_d = [{k: (v[0].to(dev), v[1].to(dev), v[2], v[3]) for k, v in lin.items()} for lin in full]

    def ref_chain():
        x = x0
        for lin in full_d:
            for name, N, K in shapes:
                Q, S, _, _ = lin[name]
                D = torch.empty((M, N), dtype=dtype, device=dev)
                _lib.check(_lib.lib.flute_b200_qgemm(x.data_ptr(), Q.data_ptr(), D.data_ptr(), S.data_ptr(), table_d.data_ptr(),
                                                     table2_d.data_ptr(), ws.data_ptr(), ws.numel(), M, N, K, bits, group, 32,
                                                     _lib.BF16, 0, local, st()))
                x = D
        return x

    ref = ref_chain()
    torch.cuda.synchronize()

    shard = [{k: (*[t.to(dev) for t in parallel.shard_packed_linear(v[0], v[1], bits, rank, world, 32)], v[2] // world, v[3])
              for k, v in lin.items()} for lin in full]
    # every hand-over between qgemms rides on the {value, sequence} word image; the plain image is kept only where a
    # non-qgemm consumer reads (the final `down`), or everywhere with TP_TEST_COUNTERS=1
    all_plain = os.environ.get("TP_TEST_COUNTERS", "0") == "1"
    fg = parallel.FusedGather(dev, rank, world, [(name, M, N, layers) for name, N, K in shapes], dtype)
    flags = _lib.FLAG_PDL | _lib.FLAG_STATIC_WEIGHTS

    def tp_chain():
        fg.begin_step()
        x = x0
        for li, lin in enumerate(shard):
            for name, N, K in shapes:
                Q, S, n_loc, _ = lin[name]
                plain = True if all_plain else (name == "down" and li == layers - 1)
                x = fg.qgemm(x, Q, S, table_d, table2_d, ws, name, n_loc, K, bits, group, flags, plain=plain)
        fg.end_step(shapes[-1][0])
        return x.clone()

    def close(a, b):
        d = (a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)
        return float(d)

    errs = []
    for _ in range(3):
        out = tp_chain()
        torch.cuda.synchronize()
        errs.append(close(out, ref))
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            gout = tp_chain()
    torch.cuda.synchronize()
    dist.barrier()
    for _ in range(20):
        graph.replay()
    torch.cuda.synchronize()
    errs.append(close(gout, ref))
    _lib.check(_lib.lib.flute_b200_check(local))
    ok = all(e < 5e-3 for e in errs) and not torch.isnan(gout.float()).any()
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print(f"[tp_fused_worker rank {rank}/{world}] M={M} rel err vs single-GPU chain: {['%.2e' % e for e in errs]} -> {'ok' if ok else 'FAIL'}",
          flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
