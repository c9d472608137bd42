"""Host time per eager qgemm call through each binding (verdict item 7): compiled TORCH_LIBRARY shim vs the torch.library
Python implementation vs the raw C ABI through ctypes.  Wall clock of N back-to-back calls that never wait for the GPU
(the kernel, ~8 us, is shorter than any of these, so the loop is host-bound; a final sync is outside the clock).

    python tools/host_overhead.py          # prints us per call for the binding this process loaded
    FLUTE_B200_PY_OPS=1 python tools/host_overhead.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import flute_b200 as flute
    from flute_b200 import _lib, ops, utils
    from flute_b200.templates import default_template_id
    dev = torch.device("cuda", 0)
    ws = utils.get_workspace_streamk(dev)
    N = K = 4096
    Q = torch.randint(-32768, 32768, (N // 4, K), dtype=torch.int16, device=dev)
    S = torch.randn((N, K // 64), device=dev).to(torch.bfloat16)
    table = torch.randn(16, device=dev).to(torch.bfloat16)
    t2 = utils.make_qmap2_from_qmap(table)
    x = torch.randn((1, K), device=dev).to(torch.bfloat16)
    D = torch.empty((1, N), dtype=torch.bfloat16, device=dev)
    tid = default_template_id(4)
    st = torch.cuda.current_stream().cuda_stream

    def op():
        return flute.qgemm(x, Q, S, table, t2, ws, 4, 64, tid, 148)

    def cabi():
        return _lib.lib.flute_b200_qgemm(x.data_ptr(), Q.data_ptr(), D.data_ptr(), S.data_ptr(), table.data_ptr(), t2.data_ptr(),
                                         ws.data_ptr(), ws.numel(), 1, N, K, 4, 64, 32, _lib.BF16, 0, 0, st)

    for name, fn in ((f"torch.ops.flute.qgemm_raw_simple [{ops.BINDING} binding]", op), ("flute_b200_qgemm via ctypes", cabi)):
        for _ in range(200):
            fn()
        torch.cuda.synchronize()
        n = 5000
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        dt = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"{name:62s} {dt / n * 1e6:7.2f} us per eager call (M=1, 4096x4096, {n} calls)", flush=True)


if __name__ == "__main__":
    main()
