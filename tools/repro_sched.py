"""Run the decode-kernel schedule cases (forced ring depth / grid) one per subprocess; print the kernel diagnostics on failure."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(M, bits, stages, grid, reps):
    import torch
    from helpers import make_case, oracle_qgemm, rel_errors
    from flute_b200 import _lib, utils
    dev = torch.device("cuda", 0)
    ws = utils.get_workspace_streamk(dev)
    N, K = (2048, 1024) if bits == 2 else (3072, 2048)
    c = make_case(M, N, K, bits, 64, "bfloat16" if bits == 4 else "float16", seed=M)
    A = c["A"].to(dev); Q, S, t2 = (c[k].to(dev) for k in ("Q", "S", "table2"))
    code = _lib.BF16 if A.dtype == torch.bfloat16 else _lib.F16
    _lib.lib.flute_b200_set_variant(2)
    try:
        for r in range(reps):
            D = torch.full((M, N), float("nan"), dtype=A.dtype, device=dev)
            rc = _lib.lib.flute_b200_qgemm_debug(A.data_ptr(), Q.data_ptr(), D.data_ptr(), S.data_ptr(), t2.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), M, N, K, bits, 64, 32, code, 0, 0,
                                                 torch.cuda.current_stream().cuda_stream, 0, stages, grid, -1, None)
            _lib.check(rc)
            torch.cuda.synchronize()
        e = rel_errors(D.cpu(), oracle_qgemm(c))
        print(f"M={M} W{bits} stages={stages} grid={grid}: ok x{reps} rel err {e[0]:.2e}", flush=True)
    except Exception as e:
        rc = _lib.lib.flute_b200_check(0)
        print(f"M={M} W{bits} stages={stages} grid={grid}: FAIL {type(e).__name__} check rc={rc} "
              f"{_lib.lib.flute_b200_last_error().decode()}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*[int(x) for x in sys.argv[1:6]])
    else:
        for (M, bits) in ((1, 4), (2, 4), (3, 2)):
            for (stages, grid) in ((2, 37), (2, 148), (3, 37), (0, 37), (2, 5)):
                subprocess.run([sys.executable, __file__, str(M), str(bits), str(stages), str(grid), "10"], timeout=120)
