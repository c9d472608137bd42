"""Run the oracle-sweep cases one by one (each in a fresh subprocess) and print the kernel diagnostics on failure."""
import subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def one(bits, dtype, i, reps):
    import torch
    from helpers import make_case, oracle_qgemm, rel_errors
    from flute_b200 import _lib, utils
    dev = torch.device("cuda", 0)
    ws = utils.get_workspace_streamk(dev)
    N0 = 2048 if bits == 3 else 1024
    cases = [(1, N0, 512, 64, "randn"), (3, N0, 1024, 128, "arange"), (16, N0, 768, 256, "randn"),
             (32, 2 * N0, 256, 64, "randn"), (53, N0, 512, 64, "randn"), (64, N0, 512, 128, "randn"),
             (100, N0, 256, 64, "randn"), (1, 4096, 4096, 64, "nf4"), (7, N0 + N0 // 2, 3584, 128, "randn")]
    M, N, K, group, table = cases[i]
    if bits == 3 and N % 512: N = (N // 512) * 512
    c = make_case(M, N, K, bits, group, dtype, seed=i, table=table)
    A = c["A"].to(dev); Q, S, t2, tab = (c[k].to(dev) for k in ("Q", "S", "table2", "table"))
    code = _lib.BF16 if A.dtype == torch.bfloat16 else _lib.F16
    try:
        for r in range(reps):
            D = torch.full((M, N), float("nan"), dtype=A.dtype, device=dev)
            rc = _lib.lib.flute_b200_qgemm(A.data_ptr(), Q.data_ptr(), D.data_ptr(), S.data_ptr(), tab.data_ptr(), t2.data_ptr(),
                                           ws.data_ptr(), ws.numel(), M, N, K, bits, group, c["tile_p"], code, 0, 0,
                                           torch.cuda.current_stream().cuda_stream)
            _lib.check(rc)
        torch.cuda.synchronize()
        print(f"W{bits} {dtype} case {i} M={M} N={N} K={K} g={group}: ok x{reps}", flush=True)
    except Exception as e:
        rc = _lib.lib.flute_b200_check(0)
        print(f"W{bits} {dtype} case {i} M={M} N={N} K={K} g={group}: FAIL {type(e).__name__}: check rc={rc} "
              f"{_lib.lib.flute_b200_last_error().decode() if hasattr(_lib.lib.flute_b200_last_error(), 'decode') else _lib.lib.flute_b200_last_error()}", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    else:
        for bits in (2, 4):
            for dtype in ("float16", "bfloat16"):
                for i in (0, 1, 2, 7, 8):
                    subprocess.run([sys.executable, __file__, str(bits), dtype, str(i), "20"], timeout=120)
