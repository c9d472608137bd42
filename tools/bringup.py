"""GPU bring-up harness: runs groups of small cases against the oracle, each group in its own
subprocess under a timeout, so a trap or hang in one group cannot take the others (or the box) down.

    python tools/bringup.py                 # all groups, report to stdout + gpurun_out/bringup.log
    python tools/bringup.py --group tmem    # one group in-process
"""
from __future__ import annotations

import argparse
import ctypes
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def T(dtype):
    import torch
    return torch.float16 if dtype == "float16" else torch.bfloat16


def to_np16(t):
    import torch
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def make_case(M, N, K, bits, group, dtype, seed=0, table_kind="randn", tile_p=32, identity=False):
    """Inputs built on CPU with numpy/torch; packed with the engine's own packer (golden-pinned)."""
    import torch
    from flute_b200 import utils
    g = torch.Generator().manual_seed(seed)
    t = T(dtype)
    if identity:
        A = torch.eye(K, dtype=t)[:M]
    else:
        A = (torch.randn((M, K), generator=g) / 100.).to(t)
    W = torch.randint(0, 2 ** bits, (K, N), generator=g, dtype=torch.int64).to(torch.uint8)
    S = torch.randn((N, K // group), generator=g).to(t)
    if table_kind == "arange":
        table = torch.arange(2 ** bits).to(t)
    else:
        table = torch.randn(2 ** bits, generator=g).to(t)
    table2 = utils.make_qmap2_from_qmap(table)
    Q = utils.pack_tile_p(W, bits, tile_p)
    return dict(A=A, W=W, S=S, table=table, table2=table2, Q=Q)


def oracle_out(c, bits, group, dtype, tile_p=32):
    from oracle import c_oracle
    D = c_oracle.qgemm(to_np16(c["A"]), c["Q"].numpy(), to_np16(c["S"]), c["table2"].numpy(), bits, group,
                       dtype == "bfloat16", tile_p)
    return D


def rel_err(D, Dref, dtype):
    from oracle import flute_oracle as O
    a = O.to_f32(D.view(np.float16) if dtype == "float16" else D, dtype).astype(np.float64)
    b = O.to_f32(Dref.view(np.float16) if dtype == "float16" else Dref, dtype).astype(np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run_qgemm(c, M, N, K, bits, group, dtype, tile_p=32, force_mb=0, force_stages=0, force_grid=0, force_streamk=-1,
              dbg=False, workspace=None):
    import torch
    from flute_b200 import _lib, utils
    dev = torch.device("cuda", 0)
    A, Q, S, t2 = (c[k].to(dev) for k in ("A", "Q", "S", "table2"))
    D = torch.full((M, N), float("nan"), dtype=A.dtype, device=dev)
    ws = workspace if workspace is not None else utils.get_workspace_streamk(dev)
    dbgbuf = torch.zeros((128, 128), dtype=torch.int32, device=dev) if dbg else None
    rc = _lib.lib.flute_b200_qgemm_debug(
        A.data_ptr(), Q.data_ptr(), D.data_ptr(), S.data_ptr(), t2.data_ptr(), ws.data_ptr(), ws.numel(),
        M, N, K, bits, group, tile_p, 0 if dtype == "float16" else 1, 0, 0,
        torch.cuda.current_stream().cuda_stream, force_mb, force_stages, force_grid, force_streamk,
        dbgbuf.data_ptr() if dbg else None)
    _lib.check(rc)
    try:
        torch.cuda.synchronize()
    except Exception as e:   # trapped kernel: fetch the reason
        try:
            _lib.check(_lib.lib.flute_b200_check(0))
        except RuntimeError as e2:
            raise RuntimeError(f"{e2}") from e
        raise
    return to_np16(D), (dbgbuf.cpu().numpy().view(np.uint32) if dbg else None), ws


def report(name, D, Dref, dtype, tol):
    e = rel_err(D, Dref, dtype)
    nbad = int((D != Dref).sum())
    nan = int(np.isnan((D.view(np.float16) if dtype == "float16" else (D.astype(np.uint32) << 16).view(np.float32)).astype(np.float32)).sum())
    ok = e <= tol and nan == 0
    print(f"  [{'OK ' if ok else 'BAD'}] {name}: rel={e:.3e} mismatched={nbad}/{D.size} nan={nan}", flush=True)
    if not ok:
        bad = np.argwhere(D != Dref)
        rows = np.unique(bad[:, 0])[:8]
        cols = np.unique(bad[:, 1])
        print(f"        bad rows(m) {rows.tolist()} ; bad cols: count {cols.size} first {cols[:16].tolist()} "
              f"mod32 {np.unique(cols % 32)[:8].tolist()} //32%4 {np.unique((cols // 32) % 4).tolist()} //128 {np.unique(cols // 128)[:8].tolist()}")
        for (m, n) in bad[:6]:
            print(f"        D[{m},{n}] = {D[m, n]:#06x} ref {Dref[m, n]:#06x}")
    return ok


# ------------------------------------------------------------------------------------------------
def group_aux():
    """dequantize + hadamard kernels vs the oracle."""
    import torch
    from flute_b200 import utils, ops
    from oracle import c_oracle
    ok = True
    for bits, tp in [(4, 32), (4, 64), (2, 32), (2, 64), (3, 32)]:
        for dtype in ("float16", "bfloat16"):
            N = 1024 if bits != 3 else 1024
            K, group = 256, 64
            c = make_case(1, N, K, bits, group, dtype, seed=bits, tile_p=tp)
            What = utils.dequantize(c["Q"].cuda(), c["S"].cuda(), c["table2"].cuda(), bits, group, tp)
            torch.cuda.synchronize()
            ref = c_oracle.dequantize(c["Q"].numpy(), to_np16(c["S"]), c["table2"].numpy(), bits, group, dtype == "bfloat16", tp)
            same = bool((to_np16(What) == ref).all())
            print(f"  [{'OK ' if same else 'BAD'}] dequantize b{bits} tp{tp} {dtype}", flush=True)
            ok &= same
    for h in (2, 64, 512, 2048, 32768):
        for dtype in ("float16", "bfloat16"):
            x = (torch.randn((3, max(h, 4096) // h * h)) ).to(T(dtype))
            y = ops.hadamard_transform(x.cuda(), h)
            torch.cuda.synchronize()
            ref = c_oracle.hadamard(to_np16(x), h, dtype == "bfloat16")
            e = rel_err(to_np16(y), ref, dtype)
            good = e < (2e-3 if dtype == "float16" else 1.2e-2)
            print(f"  [{'OK ' if good else 'BAD'}] hadamard h={h} {dtype} rel={e:.2e}", flush=True)
            ok &= good
    return ok


def group_tmem():
    """Single stage, single CTA: check the dequantised TMEM chunk, then the MMA result."""
    from oracle import c_oracle
    from flute_b200 import _lib
    _lib.lib.flute_b200_set_variant(0)   # the expected chunk layout below is the LARGE footprint's
    ok = True
    for dtype in ("float16", "bfloat16"):
        bits, group, M, N, K = 4, 64, 1, 512, 64
        c = make_case(M, N, K, bits, group, dtype, seed=1)
        D, dbg, _ = run_qgemm(c, M, N, K, bits, group, dtype, dbg=True)
        What = c_oracle.dequantize(c["Q"].numpy(), to_np16(c["S"]), c["table2"].numpy(), bits, group, dtype == "bfloat16", 32)
        # expected chunk: lane L, column j*32 + k2  -> pair (What[2k2, n], What[2k2+1, n]), n = (L//32)*128 + j*32 + L%32
        exp = np.zeros((128, 128), dtype=np.uint32)
        for L in range(128):
            for j in range(4):
                n = (L // 32) * 128 + j * 32 + (L % 32)
                lo = What[0::2, n].astype(np.uint32)
                hi = What[1::2, n].astype(np.uint32)
                exp[L, j * 32:(j + 1) * 32] = lo | (hi << 16)
        same = bool((dbg == exp).all())
        print(f"  [{'OK ' if same else 'BAD'}] TMEM chunk {dtype}: mismatches {(dbg != exp).sum()}", flush=True)
        if not same:
            bad = np.argwhere(dbg != exp)[:6]
            for (L, col) in bad:
                print(f"        chunk[{L},{col}] = {dbg[L, col]:#010x} exp {exp[L, col]:#010x}")
        ok &= same
        ok &= report(f"qgemm M=1 N=512 K=64 {dtype}", D, oracle_out(c, bits, group, dtype), dtype, 2e-3 if dtype == "float16" else 1.1e-2)
    return ok


def _sweep(cases, tolmul=1.0):
    ok = True
    for (name, M, N, K, bits, group, dtype, kw) in cases:
        mk = {k: kw.pop(k) for k in ("table_kind", "tile_p", "identity", "seed") if k in kw}
        c = make_case(M, N, K, bits, group, dtype, **mk)
        tp = mk.get("tile_p", 32)
        try:
            D, _, _ = run_qgemm(c, M, N, K, bits, group, dtype, tile_p=tp, **kw)
        except Exception as e:
            print(f"  [BAD] {name}: EXCEPTION {e}", flush=True)
            return False
        tol = 0.0 if mk.get("identity") else (2e-3 if dtype == "float16" else 1.1e-2) * tolmul
        ok &= report(name, D, oracle_out(c, bits, group, dtype, tp), dtype, tol)
    return ok


def group_w4():
    cs = []
    for dtype in ("float16", "bfloat16"):
        cs += [
            (f"w4 K=128 1cta {dtype}", 1, 512, 128, 4, 64, dtype, {}),
            (f"w4 K=1024 N=512 (stages wrap) {dtype}", 3, 512, 1024, 4, 64, dtype, {}),
            (f"w4 N=4096 K=4096 M=1 {dtype}", 1, 4096, 4096, 4, 64, dtype, {}),
            (f"w4 N=4096 K=4096 M=16 g128 {dtype}", 16, 4096, 4096, 4, 128, dtype, {}),
            (f"w4 N=1024 K=4096 M=5 g256 {dtype}", 5, 1024, 4096, 4, 256, dtype, {}),
            (f"w4 M=17 (mb 32) {dtype}", 17, 1024, 512, 4, 64, dtype, {}),
            (f"w4 M=53 (mb 64) {dtype}", 53, 1024, 512, 4, 64, dtype, {}),
            (f"w4 M=100 (2 m-tiles) {dtype}", 100, 1024, 512, 4, 64, dtype, {}),
            (f"w4 N=640 (partial n-tile) {dtype}", 7, 640, 256, 4, 64, dtype, {}),
            (f"w4 tile_P=64 {dtype}", 4, 1024, 256, 4, 64, dtype, {"tile_p": 64}),
            (f"w4 identity K=256 {dtype}", 256, 512, 256, 4, 64, dtype, {"identity": True}),
            (f"w4 identity arange {dtype}", 128, 512, 128, 4, 64, dtype, {"identity": True, "table_kind": "arange"}),
            (f"w4 streamk grid=3 {dtype}", 2, 1024, 512, 4, 64, dtype, {"force_grid": 3, "force_streamk": 1}),
            (f"w4 streamk grid=7 stages=2 {dtype}", 2, 1536, 1024, 4, 64, dtype, {"force_grid": 7, "force_streamk": 1, "force_stages": 2}),
            (f"w4 data-parallel grid=2 {dtype}", 2, 2048, 256, 4, 64, dtype, {"force_grid": 2, "force_streamk": 0}),
            (f"w4 K=3584 g128 (G=28) {dtype}", 2, 512, 3584, 4, 128, dtype, {}),
        ]
    return _sweep(cs)


def group_w2():
    cs = []
    for dtype in ("float16", "bfloat16"):
        cs += [
            (f"w2 K=64 {dtype}", 1, 1024, 64, 2, 64, dtype, {}),
            (f"w2 K=1024 M=3 {dtype}", 3, 2048, 1024, 2, 64, dtype, {}),
            (f"w2 N=4096 K=4096 M=20 g128 {dtype}", 20, 4096, 4096, 2, 128, dtype, {}),
            (f"w2 tile_P=64 {dtype}", 4, 1024, 256, 2, 64, dtype, {"tile_p": 64}),
            (f"w2 identity {dtype}", 128, 1024, 128, 2, 64, dtype, {"identity": True}),
            (f"w2 streamk grid=5 {dtype}", 2, 2048, 512, 2, 64, dtype, {"force_grid": 5, "force_streamk": 1}),
        ]
    return _sweep(cs)


def group_w3():
    cs = []
    for dtype in ("float16", "bfloat16"):
        cs += [
            (f"w3 K=64 {dtype}", 1, 2048, 64, 3, 64, dtype, {}),
            (f"w3 K=1024 M=3 {dtype}", 3, 2048, 1024, 3, 64, dtype, {}),
            (f"w3 N=4096 K=4096 M=16 g128 {dtype}", 16, 4096, 4096, 3, 128, dtype, {}),
            (f"w3 N=512 (partial tile) {dtype}", 2, 512, 256, 3, 64, dtype, {}),
            (f"w3 identity {dtype}", 128, 1024, 128, 3, 64, dtype, {"identity": True}),
            (f"w3 streamk grid=5 M=40 {dtype}", 40, 4096, 512, 3, 64, dtype, {"force_grid": 5, "force_streamk": 1}),
        ]
    return _sweep(cs)


def group_api():
    """Through the public Python API (torch op), incl. workspace reuse across calls."""
    import torch
    import flute_b200 as flute
    from flute_b200 import utils
    ok = True
    dev = torch.device("cuda", 0)
    ws = utils.get_workspace_streamk(dev)
    for rep in range(3):
        for bits in (4, 3, 2):
            dtype = "bfloat16"
            M, N, K, group = 3, 4096, 4096, 64
            c = make_case(M, N, K, bits, group, dtype, seed=rep * 10 + bits)
            out = flute.qgemm(c["A"].to(dev).view(1, M, K), c["Q"].to(dev), c["S"].to(dev), c["table"].to(dev),
                              c["table2"].to(dev), ws, bits, group, flute.templates.default_template_id(bits), 148)
            torch.cuda.synchronize()
            assert out.shape == (1, M, N)
            ok &= report(f"api rep{rep} w{bits}", to_np16(out.view(M, N)), oracle_out(c, bits, group, dtype), dtype, 1.1e-2)
    print("  workspace counters clean:", bool((ws[:65536] == 0).all().item()))
    return ok


GROUPS = {"aux": group_aux, "tmem": group_tmem, "w4": group_w4, "w2": group_w2, "w3": group_w3, "api": group_api}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--only", default=None, help="comma-separated groups for the driver mode")
    args = ap.parse_args()
    if args.group:
        from flute_b200 import _lib
        _lib.lib.flute_b200_set_timeout_ms(3000)
        if os.environ.get("FLUTE_B200_VARIANT"):
            _lib.lib.flute_b200_set_variant(int(os.environ["FLUTE_B200_VARIANT"]))
        ok = GROUPS[args.group]()
        print(f"GROUP {args.group}: {'PASS' if ok else 'FAIL'}", flush=True)
        sys.exit(0 if ok else 1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    names = args.only.split(",") if args.only else list(GROUPS)
    results = {}
    with open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "w") as log:
        for g in names:
            t0 = time.time()
            try:
                pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--group", g], cwd=ROOT,
                                    stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240)
                out, rc = pr.stdout, pr.returncode
            except subprocess.TimeoutExpired as e:
                out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), -999
            results[g] = rc
            hdr = f"===== group {g}: rc={rc} ({time.time() - t0:.1f}s) ====="
            print(hdr)
            print(out[-6000:])
            log.write(hdr + "\n" + out + "\n")
            log.flush()
    print("SUMMARY", results)
    sys.exit(0 if all(v == 0 for v in results.values()) else 1)


if __name__ == "__main__":
    main()
