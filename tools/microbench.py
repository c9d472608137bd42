"""Per-shape micro-benchmark of the qgemm kernel (not the judged bench; a tuning tool).

For each (N, K, M): enough distinct weight copies to exceed L2 several times are cycled inside one
CUDA graph, timed with CUDA events; prints achieved GB/s (algorithmic bytes) and TFLOP/s.

    python tools/microbench.py --M 1,16 --shapes llama8b [--bits 4] [--dtype bf16] [--pdl 1]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPE_SETS = {
    "llama8b": [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)],
    "llama70b": [(10240, 8192), (8192, 8192), (57344, 8192), (8192, 28672)],
    "small": [(4096, 4096)],
    "gateup": [(28672, 4096)],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", default="1")
    ap.add_argument("--shapes", default="llama8b")
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--group", type=int, default=64)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--force-mb", type=int, default=0)
    ap.add_argument("--force-stages", type=int, default=0)
    ap.add_argument("--force-grid", default="0", help="CTAs per launch (0 = engine's choice); a comma list sweeps")
    ap.add_argument("--force-streamk", type=int, default=-1)
    ap.add_argument("--json", default=None)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--ablate", type=int, default=0)
    ap.add_argument("--slack", type=int, default=0, help="SMs left idle per launch (grid = SMs - slack)")
    ap.add_argument("--l2pf", type=int, default=-1, help="decode kernel: L2 prefetch distance in stages (-1 = engine's choice)")
    ap.add_argument("--trace", type=int, default=0, help="print a per-CTA timeline of one isolated launch")
    args = ap.parse_args()

    import torch
    from flute_b200 import _lib, utils
    dev = torch.device("cuda", 0)
    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    ws = utils.get_workspace_streamk(dev)
    _lib.lib.flute_b200_set_variant((args.variant & 0xff) | (args.ablate << 8) | ((args.l2pf + 1) << 16) | (args.slack << 24))
    bits, group = args.bits, args.group
    table = torch.randn(2 ** bits, device=dev).to(dt)
    table2 = utils.make_qmap2_from_qmap(table)
    results = []
    shapes = SHAPE_SETS[args.shapes] if args.shapes in SHAPE_SETS else [tuple(int(v) for v in t.split("x")) for t in args.shapes.split(",")]
    grids = [int(g) for g in str(args.force_grid).split(",")]
    for (N, K) in shapes:
      for force_grid in grids:
        for M in [int(m) for m in args.M.split(",")]:
            wbytes = N * K * bits // 8 + N * (K // group) * 2
            abytes = wbytes + M * K * 2 + M * N * 2 + (2 ** bits) * 2 + (4 ** bits) * 4
            ncopies = max(2, min(768, (600 * 2 ** 20 + wbytes - 1) // wbytes))
            Qs = [torch.randint(-32768, 32768, (N // 16 * bits, K), dtype=torch.int16, device=dev) for _ in range(ncopies)]
            Ss = [(torch.randn((N, K // group), device=dev) / K ** 0.5).to(dt) for _ in range(ncopies)]
            A = (torch.randn((M, K), device=dev)).to(dt)
            D = torch.empty((M, N), dtype=dt, device=dev)
            flags = (_lib.FLAG_PDL | _lib.FLAG_STATIC_WEIGHTS) if args.pdl else 0

            def launch(i):
                rc = _lib.lib.flute_b200_qgemm_debug(
                    A.data_ptr(), Qs[i].data_ptr(), D.data_ptr(), Ss[i].data_ptr(), table2.data_ptr(), ws.data_ptr(),
                    ws.numel(), M, N, K, bits, group, 32, 1 if dt == torch.bfloat16 else 0, flags, 0,
                    torch.cuda.current_stream().cuda_stream, args.force_mb, args.force_stages, force_grid,
                    args.force_streamk, None)
                _lib.check(rc)

            for i in range(ncopies):
                launch(i)
            torch.cuda.synchronize()
            if args.trace and os.environ.get("FLUTE_B200_PROFILE") != "1":
                print("   --trace needs the profiling build: FLUTE_B200_PROFILE=1 (python flute_b200/build.py --profile)")
            elif args.trace:
                # three PDL-chained launches, one trace buffer each: shows how far kernel i+1's CTAs get (start, set-up)
                # while kernel i still runs, and the gap between kernel i's last exit and kernel i+1's first epilogue
                import numpy as np
                nchain = 3
                trs = [torch.zeros((512, 48), dtype=torch.int64, device=dev) for _ in range(nchain)]
                torch.cuda.synchronize()
                # captured in ONE CUDA graph (eager launches through ctypes are ~10 us apart: host-bound, no overlap to see);
                # the trace pointer is a kernel argument, fixed at capture time
                tg = torch.cuda.CUDAGraph()
                tside = torch.cuda.Stream()
                with torch.cuda.stream(tside):
                    with torch.cuda.graph(tg, stream=tside):
                        launch(0)
                        for j in range(nchain):
                            _lib.lib.flute_b200_set_trace_buffer(trs[j].data_ptr())
                            launch((1 + j) % ncopies)
                        _lib.lib.flute_b200_set_trace_buffer(None)
                        launch((1 + nchain) % ncopies)
                torch.cuda.synchronize()
                for _ in range(2):
                    tg.replay()
                torch.cuda.synchronize()
                names = ["start", "setup", "dep-ready", "mma-first", "acc-full", "epilogue", "last-epi", "exit"]
                t0 = None
                for j in range(nchain):
                    t = trs[j].cpu().numpy()
                    t = t[t[:, 0] > 0]
                    if t0 is None:
                        t0 = t[:, 0].min()
                    print(f"   trace launch {j} N={N} K={K} M={M}: {t.shape[0]} CTAs; ns since the first CTA of launch 0 started (min / median / max)")
                    for c, nm in enumerate(names):
                        col = t[:, c]; col = col[col > 0] - t0
                        if col.size:
                            print(f"     {nm:10s} {col.min():8d} {int(np.median(col)):8d} {col.max():8d}   (n={col.size})")
                    if args.trace >= 2 and j == nchain - 1:
                        # the slowest CTAs of the last launch: where their time goes (last partial segment: reds+bar, atomic, finalize)
                        tt = trs[j].cpu().numpy()
                        idx = np.argsort(-tt[:, 6])[:8]
                        print("     slowest CTAs (ns since launch-0 start): cta start dep-ready mma-first acc-full last-epi | last partial segment: reds-done atom-done fin-done seg-offset contributors last")
                        live = np.nonzero(tt[:, 0] > 0)[0]
                        mid = live[np.argsort(tt[live, 6])][len(live) // 2: len(live) // 2 + 3]
                        for b in list(idx) + [int(x) for x in mid]:
                            if tt[b, 0] == 0:
                                continue
                            r = lambda c: int(tt[b, c] - t0) if tt[b, c] > 0 else -1
                            f = int(tt[b, 47])
                            print(f"       {b:4d} {r(0):7d} {r(2):7d} {r(3):7d} {r(4):7d} {r(6):7d} | {r(44):7d} {r(45):7d} {r(46):7d} seg+{(f >> 8) & 0xfff} c={f >> 20} last={f & 1}")
                if os.environ.get("FLUTE_B200_PROFILE") == "1" and M <= 16 and bits in (2, 4) and (args.variant < 0 or args.variant >= 2):
                    prof = [(8, "producer scale blocks"), (9, "producer wait-empty"), (10, "producer W issue"),
                            (12, "producer iters"), (13, "mma wait-full"), (14, "mma wait-afull"), (15, "mma wait-pempty"),
                            (16, "mma issue+commit"),
                            (24, "dq0 wait-full"), (25, "dq0 wait-aempty"), (26, "dq0 pieces"), (27, "dq0 wait-st+arrive"),
                            (32, "dq5 wait-full"), (33, "dq5 wait-aempty"), (34, "dq5 pieces"), (35, "dq5 wait-st+arrive"),
                            (40, "apply scale block"), (41, "apply wait-pfull"), (42, "apply ld+fma"), (43, "apply epilogue")]
                    for c, nm in prof:
                        col = t[:, c]
                        print(f"     {nm:26s} min {col.min():8d} med {int(np.median(col)):8d} max {col.max():8d}")
                elif os.environ.get("FLUTE_B200_PROFILE") == "1" and bits == 4 and args.variant < 0:
                    prof = [(8, "producer scale blocks"), (9, "producer wait-empty"), (10, "producer issue"), (12, "producer iters"),
                            (13, "mma wait-accempty"), (14, "mma wait-afull"), (16, "mma issue+commit"),
                            (24, "dq0 wait-full"), (25, "dq0 wait-aslot"), (26, "dq0 pieces"), (27, "dq0 wait-st+arrive"), (28, "dq0 scales+loop"),
                            (32, "dq9 wait-full"), (33, "dq9 wait-aslot"), (34, "dq9 pieces"), (35, "dq9 wait-st+arrive"), (36, "dq9 scales+loop"),
                            (29, "dq0 epilogue wait"), (30, "dq0 epilogue work"), (37, "dq9 epilogue wait"), (38, "dq9 epilogue work")]
                    for c, nm in prof:
                        col = t[:, c]
                        print(f"     {nm:26s} min {col.min():8d} med {int(np.median(col)):8d} max {col.max():8d}")
                elif os.environ.get("FLUTE_B200_PROFILE") == "1":
                    prof = [(8, "producer wait-empty"), (9, "producer issue"), (10, "producer iters"), (11, "mma wait-full"),
                            (12, "mma wait-afull"), (13, "mma issue+commit"), (14, "mma wait-accempty"),
                            (40, "scale wait-empty"), (41, "scale load+store"), (42, "scale chunks"),
                            (16, "dq0 loop (no waits)"), (23, "dq0 wait-scale"), (17, "dq0 wait-full"), (18, "dq0 wait-aempty"),
                            (19, "dq0 run"), (20, "dq0 wait-st+arrive"), (21, "dq0 epilogue"), (22, "dq0 chunks"),
                            (24, "dq1 loop (no waits)"), (31, "dq1 wait-scale"), (25, "dq1 wait-full"), (26, "dq1 wait-aempty"),
                            (27, "dq1 run"), (28, "dq1 wait-st+arrive"), (29, "dq1 epilogue"), (30, "dq1 chunks")]
                    for c, nm in prof:
                        col = t[:, c]
                        print(f"     {nm:22s} min {col.min():8d} med {int(np.median(col)):8d} max {col.max():8d}")
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    for i in range(ncopies):
                        launch(i)
            torch.cuda.synchronize()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (args.reps * ncopies)
            gbs = abytes / us / 1e3
            tfl = 2.0 * M * N * K / us / 1e6
            _lib.check(_lib.lib.flute_b200_check(0))
            r = dict(N=N, K=K, M=M, bits=bits, us=us, gbs=gbs, hbm_frac=gbs / peaks["hbm_gbs"], tflops=tfl,
                     tc_frac=tfl / peaks["bf16_tflops"], copies=ncopies, grid=force_grid)
            results.append(r)
            print(f"N={N:6d} K={K:6d} M={M:5d} W{bits}: {us:9.2f} us  {gbs:8.1f} GB/s ({100 * r['hbm_frac']:5.1f}% HBM)  "
                  f"{tfl:8.1f} TFLOP/s ({100 * r['tc_frac']:5.1f}% TC)  copies={ncopies}" + (f"  grid={force_grid}" if force_grid else ""), flush=True)
            del Qs, Ss
            torch.cuda.empty_cache()
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        with open(args.json, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
