#!/usr/bin/env python
"""bench.py -- Llama-3-8B W4G64 bf16 decode (M = 1) through the B200 LUT-qGEMM engine.

A "step" is one decode token through every quantised linear of Llama-3-8B as vLLM fuses them
(qkv 6144x4096, o 4096x4096, gate_up 28672x4096, down 4096x14336; 32 layers = 128 GEMMs,
3.71 GB of packed weights + scales, far larger than the 126 MB L2, so every step streams from
HBM).  Attention / norms / activations are not part of the reference's hot path and are not run:
tok/s here is "linears only", the quantity SURVEY.md section 8(d) defines.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-configs]

N > 1 (torchrun): tensor parallel, every linear column-sharded N/tp per rank;
strong scaling (one token stream).  The exchange step is fused into the GEMM: the epilogue stores
its [1, N/tp] slice straight into every peer's gathered activation buffer over NVLink (symmetric
memory), each element as one 8-byte {value, sequence number} word; the next linear's activation
warp reads those words and spins on the sequence number (flute_b200/parallel.py::FusedGather).  `tp` in the JSON line reports the step with that exchange, without any exchange, and with
NCCL all-gathers instead.

Keys beyond the base contract: `roofline` (achieved HBM GB/s of the qGEMM kernel vs the measured
peak), `cpu_baseline` (the reference's dequantize-then-torch.matmul path on the host cores, bounded
sample), `e2e` (the same step through the public Python API inside one CUDA graph, with the
activation H2D copy and the result D2H copy inside the timed region), `clocks`, `gpu_launches`,
and `configs`: the other BASELINE.json configurations (prefill M = 16 / 512 / 4096 per shape, W3G64
fp16 decode, Llama-3.1-70B shapes, HIGGS + Hadamard on Gemma-2-9B shapes), each with its own roofline
fraction -- reported beside the headline, never instead of it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LAYERS = 32
SHAPES = [("qkv", 6144, 4096), ("o", 4096, 4096), ("gate_up", 28672, 4096), ("down", 4096, 14336)]
SHAPES_70B = [("qkv", 10240, 8192), ("o", 8192, 8192), ("gate_up", 57344, 8192), ("down", 8192, 28672)]
LAYERS_70B = 80
# Gemma-2-9B as vLLM fuses it (tests/shapes.py:53-61 of the reference): qkv 4096+2*2048, o, gate_up 2*14336, down
SHAPES_GEMMA = [("qkv", 8192, 3584), ("o", 3584, 4096), ("gate_up", 28672, 3584), ("down", 3584, 14336)]
LAYERS_GEMMA = 42
BITS, GROUP = 4, 64
METRIC = "llama3_8b_w4g64_decode_tok_per_s"


def algorithmic_bytes(M, N, K, bits=BITS, group=GROUP):
    """BASELINE.md section 3 / SURVEY.md section 8(d)."""
    return N * K * bits // 8 + N * (K // group) * 2 + M * K * 2 + M * N * 2 + (2 ** bits) * 2 + (4 ** bits) * 4


def step_bytes(M=1, tp=1, shapes=SHAPES, layers=LAYERS, bits=BITS):
    return layers * sum(algorithmic_bytes(M, N // tp, K, bits) for _, N, K in shapes)


def total_weights():
    return LAYERS * sum(N * K for _, N, K in SHAPES)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm": float(d["hbm_gbs"]), "tc_burst": float(d.get("bf16_tflops", 1705.2)),
                "tc_sustained": float(d.get("bf16_tflops_sustained", 1435.8)), "kind": "measured"}
    # B200_PROFILING.md fallbacks
    return {"hbm": 6650.0, "tc_burst": 1590.0, "tc_sustained": 1590.0, "kind": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# reference arm: the reference's CPU formulation on the host cores
# ----------------------------------------------------------------------------------------------
def run_reference(args):
    import torch
    from oracle import cpu_path
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dtype = torch.bfloat16
    K = 4096
    # size the per-step sample so that (steps + warmup) steps take ~2 minutes at most
    t_cal, w_cal = cpu_path.time_sample(1, 256, K, BITS, GROUP, dtype)
    budget = 120.0 / max(1, args.steps + args.warmup)
    n_cols = int(min(4096, max(256, (budget / max(t_cal, 1e-6)) * 256)) // 128 * 128)
    for _ in range(args.warmup):
        cpu_path.time_sample(1, n_cols, K, BITS, GROUP, dtype)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_path.time_sample(1, n_cols, K, BITS, GROUP, dtype, seed=i)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    # time_sample includes building the inputs; time the compute alone for the reported rate (median of 5: the
    # first call pays thread-pool start-up, which moved this number 7x between two round-1 runs)
    ts = sorted(cpu_path.time_sample(1, n_cols, K, BITS, GROUP, dtype, repeats=1)[0] for _ in range(5))
    t_compute, weights = ts[len(ts) // 2], n_cols * K
    tok_s = (weights / t_compute) / total_weights()
    sample = (f"o_proj columns [0,{n_cols}) x K=4096 (M=1, W4G64 bf16): {weights / 1e6:.2f}M of "
              f"{total_weights() / 1e9:.2f}G weights/token, median of 5, extrapolated")
    line = {
        "impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tok/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "Llama-3-8B linears W4G64 bf16 decode M=1 (reference's dequantize-then-torch.matmul on CPU)",
                   "sample": sample},
        "cpu_baseline": {"value": tok_s, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tok_s, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# own arm
# ----------------------------------------------------------------------------------------------
NF4 = [-1.0, -0.6962, -0.5251, -0.3949, -0.2844, -0.1848, -0.0911, 0.0, 0.0796, 0.1609, 0.2461, 0.3379, 0.4407, 0.5626,
       0.7230, 1.0]


def build_linears(torch, dev, shapes, layers, bits, dtype, tp=1, rank=0, seed=1234):
    """Random-init quantised linears, this rank's column shard.  Any int16 bit pattern is a valid packing at every
    bit width (4-bit: 4 nibble pairs per word; 3-bit: three planes; 2-bit: 8 pairs), so Q is drawn directly in packed
    form."""
    g = torch.Generator(device=dev).manual_seed(seed + rank)
    out = []
    for _ in range(layers):
        lin = {}
        for name, N, K in shapes:
            n_loc = N // tp
            Q = torch.randint(-32768, 32768, (n_loc // 16 * bits, K), generator=g, dtype=torch.int16, device=dev)
            S = (torch.randn((n_loc, K // GROUP), generator=g, device=dev) * (2.0 / K ** 0.5)).to(dtype)
            lin[name] = (Q, S, n_loc, K)
        out.append(lin)
    return out


class Progress:
    """Phase log on stderr + a watchdog for multi-rank runs: a rank that makes no progress for `limit` seconds
    prints what it was doing (rank 0: also a JSON line carrying the error) and exits instead of hanging the job."""

    def __init__(self, rank, world, limit=180.0):
        self.rank, self.world, self.limit = rank, world, limit
        self.t0 = time.time()
        self.last = self.t0
        self.name = "start"
        if world > 1:
            threading.Thread(target=self._watch, daemon=True).start()

    def __call__(self, name):
        self.name, self.last = name, time.time()
        print(f"[bench rank {self.rank} +{self.last - self.t0:6.1f}s] {name}", file=sys.stderr, flush=True)

    def _watch(self):
        while True:
            time.sleep(5.0)
            if time.time() - self.last > self.limit:
                msg = f"rank {self.rank}: no progress for {self.limit:.0f} s in phase '{self.name}'"
                print(f"[bench] WATCHDOG: {msg}", file=sys.stderr, flush=True)
                if self.rank == 0:
                    print(json.dumps({"metric": METRIC, "value": None, "unit": "tok/s", "n_gpus": self.world,
                                      "error": msg}), flush=True)
                os._exit(4)


def time_replays(torch, step, n, warm=3):
    """CUDA-event time of n back-to-back calls of `step` (ms per call)."""
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def capture(torch, fn):
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            out = fn()
    torch.cuda.synchronize()
    return g, out


def run_configs(torch, dev, progress, peaks):
    """The other BASELINE.json configurations on one GPU, each timed like the headline (CUDA graph, weights >> L2,
    CUDA events) and reported with the roofline fraction that bounds it."""
    import flute_b200
    from flute_b200 import _lib, utils
    from flute_b200.templates import default_template_id
    ws = utils.get_workspace_streamk(dev)
    st = lambda: torch.cuda.current_stream().cuda_stream
    flags = _lib.FLAG_PDL | _lib.FLAG_STATIC_WEIGHTS
    out = []

    def chain_cfg(name, shapes, layers, bits, dtype, hadamard=False, note=""):
        """Decode chain (M = 1) over `layers` distinct layers of `shapes`."""
        progress(f"config {name}: build")
        code = _lib.BF16 if dtype == torch.bfloat16 else _lib.F16
        lins = build_linears(torch, dev, shapes, layers, bits, dtype)
        table = torch.randn(2 ** bits, device=dev).to(dtype) if bits != 4 else torch.tensor(NF4, device=dev).to(dtype)
        table2 = utils.make_qmap2_from_qmap(table)
        if hadamard:      # HIGGS vector_size = 2: the pair table IS the grid (tests/higgs.py:7-17 of the reference)
            table2 = torch.randn((256, 2), device=dev).to(dtype).view(torch.float32).reshape(16, 16, 1).contiguous()
        tid = default_template_id(bits)
        Kin = shapes[0][2]
        x0 = torch.randn((1, Kin), device=dev).to(dtype)
        bufs = {nm: torch.empty((1, N), dtype=dtype, device=dev) for nm, N, K in shapes}
        nlaunch = [0]

        def lin_call(x, lin, nm):
            Q, S, n_loc, K = lin[nm]
            if hadamard:
                h = K & -K                                   # largest power of two dividing K
                y = flute_b200.qgemm_hadamard(x, Q, S, table, table2, ws, bits, GROUP, h, tid, 0)
                nlaunch[0] += 2
                return y
            rc = _lib.lib.flute_b200_qgemm(x.data_ptr(), Q.data_ptr(), bufs[nm].data_ptr(), S.data_ptr(), table.data_ptr(),
                                           table2.data_ptr(), ws.data_ptr(), ws.numel(), 1, n_loc, K, bits, GROUP, 32, code,
                                           flags, dev.index, st())
            _lib.check(rc)
            nlaunch[0] += 1
            return bufs[nm]

        names = [nm for nm, _, _ in shapes]
        width = {nm: K for nm, N, K in shapes}

        def token():
            x = x0
            for lin in lins:
                for nm in names:
                    x = lin_call(x[:, :width[nm]], lin, nm)
            return x

        token(); torch.cuda.synchronize()
        nlaunch[0] = 0
        g, _ = capture(torch, token)
        launches = nlaunch[0]
        ms = time_replays(torch, g.replay, 20, warm=3)
        nbytes = step_bytes(1, 1, shapes, layers, bits)
        gbs = nbytes / (ms * 1e-3) / 1e9
        rec = {"name": name, "workload": f"{layers} layers x {[s[1:] for s in shapes]} (N,K), M=1, W{bits}G{GROUP} "
                                         f"{str(dtype).split('.')[-1]}" + (", Hadamard pre-transform" if hadamard else ""),
               "metric": "tok/s (linears only)", "value": 1e3 / ms, "ms_per_step": ms, "launches_per_step": launches,
               "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                            "ceiling_tok_s": peaks["hbm"] * 1e9 / nbytes},
               "kernel": _lib.dispatch_name(1, shapes[0][1], shapes[0][2], bits, GROUP, code)}
        if note:
            rec["note"] = note
        out.append(rec)
        del lins
        torch.cuda.empty_cache()

    def shape_cfg(name, M, bits, dtype, reps):
        """One launch per (shape, M), distinct weight copies >= 512 MB cycled inside the graph."""
        progress(f"config {name}")
        code = _lib.BF16 if dtype == torch.bfloat16 else _lib.F16
        table = torch.tensor(NF4, device=dev).to(dtype)
        table2 = utils.make_qmap2_from_qmap(table)
        rows = []
        for nm, N, K in SHAPES:
            wbytes = N * K * bits // 8 + N * (K // GROUP) * 2
            ncopies = max(2, min(48, (512 * 2 ** 20 + wbytes - 1) // wbytes))
            lins = build_linears(torch, dev, [(nm, N, K)], ncopies, bits, dtype)
            A = (torch.randn((M, K), device=dev) / 100).to(dtype)
            D = torch.empty((M, N), dtype=dtype, device=dev)

            def all_copies():
                for lin in lins:
                    Q, S, n_loc, _ = lin[nm]
                    _lib.check(_lib.lib.flute_b200_qgemm(A.data_ptr(), Q.data_ptr(), D.data_ptr(), S.data_ptr(), table.data_ptr(),
                                                         table2.data_ptr(), ws.data_ptr(), ws.numel(), M, N, K, bits, GROUP, 32,
                                                         code, flags, dev.index, st()))

            all_copies(); torch.cuda.synchronize()
            g, _ = capture(torch, all_copies)
            us = time_replays(torch, g.replay, reps, warm=2) * 1e3 / ncopies
            gbs = algorithmic_bytes(M, N, K, bits) / us / 1e3
            tfl = 2.0 * M * N * K / us / 1e6
            row = {"shape": nm, "N": N, "K": K, "us": us, "GBps": gbs, "hbm_frac": gbs / peaks["hbm"], "TFLOPs": tfl,
                   "tensor_frac_burst": tfl / peaks["tc_burst"], "tensor_frac_sustained": tfl / peaks["tc_sustained"]}
            rows.append(row)
            del lins
            torch.cuda.empty_cache()
        bound = "tensor" if M >= 512 else "hbm"
        key = "tensor_frac_burst" if bound == "tensor" else "hbm_frac"
        out.append({"name": name, "workload": f"Llama-3-8B shapes, M={M}, W{bits}G{GROUP} {str(dtype).split('.')[-1]}, one launch "
                                              f"per shape, >= 512 MB of distinct weights cycled",
                    "bound": bound, "frac_min": min(r[key] for r in rows), "frac_max": max(r[key] for r in rows),
                    "kernel": _lib.dispatch_name(M, 4096, 4096, bits, GROUP, code), "shapes": rows})

    for M, reps in ((1, 10), (16, 10), (512, 5), (4096, 3)):
        shape_cfg(f"llama3_8b_w4g64_bf16_M{M}", M, 4, torch.bfloat16, reps)
    chain_cfg("llama3_8b_w3g64_fp16_decode", SHAPES, LAYERS, 3, torch.float16)
    chain_cfg("llama3.1_70b_w4g64_bf16_decode_tp1", SHAPES_70B, LAYERS_70B, 4, torch.bfloat16,
              note="all 80 layers resident (36 GB of packed weights); per-rank shards at tp = 2/4/8 are the `tp` record of --gpus N runs")
    chain_cfg("gemma2_9b_higgs_w4g64_hadamard_decode", SHAPES_GEMMA, LAYERS_GEMMA, 4, torch.bfloat16, hadamard=True)
    return out


def run_own(args):
    import datetime
    import torch
    import torch.distributed as dist
    import flute_b200
    from flute_b200 import _lib, utils, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    progress = Progress(rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        progress("init_process_group(nccl)")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=150))
        progress("first collective")
        t = torch.ones(1, device=dev)
        dist.all_reduce(t)
        torch.cuda.synchronize()
    tp = world
    peaks = measured_peaks()
    progress("build model")

    layers = build_linears(torch, dev, SHAPES, LAYERS, BITS, torch.bfloat16, tp, rank)
    table = torch.tensor(NF4, dtype=torch.bfloat16, device=dev)
    table2 = utils.make_qmap2_from_qmap(table)
    ws = utils.get_workspace_streamk(dev)
    M = 1
    x0 = torch.randn((M, 4096), device=dev).to(torch.bfloat16)

    # the model's weights are resident and static: allow weight prefetch across kernel boundaries
    flags = (_lib.FLAG_PDL | _lib.FLAG_STATIC_WEIGHTS) if os.environ.get("FLUTE_B200_PDL", "1") != "0" else 0
    from flute_b200 import ops as _ops
    _ops.set_launch_flags(pdl=bool(flags), static_weights=bool(flags))
    launches = [0]
    st = lambda: torch.cuda.current_stream().cuda_stream

    tpx = None
    if tp > 1:
        progress("tensor-parallel exchange: symmetric memory rendezvous")
        tpx = parallel.FusedGather(dev, rank, tp, [(name, M, N, LAYERS) for name, N, K in SHAPES], torch.bfloat16)
    bufs = {name: torch.empty((M, N // tp), dtype=torch.bfloat16, device=dev) for name, N, K in SHAPES}
    gathered = {name: torch.empty((tp, M, N // tp), dtype=torch.bfloat16, device=dev) for name, N, K in SHAPES}

    def linear_cabi(x, lin, name, mode, last=False):
        """mode: 'fused' (epilogue writes every peer's gathered buffer), 'none' (no exchange: local slice only),
        'nccl' (all_gather_into_tensor after the GEMM)."""
        Q, S, n_loc, K = lin[name]
        launches[0] += 1
        if tp > 1 and mode == "fused":
            # the plain image only where something other than the next qgemm reads the buffer: the step's final output
            return tpx.qgemm(x, Q, S, table, table2, ws, name, n_loc, K, BITS, GROUP, flags, plain=last)
        out = bufs[name]
        rc = _lib.lib.flute_b200_qgemm(x.data_ptr(), Q.data_ptr(), out.data_ptr(), S.data_ptr(), table.data_ptr(),
                                       table2.data_ptr(), ws.data_ptr(), ws.numel(), M, n_loc, K, BITS, GROUP, 32,
                                       _lib.BF16, flags if mode != "nccl" else 0, local_rank, st())
        _lib.check(rc)
        if tp > 1 and mode == "nccl":
            return parallel.all_gather_columns(out, out=gathered[name])
        if tp > 1:       # 'none': feed the next linear from a full-width buffer whose other columns are stale
            return gathered[name].view(M, -1)
        return out

    def linear_api(x, lin, name, mode, last=False):
        Q, S, n_loc, K = lin[name]
        out = flute_b200.qgemm_simple(x, Q, S, table, table2, ws, BITS, GROUP)
        if tp > 1:
            return parallel.all_gather_columns(out, out=gathered[name])
        return out

    def token(x, linear, mode="fused"):
        if tp > 1 and mode == "fused":
            tpx.begin_step()
        for li, lin in enumerate(layers):
            qkv = linear(x, lin, "qkv", mode)
            o = linear(qkv[:, :4096], lin, "o", mode)
            gu = linear(o, lin, "gate_up", mode)
            x = linear(gu[:, :14336], lin, "down", mode, last=(li == len(layers) - 1))
        if tp > 1 and mode == "fused":
            tpx.end_step("down")
        return x

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps, warm):
        for _ in range(warm):
            step()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            step()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- device-resident arm: one CUDA graph per step (128 PDL-chained launches) ----
    progress("eager warm-up tokens")
    main_mode = "fused" if tp > 1 else "none"
    for _ in range(2):
        launches[0] = 0
        token(x0, linear_cabi, main_mode)
        launches_per_step = launches[0]
    barrier()
    progress("graph capture")
    graph, _ = capture(torch, lambda: token(x0, linear_cabi, main_mode))
    barrier()
    progress("timed steps")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_per_step = timed(graph.replay, args.steps, max(3, args.warmup))
    clocks = sampler.stop() if rank == 0 else None
    tok_s = 1e3 / ms_per_step
    _lib.check(_lib.lib.flute_b200_check(local_rank))

    tp_record = None
    if tp > 1:
        # the same step without any exchange, and with NCCL all-gathers (eager: NCCL inside graph capture is not
        # relied upon) -- SURVEY.md section 8(e): "report scaling both with and without the collective"
        progress("tp: step without exchange")
        g_none, _ = capture(torch, lambda: token(x0, linear_cabi, "none"))
        ms_none = timed(g_none.replay, args.steps, 3)
        progress("tp: step with NCCL all-gathers (eager)")
        ms_nccl = timed(lambda: token(x0, linear_cabi, "nccl"), max(5, args.steps // 5), 2)
        tp_record = {"exchange": "fused into the GEMM epilogue: NVLink peer stores of {value, sequence} words into every rank's "
                                 "symmetric-memory buffer; the consuming launch's activation warp spins on the sequence "
                                 "numbers (no fence, no collective kernel); a system-scope publish + wait only after the step's final output",
                     "tok_s_fused_exchange": tok_s, "tok_s_without_exchange": 1e3 / ms_none,
                     "tok_s_nccl_allgather_eager": 1e3 / ms_nccl,
                     "bytes_exchanged_per_rank_per_step": LAYERS * sum(M * N // tp * 2 * (tp - 1) for _, N, K in SHAPES)}

    # ---- end-to-end arm: public Python API, pinned host activations in, result out, every step ----
    progress("e2e arm")
    x_host = torch.randn((M, 4096)).to(torch.bfloat16).pin_memory()
    y_host = torch.empty((M, 4096), dtype=torch.bfloat16).pin_memory()
    x_dev = torch.empty((M, 4096), dtype=torch.bfloat16, device=dev)
    e2e_linear = linear_api if tp == 1 else linear_cabi      # tp > 1: the fused-exchange entry point is the public API there
    for _ in range(2):
        x_dev.copy_(x_host, non_blocking=True)
        y_host.copy_(token(x_dev, e2e_linear, main_mode), non_blocking=True)
    barrier()

    def e2e_body():
        x_dev.copy_(x_host, non_blocking=True)
        y2 = token(x_dev, e2e_linear, main_mode)
        y_host.copy_(y2, non_blocking=True)

    graph2, _ = capture(torch, e2e_body)
    for _ in range(max(3, args.warmup)):
        graph2.replay()
        torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        graph2.replay()
        torch.cuda.synchronize()          # the host consumes y_host every step
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_tok_s = args.steps / e2e_s

    if rank == 0:
        nbytes = step_bytes(M, tp)            # per rank == per GPU
        achieved = nbytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if tp == 1 and os.path.exists(tpath):     # ncu dram__bytes_read+write of this bench's launches (one-off capture)
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch_avg")
        line = {
            "metric": METRIC, "value": tok_s, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "Llama-3-8B linears (qkv,o,gate_up,down x32) W4G64 bf16 decode M=1, linears only",
                       "parallelism": f"tp{tp}" if tp > 1 else "single", "packing": "tile_P=32",
                       "l2": "3.7 GB of distinct weights per step >> 126 MB L2 (no flush needed)",
                       "launch": "one CUDA graph per step, 128 qgemm launches" + (", PDL-chained" if flags else "")
                                 + (", all-gather fused into the GEMM epilogue (NVLink peer stores)" if tp > 1 else "")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm"], "peak_kind": peaks["kind"], "traffic": traffic,
                         "algorithmic_bytes_per_launch_avg": nbytes / launches_per_step,
                         "basis": "algorithmic bytes of a step / device time of the step (CUDA events); the step is 128 "
                                  "back-to-back launches of this one kernel, so inter-launch gaps count against it",
                         "kernel": _lib.dispatch_name(M, SHAPES[0][1] // tp, SHAPES[0][2], BITS, GROUP, _lib.BF16)},
            "e2e": {"value": e2e_tok_s, "unit": "tok/s", "h2d_bytes_per_step": x_host.numel() * 2,
                    "d2h_bytes_per_step": y_host.numel() * 2,
                    "api": ("flute_b200.qgemm_simple (torch op, binding=" + _ops.BINDING + ") x128" if tp == 1 else
                            "flute_b200.parallel.FusedGather.qgemm x128") + " in one CUDA graph + pinned H2D/D2H, host sync per step"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if tp_record is not None:
            line["tp"] = tp_record
        if args.gpus == 1 and not args.no_configs:
            try:
                line["configs"] = run_configs(torch, dev, progress, peaks)
            except Exception as exc:      # the headline stands on its own; say what broke instead of dropping the line
                line["configs_error"] = f"{type(exc).__name__}: {exc}"
        if args.gpus == 1 and not args.no_cpu_baseline:
            progress("cpu baseline")
            from oracle import cpu_path
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            ts = sorted(cpu_path.time_sample(1, 4096, 4096, BITS, GROUP, torch.bfloat16, repeats=1)[0] for _ in range(5))
            t_cpu, w_cpu = ts[len(ts) // 2], 4096 * 4096
            line["cpu_baseline"] = {
                "value": (w_cpu / t_cpu) / total_weights(), "unit": "tok/s", "cores": cores, "kind": "port",
                "sample": f"o_proj 4096x4096 M=1 W4G64 bf16, median of 5 ({t_cpu:.2f} s; {w_cpu / 1e6:.1f}M of "
                          f"{total_weights() / 1e9:.2f}G weights/token, extrapolated)"}
        print(json.dumps(line), flush=True)
    progress("done")
    if world > 1:
        # tear the process group down, but never let a stuck teardown turn a finished measurement into a hang
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(20.0)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
