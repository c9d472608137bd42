#!/usr/bin/env python
"""bench.py -- Llama-3-8B W4G64 bf16 decode (M = 1) through the B200 LUT-qGEMM engine.

A "step" is one decode token through every quantised linear of Llama-3-8B as vLLM fuses them
(qkv 6144x4096, o 4096x4096, gate_up 28672x4096, down 4096x14336; 32 layers = 128 GEMMs,
3.71 GB of packed weights + scales, far larger than the 126 MB L2, so every step streams from
HBM).  Attention / norms / activations are not part of the reference's hot path and are not run:
tok/s here is "linears only", the quantity SURVEY.md section 8(d) defines.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun): tensor parallel, every linear column-sharded N/tp per rank + one all-gather
of the [1, N/tp] outputs; strong scaling (one token stream).

Keys beyond the base contract: `roofline` (achieved HBM GB/s of the qGEMM kernel vs the measured
peak), `cpu_baseline` (the reference's dequantize-then-torch.matmul path on the host cores, bounded
sample), `e2e` (the same step through the public Python API inside one CUDA graph, with the
activation H2D copy and the result D2H copy inside the timed region), `clocks`, `gpu_launches`.
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
BITS, GROUP = 4, 64
METRIC = "llama3_8b_w4g64_decode_tok_per_s"


def algorithmic_bytes(M, N, K, bits=BITS, group=GROUP):
    """BASELINE.md section 3 / SURVEY.md section 8(d)."""
    return N * K * bits // 8 + N * (K // group) * 2 + M * K * 2 + M * N * 2 + (2 ** bits) * 2 + (4 ** bits) * 4


def step_bytes(M=1, tp=1):
    return LAYERS * sum(algorithmic_bytes(M, N // tp, K) for _, N, K in SHAPES)


def total_weights():
    return LAYERS * sum(N * K for _, N, K in SHAPES)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


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
    # time_sample includes building the inputs; time the compute alone once for the reported rate
    t_compute, weights = cpu_path.time_sample(1, n_cols, K, BITS, GROUP, dtype, repeats=2)
    tok_s = (weights / t_compute) / total_weights()
    sample = f"o_proj columns [0,{n_cols}) x K=4096 (M=1, W4G64 bf16): {weights / 1e6:.2f}M of {total_weights() / 1e9:.2f}G weights/token, extrapolated"
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
def build_model(torch, dev, tp, rank, seed=1234):
    """Random-init Llama-3-8B quantised linears, this rank's column shard.  Any int16 bit pattern is a
    valid 4-bit packing, so Q is drawn directly in packed form."""
    g = torch.Generator(device=dev).manual_seed(seed + rank)
    table = torch.tensor([-1.0, -0.6962, -0.5251, -0.3949, -0.2844, -0.1848, -0.0911, 0.0, 0.0796, 0.1609,
                          0.2461, 0.3379, 0.4407, 0.5626, 0.7230, 1.0], dtype=torch.bfloat16, device=dev)
    from flute_b200 import utils
    table2 = utils.make_qmap2_from_qmap(table)
    layers = []
    for _ in range(LAYERS):
        lin = {}
        for name, N, K in SHAPES:
            n_loc = N // tp
            Q = torch.randint(-32768, 32768, (n_loc // 16 * BITS, K), generator=g, dtype=torch.int16, device=dev)
            S = (torch.randn((n_loc, K // GROUP), generator=g, device=dev) * (2.0 / K ** 0.5)).to(torch.bfloat16)
            lin[name] = (Q, S, n_loc, K)
        layers.append(lin)
    return layers, table, table2


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
    # Collectives inside CUDA graphs: on for one GPU (none exist), opt-in for TP (FLUTE_B200_TP_GRAPH=1) until the
    # captured-NCCL path has been validated on the target pool; the eager path launches 128 GEMMs + 128 all-gathers
    # per token from the host.
    use_graph = (world == 1) or os.environ.get("FLUTE_B200_TP_GRAPH", "0") == "1"
    progress("build model")

    layers, table, table2 = build_model(torch, dev, tp, rank)
    ws = utils.get_workspace_streamk(dev)
    M = 1
    x0 = torch.randn((M, 4096), device=dev).to(torch.bfloat16)

    # activation buffers (one per linear type; layers chain through them)
    bufs = {name: torch.empty((M, N // tp), dtype=torch.bfloat16, device=dev) for name, N, K in SHAPES}
    gathered = {name: torch.empty((tp, M, N // tp), dtype=torch.bfloat16, device=dev) for name, N, K in SHAPES}
    # the model's weights are resident and static: allow weight prefetch across kernel boundaries
    flags = (_lib.FLAG_PDL | _lib.FLAG_STATIC_WEIGHTS) if (tp == 1 and os.environ.get("FLUTE_B200_PDL", "1") != "0") else 0
    from flute_b200 import ops as _ops
    _ops.set_launch_flags(pdl=bool(flags), static_weights=bool(flags))
    launches = [0]

    def linear_cabi(x, lin, name):
        Q, S, n_loc, K = lin[name]
        out = bufs[name]
        rc = _lib.lib.flute_b200_qgemm(x.data_ptr(), Q.data_ptr(), out.data_ptr(), S.data_ptr(), table.data_ptr(),
                                       table2.data_ptr(), ws.data_ptr(), ws.numel(), M, n_loc, K, BITS, GROUP, 32,
                                       _lib.BF16, flags, local_rank, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc)
        launches[0] += 1
        if tp > 1:
            return parallel.all_gather_columns(out, out=gathered[name])
        return out

    def linear_api(x, lin, name):
        Q, S, n_loc, K = lin[name]
        out = flute_b200.qgemm_simple(x, Q, S, table, table2, ws, BITS, GROUP)
        if tp > 1:
            return parallel.all_gather_columns(out, out=gathered[name])
        return out

    def token(x, linear):
        for lin in layers:
            qkv = linear(x, lin, "qkv")
            o = linear(qkv[:, :4096], lin, "o")
            gu = linear(o, lin, "gate_up")
            x = linear(gu[:, :14336], lin, "down")
        return x

    # ---- device-resident arm: one CUDA graph per step (128 PDL-chained launches) ----
    progress("eager warm-up tokens")
    for _ in range(2):
        launches[0] = 0
        token(x0, linear_cabi)
        launches_per_step = launches[0]
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    if use_graph:
        progress("graph capture")
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            launches[0] = 0
            with torch.cuda.graph(graph, stream=side):
                y = token(x0, linear_cabi)
            launches_per_step = launches[0]
        torch.cuda.synchronize()
        step = graph.replay
    else:
        step = lambda: token(x0, linear_cabi)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    progress("warm-up steps")
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    progress("timed steps")
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    tok_s = 1e3 / ms_per_step
    rc = _lib.lib.flute_b200_check(local_rank)
    _lib.check(rc)

    # ---- end-to-end arm: public Python API, pinned host activations in, result out, every step ----
    x_host = torch.randn((M, 4096)).to(torch.bfloat16).pin_memory()
    y_host = torch.empty((M, 4096), dtype=torch.bfloat16).pin_memory()
    x_dev = torch.empty((M, 4096), dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        x_dev.copy_(x_host, non_blocking=True)
        y_host.copy_(token(x_dev, linear_api), non_blocking=True)
    torch.cuda.synchronize()

    def e2e_eager():
        x_dev.copy_(x_host, non_blocking=True)
        y_host.copy_(token(x_dev, linear_api), non_blocking=True)

    if use_graph:
        progress("e2e graph capture")
        graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph2, stream=side):
                x_dev.copy_(x_host, non_blocking=True)
                y2 = token(x_dev, linear_api)
                y_host.copy_(y2, non_blocking=True)
        torch.cuda.synchronize()
        e2e_step = graph2.replay
    else:
        e2e_step = e2e_eager
    progress("e2e steps")
    for _ in range(max(3, args.warmup)):
        e2e_step()
        torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
        torch.cuda.synchronize()          # the host consumes y_host every step
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_tok_s = args.steps / e2e_s

    if rank == 0:
        peak, peak_kind = measured_peaks()
        nbytes = step_bytes(M, tp)            # per rank == per GPU
        achieved = nbytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch_avg")
        line = {
            "metric": METRIC, "value": tok_s, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "Llama-3-8B linears (qkv,o,gate_up,down x32) W4G64 bf16 decode M=1, linears only",
                       "parallelism": f"tp{tp}" if tp > 1 else "single", "packing": "tile_P=32",
                       "l2": "3.7 GB of distinct weights per step >> 126 MB L2 (no flush needed)",
                       "launch": ("one CUDA graph per step, 128 qgemm launches" if use_graph else
                                  "eager, 128 qgemm launches + 128 all-gathers per step from the host")
                                 + (", PDL-chained" if flags else "")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_kind": peak_kind, "traffic": traffic,
                         "algorithmic_bytes_per_launch_avg": nbytes / launches_per_step,
                         "basis": "algorithmic bytes of a step / device time of the step (CUDA events); the step is 128 "
                                  "back-to-back launches of this one kernel, so inter-launch gaps count against it",
                         "kernel": "fb::dec::qgemm_decode_kernel<4,true,1>"},
            "e2e": {"value": e2e_tok_s, "unit": "tok/s", "h2d_bytes_per_step": x_host.numel() * 2,
                    "d2h_bytes_per_step": y_host.numel() * 2,
                    "api": "flute_b200.qgemm_simple (torch op) x128 " + ("in one CUDA graph" if use_graph else "eager")
                           + " + pinned H2D/D2H, host sync per step"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if args.gpus == 1 and not args.no_cpu_baseline:
            from oracle import cpu_path
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            t_cpu, w_cpu = cpu_path.time_sample(1, 4096, 4096, BITS, GROUP, torch.bfloat16, repeats=3)
            line["cpu_baseline"] = {
                "value": (w_cpu / t_cpu) / total_weights(), "unit": "tok/s", "cores": cores, "kind": "port",
                "sample": f"o_proj 4096x4096 M=1 W4G64 bf16, best of 3 ({t_cpu:.2f} s; {w_cpu / 1e6:.1f}M of "
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
