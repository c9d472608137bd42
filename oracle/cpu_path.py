"""The reference's CPU-runnable formulation of the hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.

"dequantize, then torch.matmul": the ground truth the reference's own tests compute
(tests/kernel.py:68-71 == flute/tune.py:332-335), restated for CPU tensors:

    W_hat = table[W.long()] * repeat_interleave(S, group_size, dim=1).T      # [K, N], rounded to T
    D     = torch.mm(A, W_hat)                                               # T

bench.py times it on the host cores (all threads torch can use) as `cpu_baseline` and as the
`--impl reference` arm.  The reference's CUDA kernel cannot be built in this image (it needs
CUTLASS v3.4.1; see oracle/build_ref.sh), so this port -- not oracle/_ref -- is what gets timed:
`cpu_baseline.kind == "port"`.
"""
from __future__ import annotations

import time
from typing import Dict, Tuple

import torch


def make_inputs(M: int, N: int, K: int, num_bits: int, group_size: int, dtype: torch.dtype, seed: int = 0) -> Dict:
    g = torch.Generator().manual_seed(seed)
    return dict(
        A=(torch.randn((M, K), generator=g) / 100.).to(dtype),
        W=torch.randint(0, 2 ** num_bits, (K, N), generator=g, dtype=torch.int64).to(torch.uint8),
        S=torch.randn((N, K // group_size), generator=g).to(dtype),
        table=torch.randn(2 ** num_bits, generator=g).to(dtype),
    )


def dequant_then_matmul(A: torch.Tensor, W: torch.Tensor, S: torch.Tensor, table: torch.Tensor,
                        group_size: int) -> torch.Tensor:
    W_ = table[W.long()]
    S_ = torch.repeat_interleave(S, group_size, dim=1).T
    return torch.mm(A, W_ * S_)


def time_sample(M: int, N: int, K: int, num_bits: int, group_size: int, dtype: torch.dtype, repeats: int = 1,
                seed: int = 0) -> Tuple[float, int]:
    """Best wall-clock seconds of `repeats` runs on an (N, K) sample; returns (seconds, weights)."""
    x = make_inputs(M, N, K, num_bits, group_size, dtype, seed)
    best = float("inf")
    for _ in range(max(1, repeats)):
        t0 = time.perf_counter()
        dequant_then_matmul(x["A"], x["W"], x["S"], x["table"], group_size)
        best = min(best, time.perf_counter() - t0)
    return best, N * K
