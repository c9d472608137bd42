"""Shared helpers for the parity tests: seeded inputs, torch <-> bit-pattern glue, oracle calls."""
from __future__ import annotations

import numpy as np
import torch

FP16_TOL = 2.0e-3    # the reference's own pass/fail bound, tests/kernel.py:12
BF16_TOL = 1.1e-2    # tests/kernel.py:13  (north_star: <= 1e-2 relative; both are asserted)
NORTH_STAR_TOL = 1.0e-2


def tdtype(name: str) -> torch.dtype:
    return torch.float16 if name in ("float16", "f16") else torch.bfloat16


def bits16(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def from_bits16(a: np.ndarray, dtype: torch.dtype) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(dtype)


def make_case(M, N, K, bits, group, dtype, seed=0, table="randn", tile_p=32, identity=False, full_range=True):
    """Synthetic inputs as SURVEY.md section 8(d) specifies them (tests/kernel.py:33-63 on CPU with our seeds)."""
    from flute_b200 import utils
    g = torch.Generator().manual_seed(1234 + seed)
    t = tdtype(dtype) if isinstance(dtype, str) else dtype
    if identity:
        A = torch.eye(K, dtype=t)[:M].contiguous()
    else:
        A = (torch.randn((M, K), generator=g) / 100.).to(t)
    high = 2 ** bits if full_range else 2 ** bits - 1          # the reference draws [0, 2^b - 1)
    W = torch.randint(0, high, (K, N), generator=g, dtype=torch.int64).to(torch.uint8)
    S = torch.randn((N, K // group), generator=g).to(t)
    if table == "arange":
        tab = torch.arange(2 ** bits).to(t)
    elif table == "nf4" and bits == 4:
        tab = torch.tensor([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                            -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                            0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941,
                            0.7229568362236023, 1.0]).to(t)
    else:
        tab = torch.randn(2 ** bits, generator=g).to(t)
    return dict(A=A, W=W, S=S, table=tab, table2=utils.make_qmap2_from_qmap(tab), Q=utils.pack_tile_p(W, bits, tile_p),
                bits=bits, group=group, dtype=t, tile_p=tile_p, M=M, N=N, K=K)


def oracle_qgemm(c) -> torch.Tensor:
    from oracle import c_oracle
    D = c_oracle.qgemm(bits16(c["A"]), c["Q"].numpy(), bits16(c["S"]), c["table2"].numpy(), c["bits"], c["group"],
                       c["dtype"] == torch.bfloat16, c["tile_p"])
    return from_bits16(D, c["dtype"])


def oracle_dequant(c) -> torch.Tensor:
    from oracle import c_oracle
    W = c_oracle.dequantize(c["Q"].numpy(), bits16(c["S"]), c["table2"].numpy(), c["bits"], c["group"],
                            c["dtype"] == torch.bfloat16, c["tile_p"])
    return from_bits16(W, c["dtype"])


def rel_errors(D: torch.Tensor, D_ref: torch.Tensor):
    """Both relative Frobenius errors of tests/kernel.py:92-93."""
    d, r = D.double().cpu(), D_ref.double().cpu()
    diff = (r - d).norm()
    return (diff / d.norm().clamp_min(1e-30)).item(), (diff / r.norm().clamp_min(1e-30)).item()


def tol_for(dtype) -> float:
    return min(FP16_TOL if dtype == torch.float16 else BF16_TOL, NORTH_STAR_TOL)


def assert_same_values(a: torch.Tensor, b: torch.Tensor, what: str = "") -> None:
    """Exact equality the way the reference's identity tests define it (`(D_ == D).all()`, tests/kernel.py:91 and
    tests/higgs.py:103): every element equal as a number.  -0 and +0 compare equal -- an fp32-accumulated
    identity GEMM turns (-0) * 1 + 0 into +0 in the reference kernel as well."""
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert not torch.isnan(a).any() and not torch.isnan(b).any(), what
    bad = (a != b).sum().item()
    assert bad == 0, f"{what}: {bad} of {a.numel()} elements differ"
