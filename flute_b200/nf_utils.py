"""NormalFloat (NF) group quantiser -- the learning-free quantiser behind `prepare_model_flute`.

Same contract as the reference's `flute/nf_utils.py:14-89` (values / pivots of an N(0,1)-quantile code book, absmax
group scales, `searchsorted` against the mid-points), written device-agnostically: the reference hard-codes `.cuda()`
(nf_utils.py:32); here every tensor follows `W.device`, so the quantiser also runs in the CPU test-suite.

    W_dq, W_idx, absmax, values = nf_quantize(W, num_bits, group_size)
      W        [N, K] float                     (an nn.Linear weight: out_features x in_features)
      W_idx    [N, K] int64 code indices        -> pack W_idx.T ([K, N]) with flute_b200.utils.pack
      absmax   [N * K / group_size] group scales -> view as [N, K / group_size] (the kernel's S)
      values   [2^num_bits] code book            (the kernel's `table`; table2 = make_qmap2_from_qmap)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

# the 4-bit code book is pinned to the published NF4 constants (nf_utils.py:28-29 of the reference; QLoRA's table)
NF4_VALUES = (-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
              -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
              0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0)


def get_values_pivots(bits: int = 4, symmetric: bool = False, dtype: torch.dtype = torch.float32,
                      device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Code book `values` (ascending, max |v| = 1) and decision thresholds `pivots` (mid-points)."""
    normal = torch.distributions.normal.Normal(torch.tensor(0.0), torch.tensor(1.0))
    offset = 0.5 * (1 / 32 + 1 / 30)

    def lin(a: float, b: float, n: int) -> torch.Tensor:     # fp32 throughout, in the reference's order of operations
        a_t, b_t = torch.tensor(a), torch.tensor(b)
        return a_t + (torch.arange(n, dtype=torch.float32) / (n - 1)) * (b_t - a_t)

    if symmetric:
        v = normal.icdf(lin(offset, 1 - offset, 2 ** bits))
    else:   # asymmetric: 2^(b-1) negative levels, zero, 2^(b-1) - 1 positive levels
        v1 = -normal.icdf(lin(1 - offset, 0.5, 2 ** (bits - 1)))
        v2 = normal.icdf(lin(0.5, 1 - offset, 2 ** (bits - 1) + 1)[1:])
        v = torch.cat((v1, v2))
    v = v / v.abs().max()
    if bits == 4 and not symmetric:
        v = torch.tensor(NF4_VALUES, dtype=torch.float32)
    p = (v[1:] + v[:-1]) / 2
    return v.to(dtype=dtype, device=device).clone(), p.to(dtype=dtype, device=device).clone()


def _quantize_groups(W: torch.Tensor, group_size: int, values: torch.Tensor, pivots: torch.Tensor,
                     absmax: Optional[torch.Tensor]):
    qx = W.reshape(-1, group_size)
    if absmax is None:
        absmax = qx.abs().amax(dim=1, keepdim=True)
    else:
        absmax = absmax.reshape(-1, 1).to(qx.dtype)
    # (pivots stay fp32 even for fp16 / bf16 weights -- the reference's mixed-dtype searchsorted, nf_utils.py:44,85)
    index = torch.searchsorted(pivots, (qx / absmax).contiguous())
    return index, absmax


def nf_quantize(W: torch.Tensor, num_bits: int, group_size: int,
                custom_scales: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(fake-quantised W, code indices, group scales, code book) -- flute/nf_utils.py:50-71."""
    if W.shape[-1] % group_size != 0:
        raise ValueError(f"in_features {W.shape[-1]} is not a multiple of group_size {group_size}")
    values, pivots = get_values_pivots(num_bits, False, device=W.device)
    index, absmax = _quantize_groups(W, group_size, values, pivots, custom_scales)
    W_dq = (values.to(W.dtype)[index] * absmax).reshape(W.shape)
    return W_dq, index.reshape(W.shape), absmax.squeeze(1), values


def nf_quantize_2(W: torch.Tensor, num_bits: int, group_size: int, dtype: torch.dtype) -> torch.Tensor:
    """Fake quantisation with the KERNEL's arithmetic (code book and scale both in `dtype`, one rounded multiply):
    what `FluteLinear` computes, as a dense weight -- flute/nf_utils.py:74-89, used to test quantised models."""
    values, pivots = get_values_pivots(num_bits, False, device=W.device)
    index, absmax = _quantize_groups(W, group_size, values, pivots, None)
    return (values.to(dtype)[index] * absmax.to(dtype)).reshape(W.shape)
