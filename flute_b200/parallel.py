"""Tensor-parallel column sharding of a FLUTE-packed linear.

Rank r of `world` owns output columns [r*N/world, (r+1)*N/world).  Because packed row p only
holds columns of block p // tile_P, the shard of the packed weight is a plain ROW SLICE
(two slices for 3-bit, whose planes 1/2 live after plane 0) -- no unpack/repack, unlike the
reference's load path (flute/integrations/vllm_utils.py:265-326).  The forward exchange is one
all-gather of the [M, N/world] outputs (north_star); K is never sharded.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_block(num_bits: int, tile_P: int) -> int:
    return 512 if num_bits == 3 else (16 // num_bits) * tile_P


def shard_packed_linear(weight: torch.Tensor, scales: torch.Tensor, num_bits: int, rank: int, world: int,
                        tile_P: int = 32) -> Tuple[torch.Tensor, torch.Tensor]:
    """(weight [P, K] int16, scales [N, G]) -> this rank's (weight [P/world, K], scales [N/world, G])."""
    N = scales.shape[0]
    if N % world != 0 or (N // world) % shard_block(num_bits, tile_P) != 0:
        raise ValueError(f"N={N} cannot be column-sharded {world} ways at {num_bits} bits / tile_P={tile_P}")
    n_loc = N // world
    s = scales[rank * n_loc:(rank + 1) * n_loc].contiguous()
    if num_bits == 3:
        p0 = N // 16                      # plane-0 rows
        a = weight[rank * (p0 // world):(rank + 1) * (p0 // world)]
        p12 = N // 8
        b = weight[p0 + rank * (p12 // world):p0 + (rank + 1) * (p12 // world)]
        return torch.cat([a, b], dim=0).contiguous(), s
    P = weight.shape[0]
    return weight[rank * (P // world):(rank + 1) * (P // world)].contiguous(), s


def all_gather_columns(local: torch.Tensor, group: Optional[dist.ProcessGroup] = None,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[M, n_loc] per rank -> [M, n_loc * world] (concatenated on the last dim), one all-gather."""
    world = dist.get_world_size(group)
    M, n_loc = local.shape
    if world == 1:
        return local
    if out is None:
        out = torch.empty((world, M, n_loc), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(world * M, n_loc), local.contiguous(), group=group)
    if M == 1:
        return out.view(1, world * n_loc)          # rank-major == column order when there is one row
    return out.permute(1, 0, 2).reshape(M, world * n_loc)
