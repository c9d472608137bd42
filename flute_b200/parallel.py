"""Tensor-parallel column sharding of a FLUTE-packed linear.

Rank r of `world` owns output columns [r*N/world, (r+1)*N/world).  Because packed row p only
holds columns of block p // tile_P, the shard of the packed weight is a plain ROW SLICE
(two slices for 3-bit, whose planes 1/2 live after plane 0) -- no unpack/repack, unlike the
reference's load path (flute/integrations/vllm_utils.py:265-326).  The forward exchange is one
all-gather of the [M, N/world] outputs (north_star); K is never sharded.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence, Tuple

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


class FusedGather:
    """The forward exchange of a column-sharded linear, fused into the GEMM.

    One symmetric-memory allocation per rank (torch.distributed._symmetric_memory: CUDA VMM buffers mapped into every
    peer over NVLink) holds, for every named output, the gathered activation buffer [M, N_total] in T (the "plain image"),
    its "word image" ([M, N_total] 8-byte words {value, sequence number}) and an arrival counter.  `qgemm` calls
    `flute_b200_qgemm_tp`: the kernel's epilogue stores this rank's [M, N_total / tp] slice into every rank's word image;
    when the activations passed in live in one of these buffers, the kernel reads them from the word image and spins per
    word on the sequence number.  No collective kernel, no fence between producer and consumer, nothing to synchronise on
    the host: a whole token step captures into one CUDA graph.

        fg = FusedGather(dev, rank, tp, [("qkv", 1, 6144, 32), ...], torch.bfloat16)   # name, M, N_total, uses per step
        fg.begin_step();  y = fg.qgemm(x, Q_r, S_r, table, table2, ws, "qkv", n_loc, K, 4, 64, flags);  ...
        out = fg.qgemm(..., "down", ..., plain=True);  fg.end_step("down")           # `out` is data after end_step

    Every rank must issue the same sequence of calls per step; each named buffer is written `uses` times per step, and no
    call may read and write the same named buffer.  A call whose activations live in a gathered buffer does not wait for
    the producing launch to finish: it consumes the words as they land.
    The tensor `qgemm` returns is a handle for the next `qgemm`; it holds current data only for calls made with
    `plain=True` and only after `end_step(name)` (publish + wait: the system-scope side of the exchange).
    """

    def __init__(self, device: torch.device, rank: int, tp: int, outputs: Sequence[Tuple[str, int, int, int]],
                 dtype: torch.dtype, group: Optional[dist.ProcessGroup] = None) -> None:
        from . import _lib
        self._lib = _lib
        self.device, self.rank, self.tp, self.dtype = device, rank, tp, dtype
        group = group if group is not None else dist.group.WORLD
        offs, ll_offs, off = {}, {}, 128 * len(outputs)      # [counters, 128 B apart | per output: plain image, word image]
        for name, M, n_total, uses in outputs:
            offs[name] = off
            off += (M * n_total * 2 + 255) // 256 * 256
            ll_offs[name] = off
            off += (M * n_total * 8 + 255) // 256 * 256
        self.buf, ptrs = self._allocate(off, group)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.out: Dict[str, dict] = {}
        for i, (name, M, n_total, uses) in enumerate(outputs):
            view = self.buf[offs[name]:offs[name] + M * n_total * 2].view(dtype).view(M, n_total)
            flags = (ctypes.c_void_p * 8)(*[p + 128 * i for p in ptrs] + [None] * (8 - tp))
            self.out[name] = dict(M=M, n_total=n_total, uses=uses, view=view, base=ptrs[rank] + offs[name], nbytes=M * n_total * 2,
                                  out_peers=[p + offs[name] for p in ptrs], ll_base=ptrs[rank] + ll_offs[name],
                                  ll_peers=[p + ll_offs[name] for p in ptrs], flags=flags, my_flag=ptrs[rank] + 128 * i,
                                  calls=0, published=0)

    def _allocate(self, nbytes: int, group) -> Tuple[torch.Tensor, list]:
        """This rank's zeroed symmetric buffer and every rank's mapping of it (peer pointers, rank order)."""
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        buf.zero_()
        self.hdl = symm_mem.rendezvous(buf, group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        torch.cuda.synchronize(self.device)
        dist.barrier(group)                      # every rank's images are zero before anyone's first store can land
        return buf, ptrs

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def begin_step(self) -> None:
        for o in self.out.values():
            o["calls"] = 0
            o["published"] = 0
        st = self._stream()
        self._lib.check(self._lib.lib.flute_b200_tp_advance(self.epoch.data_ptr(), self.device.index, st))

    def _source_of(self, x: torch.Tensor):
        """The gathered buffer `x` lives in (None: a local tensor) and x's element offset inside it."""
        p = x.data_ptr()
        for o in self.out.values():
            if o["base"] <= p < o["base"] + o["nbytes"]:
                return o, (p - o["base"]) // 2
        return None, 0

    def qgemm(self, x: torch.Tensor, Q: torch.Tensor, S: torch.Tensor, table: torch.Tensor, table2: torch.Tensor,
              workspace: torch.Tensor, name: str, n_loc: int, K: int, num_bits: int, group_size: int, flags: int,
              tile_P: int = 32, plain: bool = False) -> torch.Tensor:
        """This rank's column slice of linear `name`, written into every rank's gathered buffer.  `plain=True` also keeps
        the plain image current (for `end_step` and readers outside this class)."""
        _lib = self._lib
        o = self.out[name]
        M = x.shape[0]
        if M != o["M"] or n_loc * self.tp != o["n_total"] or x.stride(-1) != 1:
            raise ValueError("flute_b200: FusedGather.qgemm shape mismatch")
        if o["calls"] >= o["uses"]:
            raise ValueError(f"flute_b200: `{name}` written more than its declared {o['uses']} times per step")
        d = _lib.TpDesc()
        d.tp, d.rank, d.n_total = self.tp, self.rank, o["n_total"]
        for r in range(self.tp):
            d.out_peers[r] = o["out_peers"][r]
            d.ll_peers[r] = o["ll_peers"][r]
        d.write_plain = 1 if plain else 0
        d.out_uses, d.out_call = o["uses"], o["calls"]
        src, elem_off = self._source_of(x)
        if src is o:
            # A launch that reads its activations from a word image does not wait for the producing grid (the words carry
            # their own readiness); what keeps a later writer of an image behind its readers is that a launch needs ALL of
            # its input before it can write anything -- which says nothing about a launch overwriting its own input.
            raise ValueError(f"flute_b200: `{name}` cannot be both the activations and the output of one call")
        if src is None and not x.is_contiguous():
            raise ValueError("flute_b200: local activations must be contiguous")
        if src is not None:
            if src["calls"] < 1:
                raise ValueError("flute_b200: FusedGather.qgemm reads a gathered buffer that was not written in this step")
            if M > 1 and x.stride(0) != src["n_total"]:
                raise ValueError("flute_b200: activations must be rows of the gathered buffer")
            d.in_ll = src["ll_base"] + 8 * elem_off          # the word image of the same elements
            d.in_ll_stride = src["n_total"]
            d.in_uses, d.in_call = src["uses"], src["calls"] - 1
        d.epoch = self.epoch.data_ptr()
        code = _lib.BF16 if self.dtype == torch.bfloat16 else _lib.F16
        st = self._stream()
        rc = _lib.lib.flute_b200_qgemm_tp(x.data_ptr(), Q.data_ptr(), S.data_ptr(), table.data_ptr(), table2.data_ptr(),
                                          workspace.data_ptr(), workspace.numel(), M, n_loc, K, num_bits, group_size, tile_P,
                                          code, flags, self.device.index, st, ctypes.byref(d))
        _lib.check(rc)
        o["calls"] += 1
        return o["view"]

    def end_step(self, name: str) -> None:
        """Make the latest `plain=True` write of `name` readable by anything on this stream: publish this rank's stores
        (system-scope fence + one arrival on every rank's counter), then wait for all `tp` arrivals of the step."""
        o = self.out[name]
        if o["published"]:
            raise ValueError(f"flute_b200: `{name}` already published in this step")
        st = self._stream()
        lib = self._lib.lib
        self._lib.check(lib.flute_b200_tp_publish(o["flags"], self.tp, self.device.index, st))
        self._lib.check(lib.flute_b200_tp_wait(o["my_flag"], self.tp, self.tp, self.epoch.data_ptr(), self.device.index, st))
        o["published"] = 1
