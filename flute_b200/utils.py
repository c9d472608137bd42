"""Host-side helpers of the qgemm path: LUT^2 builder, workspace, packers, unpack/reconstruct.

Mirrors the public functions of the reference's flute/utils.py (names, arguments, results).
The packers are closed forms of the wire format, vectorised torch that runs
on whatever device W lives on; they are pinned bit-for-bit against the reference's own
`_pack_{2,3,4}bit` through tests/golden/wire_format.npz.  `reconstruct` / `unpack` call the
dequantise kernel directly instead of multiplying by a K x K identity (utils.py:347-407).
"""
from __future__ import annotations

import math
import warnings
from functools import lru_cache
from typing import Dict, List, Optional

import torch

from . import _lib
from .templates import TEMPLATE_CONFIGS, tile_p_of

_WORKSPACES: Dict[torch.device, torch.Tensor] = {}


def make_qmap2_from_qmap(qmap: torch.Tensor) -> torch.Tensor:
    """T [2^b] -> float32 [2^b, 2^b, 1]: entry (i, j) is the T pair (qmap[i], qmap[j]), qmap[i] in
    the low half (flute/utils.py:15-33)."""
    if qmap.ndim != 1:
        raise ValueError
    if qmap.dtype not in [torch.float16, torch.bfloat16]:
        raise TypeError
    n = qmap.shape[0]
    qmap2 = torch.stack([qmap.view(n, 1).expand(n, n), qmap.view(1, n).expand(n, n)], dim=-1).contiguous()
    return qmap2.view(dtype=torch.float32)


@lru_cache(maxsize=8)
def get_device_num_sms(device: torch.device) -> int:
    return torch.cuda.get_device_properties(device).multi_processor_count


def make_workspace_streamk(device: torch.device) -> torch.Tensor:
    """Zeroed scratch, sized by the reference's formula (flute/utils.py:36-45)."""
    nbytes = _lib.lib.flute_b200_workspace_bytes(get_device_num_sms(device))
    return torch.zeros(nbytes, dtype=torch.uint8, device=device)


def get_workspace_streamk(device: torch.device) -> torch.Tensor:
    if device.type != "cuda":
        warnings.warn(f"Only CUDA devices are supported, but got: {device} ({device.type})")
    if device not in _WORKSPACES.keys():
        _WORKSPACES[device] = make_workspace_streamk(device)
    return _WORKSPACES[device]


def safe_cast(tensor: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if tensor.dtype == dtype:
        return tensor
    tensor_casted = tensor.to(dtype=dtype)
    if not (tensor_casted == tensor).all():
        raise ValueError
    return tensor_casted


# ---------------------------------------------------------------------------------------------
# packers
# ---------------------------------------------------------------------------------------------
def _words_to_int16(words: torch.Tensor, K: int) -> torch.Tensor:
    """int64 [P, K/2] holding 32-bit words -> int16 [P, K] (little endian: low half = even k')."""
    lo = (words & 0xFFFF).to(torch.int32)
    hi = ((words >> 16) & 0xFFFF).to(torch.int32)
    both = torch.stack([lo, hi], dim=-1).reshape(words.shape[0], K)
    both = torch.where(both >= 32768, both - 65536, both)
    return both.to(torch.int16).contiguous()


def pack_tile_p(W: torch.Tensor, num_bits: int, tile_P: int = 32) -> torch.Tensor:
    """uint8-valued [K, N] -> int16 [N/16*bits, K] in the FLUTE wire format for `tile_P`."""
    if W.ndim != 2:
        raise NotImplementedError
    if num_bits not in (2, 3, 4):
        raise ValueError
    if num_bits == 3 and tile_P != 32:
        raise NotImplementedError
    K, N = W.shape
    Wq = safe_cast(W, torch.uint8)
    if int(Wq.max()) >= 2 ** num_bits:
        raise OverflowError
    Wl = Wq.to(torch.int64)
    code = (Wl[0::2] << num_bits) | Wl[1::2]                          # [K/2, N] pair codes
    if num_bits == 3:
        if K % 2 or N % 512:
            raise ValueError
        nb = N // 512
        code = code.view(K // 2, nb, 16, 32)                           # [k2, nb, j, t]
        words = []
        for w in range(3):
            acc = torch.zeros((K // 2, nb, 32), dtype=torch.int64, device=W.device)
            for j in range(w, 15, 3):
                acc |= code[:, :, j, :] << (6 * (j // 3))
            acc |= ((code[:, :, 15, :] >> (2 * w)) & 3) << 30
            words.append(acc)
        plane0 = words[0].reshape(K // 2, nb * 32).T
        plane12 = torch.stack([words[1], words[2]], dim=2).reshape(K // 2, nb * 64).T
        return _words_to_int16(torch.cat([plane0, plane12], dim=0), K)
    fields = 32 // (2 * num_bits)
    blk = fields * tile_P
    if K % 2 or N % blk:
        raise ValueError
    code = code.view(K // 2, N // blk, fields, tile_P)
    acc = torch.zeros((K // 2, N // blk, tile_P), dtype=torch.int64, device=W.device)
    for j in range(fields):
        acc |= code[:, :, j, :] << (2 * num_bits * j)
    return _words_to_int16(acc.reshape(K // 2, N // blk * tile_P).T, K)


def get_template_config(num_bits: int, template_id: int, num_sms: int) -> Dict:
    config = TEMPLATE_CONFIGS[(num_bits, template_id)]
    return {
        "tileM": config["TileM"],
        "tileK": config["TileK"],
        "tileP": config["TileP"],
        "blocks": config["SMs_Multiple"] * num_sms,
    }


def get_template_ids(num_bits: int) -> List[int]:
    return [i for b, i in TEMPLATE_CONFIGS.keys() if b == num_bits]


def is_template_supported(M: int, N: int, K: int, num_bits: int, template_id: int, num_sms: int) -> bool:
    """The reference rejects templates with fewer tiles than CTAs (flute/utils.py:322-344).  The new
    kernel sizes its own grid, so every template whose packing it can read is supported."""
    cfg = TEMPLATE_CONFIGS.get((num_bits, template_id))
    if cfg is None:
        return False
    if num_bits == 3 and cfg["TileP"] != 32:
        return False
    block = 512 if num_bits == 3 else (16 // num_bits) * cfg["TileP"]
    return K % 64 == 0 and N % block == 0


def pack(W: torch.Tensor, num_bits: int, template_ids: Optional[List[int]] = None, num_sms: Optional[int] = None,
         group_size: Optional[int] = None) -> torch.Tensor:
    """flute.utils.pack(W, num_bits, template_ids, num_sms) (flute/utils.py:269-299); the legacy
    call `pack(W, num_bits=, group_size=)` of integrations/vllm_utils.py:314-317 packs with tile_P = 32."""
    if W.ndim != 2:
        raise NotImplementedError
    if template_ids is None:
        tile_P = 32
    else:
        tile_Ps = {tile_p_of(num_bits, t) for t in template_ids}
        if len(tile_Ps) != 1:
            raise ValueError
        tile_P = tile_Ps.pop()
    return pack_tile_p(W, num_bits, tile_P)


# ---------------------------------------------------------------------------------------------
# reconstruct / unpack
# ---------------------------------------------------------------------------------------------
def dequantize(weight: torch.Tensor, scales: torch.Tensor, tables2: torch.Tensor, num_bits: int, group_size: int,
               tile_P: int = 32) -> torch.Tensor:
    """Packed weights -> dense W_hat [K, N] in scales.dtype, on the GPU."""
    if weight.dtype != torch.int16 or tables2.dtype != torch.float32:
        raise TypeError
    if scales.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError
    if not (weight.is_cuda and scales.is_cuda and tables2.is_cuda):
        raise ValueError("flute_b200: dequantize needs CUDA tensors (there is no CPU path)")
    weight, scales, tables2 = weight.contiguous(), scales.contiguous(), tables2.contiguous()
    N, K = scales.shape[0], weight.shape[1]
    out = torch.empty((K, N), dtype=scales.dtype, device=scales.device)
    dev = scales.device.index if scales.device.index is not None else torch.cuda.current_device()
    rc = _lib.lib.flute_b200_dequantize(
        weight.data_ptr(), scales.data_ptr(), tables2.data_ptr(), out.data_ptr(), N, K, num_bits, group_size, tile_P,
        _lib.F16 if scales.dtype == torch.float16 else _lib.BF16, dev, torch.cuda.current_stream(scales.device).cuda_stream)
    _lib.check(rc)
    return out


def reconstruct(weight: torch.Tensor, scales: torch.Tensor, tables: torch.Tensor, tables2: torch.Tensor,
                workspace: torch.Tensor, num_bits: int, group_size: int, template_id: int, num_sms: int) -> torch.Tensor:
    """W.T of the dequantised weight, [N, K] (flute/utils.py:347-376), without the identity GEMM."""
    return dequantize(weight, scales, tables2, num_bits, group_size, tile_p_of(num_bits, template_id)).T


def unpack(weight: torch.Tensor, scales: torch.Tensor, workspace: torch.Tensor, num_bits: int, group_size: int,
           template_id_packed: Optional[int] = None, num_sms_packed: Optional[int] = None) -> torch.Tensor:
    """Quantisation indices as a T-valued [N, K] tensor (flute/utils.py:379-407)."""
    ones = torch.ones_like(scales)
    tables = torch.arange(2 ** num_bits, dtype=scales.dtype, device=scales.device)
    tables2 = make_qmap2_from_qmap(tables)
    if template_id_packed is None:   # legacy call without template id: default packing
        from .templates import default_template_id
        template_id_packed = default_template_id(num_bits)
    return reconstruct(weight=weight, scales=ones, tables=tables, tables2=tables2, workspace=workspace,
                       num_bits=num_bits, group_size=group_size, template_id=template_id_packed,
                       num_sms=num_sms_packed if num_sms_packed is not None else 0)
