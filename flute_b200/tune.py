"""`flute.tune` surface for the B200 engine.

The reference auto-tunes over 216 CUTLASS templates with triton's do_bench and packs for the
winner's TileP (flute/tune.py:126-257,395-463).  The new kernel has no template zoo: it reads
either packing and sizes its own grid, so "tuning" collapses to choosing the canonical
tile_P = 32 packing and running the reference's own correctness `check`
(flute/tune.py:294-392) against the kernel.
"""
from __future__ import annotations

import warnings
from typing import Dict, NamedTuple, Optional, Tuple

import click
import torch

from . import utils
from .templates import default_template_id

FP16_ERROR_THRESHOLD = 2.0e-3
BF16_ERROR_THRESHOLD = 1.1e-2


class TuneMetaData(NamedTuple):
    M: int
    N: int
    K: int
    num_bits: int
    group_size: int
    num_sms: int
    dtype: torch.dtype
    device: torch.device
    template_id: int

    def to_dict(self) -> Dict:
        data = self._asdict()
        data["dtype"] = str(data["dtype"])
        data["device"] = str(data["device"])
        return data

    @classmethod
    def from_dict(cls, data: Dict) -> "TuneMetaData":
        data = dict(data)
        names = {"torch.float32": torch.float32, "torch.float16": torch.float16, "torch.bfloat16": torch.bfloat16}
        if data.get("dtype") not in names:
            raise ValueError(f"Invalid dtype {data.get('dtype')}")
        data["dtype"] = names[data["dtype"]]
        data["device"] = torch.device(data["device"])
        return cls(**data)


@torch.no_grad()
def check(weight: torch.Tensor, weight_packed: torch.Tensor, metadata: TuneMetaData, uniform: bool,
          identity: bool, raise_on_failure: bool = False) -> bool:
    """The reference's pack-time self check (flute/tune.py:294-392): identity input must reproduce
    table[W] * S exactly; random input must be within 2.0e-3 (fp16) / 1.1e-2 (bf16) relative."""
    import flute_b200 as flute

    if identity:
        inputs = torch.eye(metadata.K, dtype=metadata.dtype, device=metadata.device)
    else:
        inputs = torch.randn((metadata.M, metadata.K), dtype=metadata.dtype, device=metadata.device) / 100.
    scales = torch.randn((metadata.N, metadata.K // metadata.group_size), dtype=metadata.dtype, device=metadata.device)
    if uniform:
        tables = torch.arange(2 ** metadata.num_bits, dtype=metadata.dtype, device=metadata.device)
    else:
        tables = torch.randn(2 ** metadata.num_bits, dtype=metadata.dtype, device=metadata.device)
    tables2 = utils.make_qmap2_from_qmap(tables)
    workspace = utils.get_workspace_streamk(metadata.device)

    weight_ = tables[utils.safe_cast(weight, dtype=torch.int64)]
    scales_ = torch.repeat_interleave(scales, metadata.group_size, dim=1).T
    output_ = torch.mm(inputs, weight_ * scales_)
    output = flute.qgemm(inputs, weight_packed, scales, tables, tables2, workspace, metadata.num_bits,
                         metadata.group_size, metadata.template_id, metadata.num_sms)

    equal = bool((output_ == output).all().item())
    error = ((output_ - output).norm() / output.norm()).item()
    error_ = ((output_ - output).norm() / output_.norm()).item()
    message = (f"WARNING: M={metadata.M} N={metadata.N} K={metadata.K} num_bits={metadata.num_bits} "
               f"group_size={metadata.group_size} dtype={metadata.dtype} uniform={uniform} "
               f"error={error:.3e} error_={error_:.3e}")
    if identity:
        ok = equal
    else:
        threshold = FP16_ERROR_THRESHOLD if metadata.dtype == torch.float16 else BF16_ERROR_THRESHOLD
        ok = (error < threshold) and (error_ < threshold)
    if not ok:
        click.secho(message, fg="red")
        if raise_on_failure:
            raise RuntimeError(message)
    return ok


def tune_and_pack(inputs: torch.Tensor, weight: torch.Tensor, num_bits: int, group_size: int, num_seeds: int = 3,
                  check_correctness: bool = True, check_num_seeds: int = 3) -> Tuple[torch.Tensor, TuneMetaData]:
    """Pack `weight` ([K, N] indices) for the engine and describe it (flute/tune.py:395-463)."""
    if inputs.ndim != 2 or weight.ndim != 2 or inputs.shape[1] != weight.shape[0]:
        raise ValueError
    M = inputs.shape[0]
    K, N = weight.shape
    device = inputs.device
    num_sms = utils.get_device_num_sms(device)
    template_id = default_template_id(num_bits)
    weight_packed = utils.pack(W=weight, num_bits=num_bits, template_ids=[template_id], num_sms=num_sms)
    metadata = TuneMetaData(M=M, N=N, K=K, num_bits=num_bits, group_size=group_size, num_sms=num_sms,
                            dtype=inputs.dtype, device=device, template_id=template_id)
    if check_correctness:
        weight = weight.to(device=device)
        weight_packed = weight_packed.to(device=device)
        for uniform in [True, False]:
            for identity in [True, False]:
                for seed in range(check_num_seeds):
                    torch.manual_seed(seed)
                    check(weight=weight, weight_packed=weight_packed, metadata=metadata, uniform=uniform,
                          identity=identity)
    return weight_packed, metadata


def qgemm_v2(input: torch.Tensor, weight: torch.Tensor, scales: torch.Tensor, table: torch.Tensor,
             table2: torch.Tensor, workspace: torch.Tensor, metadata: TuneMetaData,
             hadamard_size: Optional[int] = None) -> torch.Tensor:
    import flute_b200 as flute

    if hadamard_size is None:
        return flute.qgemm(input, weight, scales, table, table2, workspace, metadata.num_bits, metadata.group_size,
                           metadata.template_id, metadata.num_sms)
    return flute.qgemm_hadamard(input, weight, scales, table, table2, workspace, metadata.num_bits,
                                metadata.group_size, hadamard_size, metadata.template_id, metadata.num_sms)


def maybe_tune_and_repack(weight: torch.Tensor, scales: torch.Tensor, metadata: TuneMetaData,
                          example_batch_size: Optional[int] = None) -> Tuple[torch.Tensor, TuneMetaData]:
    """The reference unpacks and repacks a checkpoint for the local GPU's SM count
    (flute/tune.py:534-591).  The B200 kernel reads the stored packing as is -- the template id only
    tells it the tile_P -- so the weight is returned untouched with refreshed metadata."""
    device = weight.device if weight.device.type == "cuda" else torch.device("cuda")
    if weight.device.type != "cuda":
        warnings.warn(f"[FLUTE]: Moving data from {weight.device} to {device}.")
    num_sms = utils.get_device_num_sms(device)
    new_meta = metadata._replace(M=example_batch_size or metadata.M, num_sms=num_sms)
    return weight, new_meta
