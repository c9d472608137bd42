"""flute_b200 -- Blackwell-native (sm_100a) LUT-quantized GEMM engine behind FLUTE's Python API.

Exports the reference's operator surface (flute/__init__.py:12-69):
    qgemm(input, weight, scales, tables, tables2, workspace, num_bits, group_size, template_id, num_sms)
    qgemm_hadamard(..., num_bits, group_size, hadamard_size, template_id, num_sms)
and the legacy names its in-tree vLLM shim still calls (integrations/vllm_utils.py:86,335-344):
    qgemm_simple(input, weight, scales, tables, tables2, workspace, num_bits, group_size)
    NUM_SMS
plus TEMPLATE_CONFIGS, __version__, and the `utils` / `tune` / `ops` submodules.
The hot path is one hand-written tcgen05/TMA kernel in libflute_b200.so, reached through a
C ABI (include/flute_b200.h).  No CPU fallback exists: importing without the built library fails.
"""
from __future__ import annotations

from typing import Callable, cast

import torch

from . import _lib          # noqa: F401  (fails loudly when libflute_b200.so is missing)
from . import ops           # registers torch.ops.flute.qgemm_raw_simple[_hadamard]
from .templates import TEMPLATE_CONFIGS, default_template_id

__version__ = "0.4.2+b200.0.1"

qgemm = cast(Callable[..., torch.Tensor], torch.ops.flute.qgemm_raw_simple)
qgemm_hadamard = cast(Callable[..., torch.Tensor], torch.ops.flute.qgemm_raw_simple_hadamard)


def qgemm_simple(input: torch.Tensor, weight: torch.Tensor, scales: torch.Tensor, tables: torch.Tensor,
                 tables2: torch.Tensor, workspace: torch.Tensor, num_bits: int, group_size: int) -> torch.Tensor:
    """Legacy 8-argument entry point: weights packed with the default layout (tile_P = 32)."""
    return qgemm(input, weight, scales, tables, tables2, workspace, num_bits, group_size,
                 default_template_id(num_bits), 0)


def __getattr__(name: str):
    if name == "NUM_SMS":   # legacy module constant, resolved lazily so the import works without a GPU
        from . import utils
        return utils.get_device_num_sms(torch.device("cuda", torch.cuda.current_device()))
    raise AttributeError(name)


from . import utils         # noqa: E402
from . import tune          # noqa: E402
