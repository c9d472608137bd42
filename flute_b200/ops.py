"""Torch operator surface: the reference's schemas, verbatim, backed by the C ABI.

Two bindings register `torch.ops.flute.*`; `BINDING` says which one is live:
  "cpp"     flute_b200/_flute_b200_torch.so -- a compiled TORCH_LIBRARY(flute) / TORCH_LIBRARY_IMPL(flute, CUDA)
            shim over the C ABI (csrc/torch_binding.cpp; the counterpart of flute/csrc/qgemm.cpp:251-260), a few
            microseconds of host time per call.  Used whenever it has been built (build.py --torch,
            __graft_entry__.build()).
  "python"  the torch.library Python implementation below over ctypes (tens of microseconds per eager call; the
            same C-ABI call).  Used when the shim is absent or FLUTE_B200_PY_OPS=1.
Either way the kernels are the ones in libflute_b200.so; neither binding computes anything itself.

Schemas are those of `TORCH_LIBRARY(flute, m)` (flute/csrc/qgemm.cpp:251-254); the fake
implementations carry the input contract of flute/ops.py:4-83 so torch.compile / opcheck
behave as they do with the reference.  The CUDA implementations hand raw pointers, the
current device and the CURRENT STREAM (qgemm.cpp:101-105: CUDA-graph safe) to
libflute_b200.so.  There is no CPU implementation and no fallback.
"""
from __future__ import annotations

import torch

from . import _lib
from .templates import tile_p_of

_SCHEMA_QGEMM = (
    "(Tensor input, Tensor weight, Tensor scales, Tensor table, Tensor table2, Tensor(a!) workspace, "
    "int num_bits, int group_size, int template_id, int num_sms) -> Tensor")
_SCHEMA_QGEMM_HADAMARD = (
    "(Tensor input, Tensor weight, Tensor scales, Tensor table, Tensor table2, Tensor(a!) workspace, "
    "int num_bits, int group_size, int hadamard_size, int template_id, int num_sms) -> Tensor")

import os

# Programmatic dependent launch (FLUTE_B200_PDL, default on): the kernel's prologue (barrier init, TMEM
# allocation) overlaps the tail of the previous kernel in the stream; it waits (griddepcontrol.wait) before
# reading any tensor.  Safe next to arbitrary neighbours: kernels launched without the attribute serialise.
# FLUTE_B200_STATIC_WEIGHTS=1 additionally lets it start streaming packed weights / scales / tables before
# that wait -- valid when they are not produced by work still in flight on the stream (a served model).
LAUNCH_FLAGS = _lib.FLAG_PDL if os.environ.get("FLUTE_B200_PDL", "1") != "0" else 0
if os.environ.get("FLUTE_B200_STATIC_WEIGHTS", "0") == "1":
    LAUNCH_FLAGS |= _lib.FLAG_STATIC_WEIGHTS


def set_launch_flags(pdl: bool = True, static_weights: bool = False) -> None:
    """Process-wide launch behaviour of flute.qgemm (see the comment above)."""
    global LAUNCH_FLAGS
    LAUNCH_FLAGS = (_lib.FLAG_PDL if pdl else 0) | (_lib.FLAG_STATIC_WEIGHTS if (pdl and static_weights) else 0)
    if _cpp is not None:
        _cpp.flute_b200_torch_set_launch_flags(LAUNCH_FLAGS)

NAMESPACE = "flute"
_TORCH_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_flute_b200_torch.so")
BINDING = "python"
_cpp = None
try:
    if os.environ.get("FLUTE_B200_PY_OPS", "0") != "1" and os.path.exists(_TORCH_LIB_PATH):
        import ctypes as _ctypes
        torch.ops.load_library(_TORCH_LIB_PATH)          # runs TORCH_LIBRARY(flute) + the CUDA impl registration
        _cpp = _ctypes.CDLL(_TORCH_LIB_PATH)             # same handle: for flute_b200_torch_set_launch_flags
        _cpp.flute_b200_torch_set_launch_flags.argtypes = [_ctypes.c_int]
        _cpp.flute_b200_torch_set_launch_flags.restype = None
        _cpp.flute_b200_torch_set_launch_flags(LAUNCH_FLAGS)
        BINDING = "cpp"
    else:
        torch.library.define(f"{NAMESPACE}::qgemm_raw_simple", _SCHEMA_QGEMM)
        torch.library.define(f"{NAMESPACE}::qgemm_raw_simple_hadamard", _SCHEMA_QGEMM_HADAMARD)
        torch.library.define(f"{NAMESPACE}::hadamard_transform", "(Tensor input, int hadamard_size) -> Tensor")
except RuntimeError as exc:   # the reference's own extension is loaded in this process
    raise ImportError(
        "torch.ops.flute.qgemm_raw_simple is already registered (is the reference `flute` package "
        "imported?). flute_b200 replaces it and cannot coexist in one process.") from exc


def _dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float16:
        return _lib.F16
    if dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(f"flute_b200: unsupported dtype {dtype} (fp16 / bf16 only)")


def _check_inputs(input, weight, scales, table, table2, workspace, num_bits, group_size) -> None:
    """The formal input contract (flute/ops.py:17-49) plus what the reference silently assumes
    (contiguity, qgemm.cpp:71 reads raw data_ptr)."""
    if not all([
        input.ndim >= 2,
        weight.ndim == 2,
        scales.ndim == 2,
        table.ndim == 1,
        table2.ndim == 3,
        workspace.ndim == 1,
    ]):
        raise ValueError("flute_b200: bad tensor ranks")
    dtype = input.dtype
    if dtype not in (torch.float16, torch.bfloat16):
        raise TypeError("flute_b200: input must be fp16 or bf16")
    if not all([
        weight.dtype == torch.int16,
        scales.dtype == dtype,
        table.dtype == dtype,
        table2.dtype == torch.float32,
        workspace.dtype == torch.uint8,
    ]):
        raise TypeError("flute_b200: dtype mismatch")
    if not all([
        weight.shape[1] == input.shape[-1],                         # K
        weight.shape[1] == scales.shape[1] * group_size,            # K
        weight.shape[0] == int(num_bits * (scales.shape[0] / 16)),  # P
        table.shape[0] == 2 ** num_bits,
        table2.shape[0] == 2 ** num_bits,
        table2.shape[1] == 2 ** num_bits,
        table2.shape[2] == 1,
    ]):
        raise ValueError("flute_b200: shape mismatch")


def _qgemm_cuda(input, weight, scales, table, table2, workspace, num_bits, group_size, template_id, num_sms):
    _check_inputs(input, weight, scales, table, table2, workspace, num_bits, group_size)
    for t in (weight, scales, table, table2, workspace):
        if not t.is_contiguous() or t.device != input.device:
            raise ValueError("flute_b200: weight/scales/tables/workspace must be contiguous and on the input's device")
    K = input.shape[-1]
    N = scales.shape[0]
    x = input.reshape(-1, K)
    if not x.is_contiguous():
        x = x.contiguous()
    M = x.shape[0]
    out = torch.empty((M, N), dtype=input.dtype, device=input.device)
    if M > 0:
        dev = input.device.index if input.device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(input.device).cuda_stream
        rc = _lib.lib.flute_b200_qgemm(
            x.data_ptr(), weight.data_ptr(), out.data_ptr(), scales.data_ptr(), table.data_ptr(), table2.data_ptr(),
            workspace.data_ptr(), workspace.numel(), M, N, K, num_bits, group_size,
            tile_p_of(num_bits, template_id), _dtype_code(input.dtype), LAUNCH_FLAGS, dev, stream)
        _lib.check(rc)
    return out.reshape(input.shape[:-1] + (N,))


def hadamard_transform(x: torch.Tensor, hadamard_size: int) -> torch.Tensor:
    """apply_hadamard (flute/csrc/qgemm.cpp:201-211): x.reshape(-1, h) @ H_h / sqrt(h), out of place."""
    if x.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError("Only fp16 and bf16 supported currently")
    if x.shape[-1] % hadamard_size != 0:
        raise ValueError("flute_b200: last dimension must be a multiple of hadamard_size")
    xc = x if x.is_contiguous() else x.contiguous()
    out = torch.empty_like(xc)
    if xc.numel() > 0:
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = _lib.lib.flute_b200_hadamard(xc.data_ptr(), out.data_ptr(), xc.numel() // hadamard_size, hadamard_size,
                                          _dtype_code(x.dtype), dev, stream)
        _lib.check(rc)
    return out


def _qgemm_hadamard_cuda(input, weight, scales, table, table2, workspace, num_bits, group_size, hadamard_size,
                         template_id, num_sms):
    return _qgemm_cuda(hadamard_transform(input, hadamard_size), weight, scales, table, table2, workspace, num_bits,
                       group_size, template_id, num_sms)


if BINDING == "python":
    torch.library.impl(f"{NAMESPACE}::qgemm_raw_simple", "CUDA")(_qgemm_cuda)
    torch.library.impl(f"{NAMESPACE}::qgemm_raw_simple_hadamard", "CUDA")(_qgemm_hadamard_cuda)
    torch.library.impl(f"{NAMESPACE}::hadamard_transform", "CUDA")(hadamard_transform)


@torch.library.register_fake(f"{NAMESPACE}::qgemm_raw_simple")
def _qgemm_raw_simple_abstract(input, weight, scales, table, table2, workspace, num_bits, group_size, template_id,
                               num_sms):
    _check_inputs(input, weight, scales, table, table2, workspace, num_bits, group_size)
    N = scales.shape[0]
    return torch.empty(input.shape[:-1] + (N,), dtype=input.dtype, device=input.device)


@torch.library.register_fake(f"{NAMESPACE}::hadamard_transform")
def _hadamard_transform_abstract(input, hadamard_size):
    return torch.empty_like(input, memory_format=torch.contiguous_format)


@torch.library.register_fake(f"{NAMESPACE}::qgemm_raw_simple_hadamard")
def _qgemm_raw_simple_hadamard_abstract(input, weight, scales, table, table2, workspace, num_bits, group_size,
                                        hadamard_size, template_id, num_sms):
    return _qgemm_raw_simple_abstract(input, weight, scales, table, table2, workspace, num_bits, group_size,
                                      template_id, num_sms)
