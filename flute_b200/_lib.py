"""ctypes binding of libflute_b200.so -- the only way Python reaches the kernels.

There is NO fallback: if the library is missing the import fails loudly, and every entry
point raises RuntimeError on a non-zero return code (the reference raises through
AT_ERROR / C10_CUDA_KERNEL_LAUNCH_CHECK, flute/csrc/qgemm.cpp:82,153,171).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflute_b200_prof.so" if os.environ.get("FLUTE_B200_PROFILE") == "1"
                        else "libflute_b200.so")   # _prof: per-role cycle counters, tools/microbench.py only
if os.environ.get("FLUTE_B200_LIB"):                # tools only: A/B an older build of the same C ABI on one box
    LIB_PATH = os.path.join(_HERE, os.environ["FLUTE_B200_LIB"])

# every symbol include/flute_b200.h declares
EXPORTS = (
    "flute_b200_qgemm",
    "flute_b200_qgemm_host",
    "flute_b200_hadamard",
    "flute_b200_dequantize",
    "flute_b200_workspace_bytes",
    "flute_b200_num_sms",
    "flute_b200_max_batch_tile",
    "flute_b200_last_error",
    "flute_b200_error_string",
    "flute_b200_version",
    "flute_b200_set_timeout_ms",
    "flute_b200_check",
    "flute_b200_qgemm_debug",
    "flute_b200_set_trace_buffer",
    "flute_b200_set_variant",
    "flute_b200_dispatch_name",
    "flute_b200_decode_grid",
    "flute_b200_qgemm_tp",
    "flute_b200_tp_publish",
    "flute_b200_tp_advance",
    "flute_b200_tp_wait",
)

F16, BF16 = 0, 1
FLAG_PDL = 1
FLAG_STATIC_WEIGHTS = 2

_vp, _i, _sz, _l = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_long
_u = ctypes.c_uint


class TpDesc(ctypes.Structure):
    """`flute_b200_tp` of include/flute_b200.h (tensor-parallel fused exchange descriptor)."""
    _fields_ = [("tp", _i), ("rank", _i), ("n_total", _i), ("out_peers", _vp * 8), ("ll_peers", _vp * 8), ("write_plain", _i),
                ("out_uses", _u), ("out_call", _u), ("in_ll", _vp), ("in_ll_stride", _i), ("in_uses", _u), ("in_call", _u),
                ("epoch", _vp)]



def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python flute_b200/build.py` "
            "(or __graft_entry__.build()). flute_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.flute_b200_qgemm.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]
    lib.flute_b200_qgemm.restype = _i
    lib.flute_b200_qgemm_host.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                          _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]
    lib.flute_b200_qgemm_host.restype = _i
    lib.flute_b200_qgemm_debug.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp,
                                           _i, _i, _i, _i, _vp]
    lib.flute_b200_qgemm_debug.restype = _i
    lib.flute_b200_hadamard.argtypes = [_vp, _vp, _l, _i, _i, _i, _vp]
    lib.flute_b200_hadamard.restype = _i
    lib.flute_b200_dequantize.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]
    lib.flute_b200_dequantize.restype = _i
    lib.flute_b200_workspace_bytes.argtypes = [_i]
    lib.flute_b200_workspace_bytes.restype = _sz
    lib.flute_b200_num_sms.argtypes = [_i]
    lib.flute_b200_num_sms.restype = _i
    lib.flute_b200_max_batch_tile.argtypes = [_i]
    lib.flute_b200_max_batch_tile.restype = _i
    lib.flute_b200_last_error.argtypes = []
    lib.flute_b200_last_error.restype = ctypes.c_char_p
    lib.flute_b200_error_string.argtypes = [_i]
    lib.flute_b200_error_string.restype = ctypes.c_char_p
    lib.flute_b200_version.argtypes = []
    lib.flute_b200_version.restype = _i
    lib.flute_b200_set_timeout_ms.argtypes = [_l]
    lib.flute_b200_set_timeout_ms.restype = None
    lib.flute_b200_set_variant.argtypes = [_i]
    lib.flute_b200_set_variant.restype = None
    lib.flute_b200_set_trace_buffer.argtypes = [_vp]
    lib.flute_b200_set_trace_buffer.restype = None
    lib.flute_b200_qgemm_tp.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp,
                                        ctypes.POINTER(TpDesc)]
    lib.flute_b200_qgemm_tp.restype = _i
    lib.flute_b200_tp_publish.argtypes = [ctypes.POINTER(_vp), _i, _i, _vp]
    lib.flute_b200_tp_publish.restype = _i
    lib.flute_b200_tp_advance.argtypes = [_vp, _i, _vp]
    lib.flute_b200_tp_advance.restype = _i
    lib.flute_b200_tp_wait.argtypes = [_vp, _u, _u, _vp, _i, _vp]
    lib.flute_b200_tp_wait.restype = _i
    lib.flute_b200_dispatch_name.argtypes = [_i, _i, _i]
    lib.flute_b200_dispatch_name.restype = ctypes.c_char_p
    lib.flute_b200_decode_grid.argtypes = [ctypes.c_longlong, _i, _i, _i]
    lib.flute_b200_decode_grid.restype = _i
    lib.flute_b200_check.argtypes = [_i]
    lib.flute_b200_check.restype = _i
    return lib


lib = _load()
if os.environ.get("FLUTE_B200_VARIANT"):          # tools / A-B runs only: pin a kernel variant for the whole process
    lib.flute_b200_set_variant(int(os.environ["FLUTE_B200_VARIANT"], 0))


def dispatch_name(M: int, N: int, K: int, num_bits: int, group_size: int, dtype_code: int) -> str:
    """The kernel flute_b200_qgemm dispatches to for this problem (automatic selection)."""
    return lib.flute_b200_dispatch_name(M, num_bits, dtype_code).decode()


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.flute_b200_last_error().decode() or lib.flute_b200_error_string(rc).decode()
        raise RuntimeError(f"flute_b200: {msg} (code {rc})")
