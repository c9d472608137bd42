"""ctypes binding of oracle/libflute_oracle.so (flute_oracle.c).  TEST INFRASTRUCTURE ONLY.

Same functions as oracle/flute_oracle.py, threaded, for sizes numpy is too slow for.
T arrays cross this boundary as uint16 bit patterns (`bits16`).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libflute_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "flute_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libflute_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def bits16(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype == np.float16:
        return a.view(np.uint16)
    if a.dtype in (np.uint16, np.int16):
        return a.view(np.uint16)
    raise TypeError(a.dtype)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def pack(W: np.ndarray, num_bits: int, tile_p: int = 32) -> np.ndarray:
    W = np.ascontiguousarray(W, dtype=np.uint8)
    K, N = W.shape
    Q = np.empty((N // 16 * num_bits, K), dtype=np.int16)
    rc = lib().oracle_pack(_p(W), N, K, num_bits, tile_p, _p(Q))
    if rc:
        raise ValueError(f"oracle_pack rc={rc}")
    return Q


def unpack(Q: np.ndarray, num_bits: int, tile_p: int = 32) -> np.ndarray:
    Q = np.ascontiguousarray(Q, dtype=np.int16)
    P, K = Q.shape
    N = P * 16 // num_bits
    W = np.empty((K, N), dtype=np.uint8)
    rc = lib().oracle_unpack(_p(Q), N, K, num_bits, tile_p, _p(W))
    if rc:
        raise ValueError(f"oracle_unpack rc={rc}")
    return W


def dequantize(Q, S, table2, num_bits, group_size, is_bf16, tile_p=32) -> np.ndarray:
    Q = np.ascontiguousarray(Q, dtype=np.int16)
    S = bits16(S)
    t2 = np.ascontiguousarray(table2).view(np.uint32).reshape(-1)
    P, K = Q.shape
    N = P * 16 // num_bits
    out = np.empty((K, N), dtype=np.uint16)
    rc = lib().oracle_dequantize(_p(Q), _p(S), _p(t2), N, K, num_bits, group_size, tile_p, int(is_bf16), _p(out))
    if rc:
        raise ValueError(f"oracle_dequantize rc={rc}")
    return out


def qgemm(A, Q, S, table2, num_bits, group_size, is_bf16, tile_p=32) -> np.ndarray:
    A = bits16(A)
    Q = np.ascontiguousarray(Q, dtype=np.int16)
    S = bits16(S)
    t2 = np.ascontiguousarray(table2).view(np.uint32).reshape(-1)
    M, K = A.shape
    N = Q.shape[0] * 16 // num_bits
    D = np.empty((M, N), dtype=np.uint16)
    rc = lib().oracle_qgemm(_p(A), _p(Q), _p(S), _p(t2), M, N, K, num_bits, group_size, tile_p, int(is_bf16), _p(D))
    if rc:
        raise ValueError(f"oracle_qgemm rc={rc}")
    return D


def dequant_then_matmul(A, W, S, table, group_size, is_bf16) -> np.ndarray:
    """The reference tests' CPU formulation (tests/kernel.py:68-71) from UNPACKED indices."""
    A = bits16(A)
    W = np.ascontiguousarray(W, dtype=np.uint8)
    S = bits16(S)
    table = bits16(table)
    M, K = A.shape
    N = W.shape[1]
    What = np.empty((K, N), dtype=np.uint16)
    D = np.empty((M, N), dtype=np.uint16)
    rc = lib().oracle_dequant_then_matmul(_p(A), _p(W), _p(S), _p(table), M, N, K, group_size, int(is_bf16), _p(What), _p(D))
    if rc:
        raise ValueError(f"oracle_dequant_then_matmul rc={rc}")
    return D


def hadamard(X, had_size, is_bf16) -> np.ndarray:
    X = bits16(X)
    shape = X.shape
    rows = X.size // had_size
    Y = np.empty(X.size, dtype=np.uint16)
    rc = lib().oracle_hadamard(_p(X), ctypes.c_long(rows), had_size, int(is_bf16), _p(Y))
    if rc:
        raise ValueError(f"oracle_hadamard rc={rc}")
    return Y.reshape(shape)
