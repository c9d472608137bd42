"""CPU oracle for the FLUTE LUT-quantized GEMM hot path  --  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, what the reference computes on the path
`flute.qgemm` / `flute.qgemm_hadamard`.  It is the
checker the CUDA path is compared against.  Nothing under `flute_b200/` may
import it; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` do.

Parity status: PINNED for the integer wire format (packers checked bit-for-bit
against the reference's own `flute/utils.py::_pack_{2,3,4}bit` and
`make_qmap2_from_qmap`, imported from /root/reference by
`tests/golden/make_golden.py`; vectors committed under `tests/golden/`), and
pinned-with-tolerance for the accumulated GEMM (the reference's own tests use
a relative Frobenius bound, `tests/kernel.py:12-13`).  The Hadamard
pre-transform is PARITY UNPINNED: the reference holds no test for it
(SURVEY.md section 4), so it is restated from its definition (orthonormal
Sylvester-ordered Walsh-Hadamard, `hadamard_transform_cuda.cu:141-144`).

All citations are into /root/reference.

Conventions
-----------
* `W`      : uint8 [K, N] quantisation indices, values in [0, 2**bits).
* `Q`      : int16 [P, K] packed indices, P = N/16*bits      (utils.py:59-253)
* `S`      : T [N, K/group] group scales                      (qgemm_kernel.hpp:137)
* `table`  : T [2**bits]
* `table2` : float32 [2**bits, 2**bits, 1] = bit view of T pairs (utils.py:15-33)
* T is fp16 (np.float16) or bf16.  numpy has no bf16, so bf16 arrays are carried
  as np.uint16 bit patterns; helpers below convert with round-to-nearest-even.
"""
from __future__ import annotations

import numpy as np

FP16 = "float16"
BF16 = "bfloat16"


# ----------------------------------------------------------------------------
# T <-> float32 helpers (fp16 native, bf16 as uint16 bit patterns)
# ----------------------------------------------------------------------------
def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    bits = np.ascontiguousarray(bits, dtype=np.uint16)
    return (bits.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 (what `__float2bfloat16_rn` and
    `cutlass::NumericConverter<bfloat16_t, float, round_to_nearest>` do,
    conversion_utils.hpp:43-44).  NaNs are quietened."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    rounded = (u + np.uint32(0x7FFF) + ((u >> 16) & np.uint32(1))) >> 16
    nan = np.isnan(x)
    out = rounded.astype(np.uint16)
    if nan.any():
        out = np.where(nan, ((u >> 16) | np.uint32(0x0040)).astype(np.uint16), out)
    return out


def to_f32(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == FP16:
        return np.asarray(x, dtype=np.float16).astype(np.float32)
    if dtype == BF16:
        return bf16_bits_to_f32(x)
    raise TypeError(dtype)


def from_f32(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == FP16:
        return np.asarray(x, dtype=np.float32).astype(np.float16)
    if dtype == BF16:
        return f32_to_bf16_bits(x)
    raise TypeError(dtype)


def storage_dtype(dtype: str):
    return np.float16 if dtype == FP16 else np.uint16


# ----------------------------------------------------------------------------
# Wire format: pack / unpack  (utils.py:59-253, packbits_utils.py:84-140,191-220)
# ----------------------------------------------------------------------------
def block_columns(num_bits: int, tile_p: int) -> int:
    """Columns of N covered by `tile_p` packed rows (config.hpp:218-236)."""
    if num_bits == 3:
        return 16 * tile_p
    return (16 // num_bits) * tile_p


def packed_rows(N: int, num_bits: int) -> int:
    return N // 16 * num_bits


def _check_pack_args(W: np.ndarray, num_bits: int, tile_p: int):
    if W.ndim != 2:
        raise NotImplementedError
    if num_bits not in (2, 3, 4):
        raise ValueError("num_bits")
    if num_bits == 3 and tile_p != 32:
        raise NotImplementedError  # utils.py:138-139
    K, N = W.shape
    if K % 2 != 0 or N % block_columns(num_bits, tile_p) != 0:
        raise ValueError("shape")
    if W.max(initial=0) >= (1 << num_bits):
        raise OverflowError  # packbits_utils.py:31-32


def pack(W: np.ndarray, num_bits: int, tile_p: int = 32) -> np.ndarray:
    """uint8 [K, N] -> int16 [P, K]; closed form of utils._pack_{2,3,4}bit.

    View the result as uint32 [P, K/2] (little endian, low int16 = even k').
    4-bit: byte j of word (p, k2) = (q[2k2, n] << 4) | q[2k2+1, n],
           n = (p // tP) * 4tP + j * tP + p % tP.
    2-bit: nibble j (0..7) = (q[2k2, n] << 2) | q[2k2+1, n],
           n = (p // tP) * 8tP + j * tP + p % tP.
    3-bit: three words per (nb, t, k2); see `_pack3`.
    """
    W = np.ascontiguousarray(W, dtype=np.uint8)
    _check_pack_args(W, num_bits, tile_p)
    K, N = W.shape
    if num_bits == 3:
        return _pack3(W)
    fields = 32 // (2 * num_bits)          # 4 (4-bit) or 8 (2-bit) pair-fields per word
    blk = fields * tile_p
    # pair code for every (k2, n)
    code = (W[0::2].astype(np.uint32) << num_bits) | W[1::2].astype(np.uint32)   # [K/2, N]
    code = code.reshape(K // 2, N // blk, fields, tile_p)                       # n = b*blk + j*tP + t
    shifts = (np.arange(fields, dtype=np.uint32) * (2 * num_bits)).reshape(1, 1, fields, 1)
    words = np.bitwise_or.reduce(code << shifts, axis=2)                         # [K/2, N/blk, tP]
    words = words.reshape(K // 2, N // blk * tile_p).T                           # [P, K/2]
    return np.ascontiguousarray(words).view(np.int16).reshape(-1, K)


def _pack3(W: np.ndarray) -> np.ndarray:
    """3-bit, tile_P = 32 (utils.py:136-253).  For block nb (512 columns), row t,
    k2: 16 six-bit pair codes c_j = (q[2k2, n_j] << 3) | q[2k2+1, n_j],
    n_j = nb*512 + j*32 + t.  c_0..c_14 go five to a word at bits [6*(j//3), +6)
    of word j%3; c_15 is split two bits per word into bits [30, 32).
    Word 0 lives at row nb*32+t, words 1/2 at rows N/16 + nb*64 + {0,32} + t."""
    K, N = W.shape
    code = (W[0::2].astype(np.uint32) << 3) | W[1::2].astype(np.uint32)          # [K/2, N]
    code = code.reshape(K // 2, N // 512, 16, 32)                                # [k2, nb, j, t]
    words = np.zeros((3, K // 2, N // 512, 32), dtype=np.uint32)
    for j in range(15):
        words[j % 3] |= code[:, :, j, :] << np.uint32(6 * (j // 3))
    for w in range(3):
        words[w] |= ((code[:, :, 15, :] >> np.uint32(2 * w)) & np.uint32(3)) << np.uint32(30)
    nb = N // 512
    plane0 = words[0].reshape(K // 2, nb * 32).T                                 # [N/16, K/2]
    plane12 = np.stack([words[1], words[2]], axis=2)                             # [k2, nb, 2, t]
    plane12 = plane12.reshape(K // 2, nb * 64).T                                 # [N/8, K/2]
    Q32 = np.concatenate([plane0, plane12], axis=0)
    return np.ascontiguousarray(Q32).view(np.int16).reshape(-1, K)


def unpack(Q: np.ndarray, num_bits: int, tile_p: int = 32) -> np.ndarray:
    """int16 [P, K] -> uint8 [K, N].  Inverse of `pack`; the integer half of what
    `packbits_utils.hpp:82-142` (2/4-bit) and `:322-363` (3-bit) do in registers."""
    Q = np.ascontiguousarray(Q)
    if Q.dtype != np.int16 or Q.ndim != 2:
        raise TypeError
    P, K = Q.shape
    N = P * 16 // num_bits
    Q32 = Q.view(np.uint32).reshape(P, K // 2)
    mask = np.uint32((1 << num_bits) - 1)
    if num_bits == 3:
        nb = N // 512
        w0 = Q32[: N // 16].reshape(nb, 32, K // 2)
        w12 = Q32[N // 16:].reshape(nb, 2, 32, K // 2)
        ws = [w0, w12[:, 0], w12[:, 1]]
        code = np.empty((nb, 16, 32, K // 2), dtype=np.uint32)
        for j in range(15):
            code[:, j] = (ws[j % 3] >> np.uint32(6 * (j // 3))) & np.uint32(0x3F)
        code[:, 15] = (((ws[0] >> np.uint32(30)) & np.uint32(3))
                       | (((ws[1] >> np.uint32(30)) & np.uint32(3)) << np.uint32(2))
                       | (((ws[2] >> np.uint32(30)) & np.uint32(3)) << np.uint32(4)))
        code = code.reshape(N, K // 2).T                                         # [k2, n]
    else:
        fields = 32 // (2 * num_bits)
        blk = fields * tile_p
        words = Q32.reshape(N // blk, tile_p, K // 2)
        shifts = (np.arange(fields, dtype=np.uint32) * (2 * num_bits)).reshape(1, fields, 1, 1)
        code = (words[:, None, :, :] >> shifts) & np.uint32((1 << (2 * num_bits)) - 1)
        code = code.reshape(N, K // 2).T                                         # [k2, n]
    W = np.empty((K, N), dtype=np.uint8)
    W[0::2] = (code >> np.uint32(num_bits)) & mask
    W[1::2] = code & mask
    return W


def pair_codes(Q: np.ndarray, num_bits: int, tile_p: int = 32) -> np.ndarray:
    """int16 [P, K] -> uint16 [K/2, N]: the index each (k2, n) presents to table2,
    i.e. (q[2k2, n] << bits) | q[2k2+1, n]  (packbits_utils.hpp:99-107)."""
    W = unpack(Q, num_bits, tile_p).astype(np.uint16)
    return (W[0::2] << num_bits) | W[1::2]


# ----------------------------------------------------------------------------
# Lookup tables (utils.py:15-33, integrations/higgs.py:50-71)
# ----------------------------------------------------------------------------
def make_qmap2_from_qmap(table: np.ndarray, dtype: str) -> np.ndarray:
    """T [2^b] -> float32 [2^b, 2^b, 1]; entry (i, j) holds the T pair
    (table[i], table[j]) with table[i] in the LOW half (even k)."""
    t = np.ascontiguousarray(table).view(np.uint16).astype(np.uint32)
    n = t.shape[0]
    pairs = t.reshape(n, 1) | (t.reshape(1, n) << 16)
    return np.ascontiguousarray(pairs).view(np.float32).reshape(n, n, 1)


def table2_halves(table2: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """float32 [2^b, 2^b, 1] -> (low uint16 [4^b], high uint16 [4^b]) bit patterns."""
    u = np.ascontiguousarray(table2).view(np.uint32).reshape(-1)
    return (u & np.uint32(0xFFFF)).astype(np.uint16), (u >> np.uint32(16)).astype(np.uint16)


def higgs_to_flute(codes: np.ndarray, grid: np.ndarray, num_bits: int, dtype: str):
    """HIGGS vector_size=2 -> (W uint8 [K, N], table2) per integrations/higgs.py:50-71.
    `codes` uint8 [K/2, N] (already transposed the way `prepare_data` takes it),
    `grid` T [4^b, 2].  High field of the code is the EVEN k index."""
    mask = (1 << num_bits) - 1
    K2, N = codes.shape
    W = np.empty((K2 * 2, N), dtype=np.uint8)
    W[0::2] = (codes >> num_bits) & mask
    W[1::2] = codes & mask
    g = np.ascontiguousarray(grid).view(np.uint16).astype(np.uint32)            # [4^b, 2]
    pairs = g[:, 0] | (g[:, 1] << 16)
    n = 1 << num_bits
    return W, np.ascontiguousarray(pairs).view(np.float32).reshape(n, n, 1)


# ----------------------------------------------------------------------------
# Dequantisation  (packbits_utils.hpp:105,139,343-361; nf_utils.py:74-89)
# ----------------------------------------------------------------------------
def dequantize(Q: np.ndarray, S: np.ndarray, table2: np.ndarray, num_bits: int,
               group_size: int, dtype: str, tile_p: int = 32) -> np.ndarray:
    """-> W_hat [K, N] in T storage.  W_hat[k, n] = round_T(table2[code].{lo,hi} * S[n, k // group]):
    ONE multiply in T (`__hmul2`); the fp32 product of two T values is exact, so
    rounding it once to T reproduces the hardware result bit for bit."""
    code = pair_codes(Q, num_bits, tile_p)                                       # [K/2, N]
    lo, hi = table2_halves(table2)
    K = Q.shape[1]
    N = code.shape[1]
    if dtype == FP16:
        vlo = lo.view(np.float16).astype(np.float32)[code]
        vhi = hi.view(np.float16).astype(np.float32)[code]
    else:
        vlo = bf16_bits_to_f32(lo)[code]
        vhi = bf16_bits_to_f32(hi)[code]
    vals = np.empty((K, N), dtype=np.float32)
    vals[0::2] = vlo
    vals[1::2] = vhi
    scale = to_f32(S, dtype)                                                     # [N, G]
    scale_kn = np.repeat(scale, group_size, axis=1).T                            # [K, N]
    return from_f32(vals * scale_kn, dtype)


# ----------------------------------------------------------------------------
# GEMM  (tests/kernel.py:68-71 == tune.py:332-335; qgemm_kernel.hpp:705-710)
# ----------------------------------------------------------------------------
def qgemm(A: np.ndarray, Q: np.ndarray, S: np.ndarray, table2: np.ndarray, num_bits: int,
          group_size: int, dtype: str, tile_p: int = 32) -> np.ndarray:
    """D = round_T(A @ W_hat), accumulated wider than T (fp64 here; the reference
    accumulates fp32 inside mma.sync and, across CTAs, in T with atomics -- hence its own
    tests compare with a tolerance, tests/kernel.py:12-13)."""
    W_hat = to_f32(dequantize(Q, S, table2, num_bits, group_size, dtype, tile_p), dtype)
    A32 = to_f32(A, dtype)
    D = A32.astype(np.float64) @ W_hat.astype(np.float64)
    return from_f32(D.astype(np.float32), dtype)


def dense_reference(A: np.ndarray, W: np.ndarray, S: np.ndarray, table: np.ndarray,
                    group_size: int, dtype: str) -> np.ndarray:
    """The reference tests' ground truth, from UNPACKED indices:
    `torch.mm(A, table[W] * repeat_interleave(S, group, 1).T)`  (tests/kernel.py:68-71)."""
    tv = to_f32(table, dtype)[W.astype(np.int64)]                                # [K, N]
    scale_kn = np.repeat(to_f32(S, dtype), group_size, axis=1).T
    W_hat = to_f32(from_f32(tv * scale_kn, dtype), dtype)
    D = to_f32(A, dtype).astype(np.float64) @ W_hat.astype(np.float64)
    return from_f32(D.astype(np.float32), dtype)


def rel_error(D: np.ndarray, D_ref: np.ndarray, dtype: str) -> tuple[float, float]:
    """The two relative Frobenius errors of tests/kernel.py:92-93."""
    d = to_f32(D, dtype).astype(np.float64)
    r = to_f32(D_ref, dtype).astype(np.float64)
    diff = np.linalg.norm(r - d)
    return float(diff / max(np.linalg.norm(d), 1e-30)), float(diff / max(np.linalg.norm(r), 1e-30))


# ----------------------------------------------------------------------------
# Hadamard pre-transform  (qgemm.cpp:201-211, hadamard_transform_cuda.cu:141-144)
# ----------------------------------------------------------------------------
def hadamard(x: np.ndarray, had_size: int, dtype: str) -> np.ndarray:
    """x [..., K] -> reshape(-1, h) @ H_h / sqrt(h), Sylvester order, computed in
    fp64 and rounded once to T.  PARITY UNPINNED (no reference test)."""
    if had_size & (had_size - 1) or had_size <= 0 or had_size > (1 << 15):
        raise ValueError("had_size")
    shape = x.shape
    v = to_f32(x, dtype).astype(np.float64).reshape(-1, had_size)
    h = 1
    while h < had_size:
        v = v.reshape(-1, had_size // (2 * h), 2, h)
        a = v[:, :, 0, :] + v[:, :, 1, :]
        b = v[:, :, 0, :] - v[:, :, 1, :]
        v = np.stack([a, b], axis=2).reshape(-1, had_size)
        h *= 2
    v = v / np.sqrt(float(had_size))
    return from_f32(v.astype(np.float32), dtype).reshape(shape)


def qgemm_hadamard(A, Q, S, table2, num_bits, group_size, had_size, dtype, tile_p=32):
    """qgemm_raw_simple_hadamard (qgemm.cpp:214-244): FHT(A) rounded to T, then qgemm."""
    return qgemm(hadamard(A, had_size, dtype), Q, S, table2, num_bits, group_size, dtype, tile_p)
