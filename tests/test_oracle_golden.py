"""Pins the oracle (numpy and C) to vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py imports flute/utils.py::_pack_{2,3,4}bit, make_qmap2_from_qmap and
evaluates tests/kernel.py:68-71 / tests/higgs.py:7-17 with torch on CPU)."""
import numpy as np
import pytest

from conftest import golden_cases
from oracle import c_oracle, flute_oracle as O


def _parse_pack(tokens):
    return int(tokens[1][1:]), int(tokens[2][2:])          # bits, tile_P


def test_numpy_packers_match_reference(golden):
    n = 0
    for base, tok in list(golden_cases(golden, "pack_")) + list(golden_cases(golden, "packramp_")):
        bits, tp = _parse_pack(tok)
        W, Q = golden[base + "_W"], golden[base + "_Q"]
        assert (O.pack(W, bits, tp) == Q).all(), base
        assert (O.unpack(Q, bits, tp) == W).all(), base
        n += 1
    assert n >= 13


def test_c_packers_match_reference(golden):
    for base, tok in list(golden_cases(golden, "pack_")) + list(golden_cases(golden, "packramp_")):
        bits, tp = _parse_pack(tok)
        W, Q = golden[base + "_W"], golden[base + "_Q"]
        assert (c_oracle.pack(W, bits, tp) == Q).all(), base
        assert (c_oracle.unpack(Q, bits, tp) == W).all(), base


def test_qmap2_matches_reference(golden):
    for bits in (2, 3, 4):
        for name, dt in (("f16", O.FP16), ("bf16", O.BF16)):
            table = golden[f"qmap2_b{bits}_{name}_table"]
            ref = golden[f"qmap2_b{bits}_{name}_table2"]
            got = O.make_qmap2_from_qmap(table, dt)
            assert got.shape == ref.shape == (2 ** bits, 2 ** bits, 1)
            assert (got.view(np.uint32) == ref.view(np.uint32)).all()
            # low half = first index = even k
            lo, hi = O.table2_halves(got)
            assert (lo.reshape(2 ** bits, 2 ** bits)[3] == table[3]).all() and (hi.reshape(2 ** bits, 2 ** bits)[:, 1] == table[1]).all()


def _gt(golden, base, tok):
    bits, group = int(tok[1][1:]), int(tok[2][1:])
    dt = O.FP16 if tok[3] == "f16" else O.BF16
    view = (lambda a: a.view(np.float16)) if dt == O.FP16 else (lambda a: a)
    return bits, group, dt, view


def test_dequant_bit_exact_vs_reference_formula(golden):
    """W_hat = table[W] * repeat_interleave(S).T evaluated by torch (tests/kernel.py:68-70) -- bit exact."""
    n = 0
    for base, tok in golden_cases(golden, "gt"):
        bits, group, dt, view = _gt(golden, base, tok)
        What = O.dequantize(golden[base + "_Q"], view(golden[base + "_S"]), golden[base + "_table2"], bits, group, dt)
        assert (np.ascontiguousarray(What).view(np.uint16) == golden[base + "_What"]).all(), base
        Wc = c_oracle.dequantize(golden[base + "_Q"], golden[base + "_S"], golden[base + "_table2"], bits, group, dt == O.BF16)
        assert (Wc == golden[base + "_What"]).all(), base
        n += 1
    assert n >= 8


def test_gemm_within_reference_tolerance(golden):
    """D = torch.mm(A, W_hat) in T on CPU (tests/kernel.py:71): 2.0e-3 / 1.1e-2 both ways."""
    for base, tok in golden_cases(golden, "gt"):
        bits, group, dt, view = _gt(golden, base, tok)
        A, S = view(golden[base + "_A"]), view(golden[base + "_S"])
        D = O.qgemm(A, golden[base + "_Q"], S, golden[base + "_table2"], bits, group, dt)
        tol = 2.0e-3 if dt == O.FP16 else 1.1e-2
        e1, e2 = O.rel_error(D, view(golden[base + "_D"]), dt)
        assert e1 < tol and e2 < tol, (base, e1, e2)
        Dc = c_oracle.qgemm(golden[base + "_A"], golden[base + "_Q"], golden[base + "_S"], golden[base + "_table2"],
                            bits, group, dt == O.BF16)
        e1, e2 = O.rel_error(view(Dc), view(golden[base + "_D"]), dt)
        assert e1 < tol and e2 < tol, (base, e1, e2)
        # numpy and C oracles agree to the last bit except where fp64 summation order matters
        assert (np.ascontiguousarray(D).view(np.uint16) == Dc).mean() > 0.99


def test_higgs_vector_dequant_bit_exact(golden):
    """tests/higgs.py:7-17,60-104: grid[codes] * scales == dequant of the FLUTE form, table2 = pair grid."""
    n = 0
    for base, tok in golden_cases(golden, "higgs"):
        bits = int(tok[1][1:])
        dt = O.FP16 if tok[2] == "f16" else O.BF16
        view = (lambda a: a.view(np.float16)) if dt == O.FP16 else (lambda a: a)
        codes, grid, scales = golden[base + "_codes"], golden[base + "_grid"], view(golden[base + "_scales"])
        W, t2 = O.higgs_to_flute(np.ascontiguousarray(codes.T), grid, bits, dt)
        What = O.dequantize(O.pack(W, bits, 32), scales, t2, bits, 64, dt)
        assert (np.ascontiguousarray(What).view(np.uint16).T == golden[base + "_dense"]).all(), base
        n += 1
    assert n == 3


@pytest.mark.parametrize("bits,tile_p", [(4, 32), (4, 64), (2, 32), (2, 64), (3, 32)])
def test_pack_unpack_roundtrip_property(bits, tile_p):
    rng = np.random.default_rng(bits * 100 + tile_p)
    blk = O.block_columns(bits, tile_p)
    for K, nblk in [(2, 1), (64, 3), (192, 2)]:
        W = rng.integers(0, 1 << bits, size=(K, blk * nblk), dtype=np.uint8)
        Q = O.pack(W, bits, tile_p)
        assert Q.shape == (O.packed_rows(W.shape[1], bits), K) and Q.dtype == np.int16
        assert (O.unpack(Q, bits, tile_p) == W).all()
        assert (c_oracle.pack(W, bits, tile_p) == Q).all()


def test_pack_rejects_bad_input():
    with pytest.raises(OverflowError):
        O.pack(np.full((2, 128), 16, dtype=np.uint8), 4)
    with pytest.raises(ValueError):
        O.pack(np.zeros((2, 100), dtype=np.uint8), 4)
    with pytest.raises(NotImplementedError):
        O.pack(np.zeros((2, 1024), dtype=np.uint8), 3, 64)      # utils.py:138-139


@pytest.mark.parametrize("h", [1, 2, 8, 256, 4096])
def test_hadamard_oracle_properties(h):
    """Parity unpinned upstream (no reference test): orthonormal, involutive, matches the explicit Sylvester matrix."""
    rng = np.random.default_rng(h)
    x = rng.standard_normal((3, 4 * h)).astype(np.float16)
    y = O.hadamard(x, h, O.FP16)
    yc = c_oracle.hadamard(x, h, False).view(np.float16)
    assert np.abs(y.astype(np.float32) - yc.astype(np.float32)).max() <= 2e-3 * max(1.0, np.abs(y).max())
    if h <= 256:
        H = np.array([[1.0]])
        while H.shape[0] < h:
            H = np.block([[H, H], [H, -H]])
        ref = (x.astype(np.float64).reshape(-1, h) @ H / np.sqrt(h)).reshape(x.shape)
        assert np.abs(y.astype(np.float64) - ref).max() < 5e-3 * max(1.0, np.abs(ref).max())
    back = O.hadamard(y, h, O.FP16).astype(np.float32)
    assert np.abs(back - x.astype(np.float32)).max() < 2e-2 * max(1.0, np.abs(x).max())
    with pytest.raises(ValueError):
        O.hadamard(x, 3, O.FP16)
