"""Host-side logic of the drop-in surface, on CPU (no kernel launches)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import flute_b200 as flute
from flute_b200 import _lib, templates, utils, tune, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "flute_b200.h")).read()
    declared = set(re.findall(r"FLUTE_B200_API[^;]*?\b(flute_b200_\w+)\s*\(", header))
    assert declared, "no prototypes found in include/flute_b200.h"
    assert declared == set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libflute_b200.so does not export {name}"


def test_library_has_no_libcuda_link_dependency():
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out and "libcudart" not in out


def test_version_and_workspace_formula():
    assert _lib.lib.flute_b200_version() == 100
    # flute/utils.py:36-45: num_sms*4 * 256 * (4*64*8) + 4 * num_sms*4
    for sms in (84, 108, 148):
        assert _lib.lib.flute_b200_workspace_bytes(sms) == sms * 4 * 256 * 2048 + 16 * sms
    assert _lib.lib.flute_b200_workspace_bytes(0) == 0
    assert _lib.lib.flute_b200_max_batch_tile(4) == 64 and _lib.lib.flute_b200_max_batch_tile(5) < 0


def test_dispatch_names():
    """Which kernel the automatic dispatch picks (csrc/qgemm_sm100.cu qgemm_launch): the decode kernel for M <= 16 at 4 bits
    and M <= 4 at 2 bits, the prefill kernel for 4-bit M > 16, the general kernel for the rest (all of 3-bit)."""
    name = lambda M, bits, dt=_lib.BF16: _lib.lib.flute_b200_dispatch_name(M, bits, dt).decode()
    assert name(1, 4) == "fb::dec::qgemm_decode_kernel<4,bf16,MC=1>"
    assert name(4, 4, _lib.F16) == "fb::dec::qgemm_decode_kernel<4,f16,MC=4>"
    assert name(5, 4) == name(16, 4) == "fb::dec::qgemm_decode_kernel<4,bf16,MC=16>"
    assert name(17, 4) == name(4096, 4) == "fb::pre::qgemm_prefill_kernel<bf16>"
    assert name(3, 2) == "fb::dec::qgemm_decode_kernel<2,bf16,MC=4>" and name(5, 2) == "fb::qgemm_sm100_kernel<2,bf16,LARGE>"
    assert name(1, 3, _lib.F16) == name(64, 3, _lib.F16) == "fb::qgemm_sm100_kernel<3,f16,LARGE>"
    assert name(1, 5) == "unsupported"


def test_decode_grid_rule():
    """CTAs per decode-shaped launch (csrc/qgemm_sm100.h decode_grid_for): the measured choices of
    profiles/r02_decode_grid_sweep.log / r02_w3_grid_sweep.log, and the invariants of the rule."""
    grid = _lib.lib.flute_b200_decode_grid
    # (column tiles x k_iters, k_iters, SMs, bits) -> CTAs
    table = [((8 * 64, 64, 148, 4), 128),      # 4096x4096: 4 stages per CTA, 16 contributors per tile
             ((12 * 64, 64, 148, 4), 144),     # 6144x4096: no aligned share near the machine size
             ((56 * 64, 64, 148, 4), 144),     # 28672x4096: bandwidth-bound, every SM but four
             ((8 * 224, 224, 148, 4), 144),    # 4096x14336
             ((6 * 64, 64, 148, 4), 96),       # 3072x4096 (tp 2)
             ((7 * 64, 64, 148, 4), 112),      # 3584x4096 (tp 8)
             ((4 * 224, 224, 148, 4), 128),    # 2048x14336 (tp 2): 7 stages per CTA
             ((2 * 64, 64, 148, 4), 64),       # 1024x4096 (tp 4): two stages per CTA instead of one
             ((1 * 64, 64, 148, 4), 64),       # 512x4096 (tp 8): one stage per CTA
             ((2 * 64, 64, 148, 3), 128),      # 3-bit (2048-column tiles), 4096x4096: the general kernel keeps one stage per CTA
             ((14 * 64, 64, 148, 3), 112),     # 3-bit 28672x4096
             ((2 * 224, 224, 148, 3), 112),    # 3-bit 4096x14336
             ((3 * 64, 64, 148, 3), 96)]       # 3-bit 6144x4096
    for args, want in table:
        assert grid(*args) == want, (args, grid(*args), want)
    assert grid(0, 64, 148, 4) == 0 and grid(10, 0, 148, 4) == 0
    for sms in (8, 16, 84, 132, 148):
        max_grid = sms - 4 if sms > 16 else sms
        for bits in (3, 4):
            for k_iters in (1, 2, 7, 16, 56, 64, 112, 128, 224, 448):
                for tiles in (1, 2, 3, 5, 8, 12, 20, 56, 112):
                    total = tiles * k_iters
                    g = grid(total, k_iters, sms, bits)
                    assert 1 <= g <= min(total, max_grid)
                    if g < min(total, max_grid):      # fewer CTAs than possible only for tile-aligned equal shares
                        share = total // g
                        assert total % g == 0 and k_iters % share == 0 and 2 <= share <= 8
                        assert g * 5 >= min(total, max_grid) * 3 or (total <= max_grid and g >= 64)


def _call_qgemm(M=1, N=512, K=256, bits=4, group=64, tile_p=32, dtype=0, ptr=0x1000):
    p = ctypes.c_void_p(ptr)
    return _lib.lib.flute_b200_qgemm(p, p, p, p, p, p, p, 1 << 20, M, N, K, bits, group, tile_p, dtype, 0, 0, None)


def test_workspace_layout_fits_every_baseline_shape():
    """[64 KB counters | zero-invariant fp32 accumulators | ... | 2 x 128 KB per SM of prefill
    partial-tile slots].  The reference-sized workspace must hold the largest accumulator region any BASELINE.json
    shape needs below the scratch tail (mirrors the checks in csrc/qgemm_*_sm100.cu)."""
    from flute_b200 import _lib
    sms = 148
    total = _lib.lib.flute_b200_workspace_bytes(sms)
    scratch = sms * 2 * 131072 + 256
    counters = 65536
    shapes = [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336), (10240, 8192), (8192, 8192), (57344, 8192),
              (8192, 28672), (28672, 3584)]
    for (N, K) in shapes:
        # decode kernel, 4-bit: one [4 fields][16][128] fp32 block per 512-column tile
        need = counters + ((N + 511) // 512) * 4 * 16 * 128 * 4
        assert need + scratch <= total, (N, K)
        # general kernel, Stream-K at M = 64 (4 fields x 64 rows x 128 lanes fp32 per tile, one activation-row tile)
        need = counters + ((N + 511) // 512) * 4 * 64 * 128 * 4
        assert need + scratch <= total, (N, K)
        assert ((N + 511) // 512) * 2 * ((4096 + 127) // 128) * 4 <= counters        # prefill tile counters at M = 4096


@pytest.mark.parametrize("kwargs,code", [
    (dict(bits=5), -1), (dict(bits=1), -1),                       # AT_ERROR("Unsupported `num_bits`")  qgemm.cpp:171
    (dict(group=32), -2), (dict(group=96), -2),                   # AT_ERROR("Unsupported `group_size`") qgemm.cpp:153
    (dict(dtype=2), -3),                                          # dtype dispatch                       qgemm.cpp:176-193
    (dict(K=100), -4), (dict(K=320, group=128), -4), (dict(N=500), -4), (dict(M=-1), -4),
    (dict(bits=3, N=1024 + 256), -4),                             # 3-bit needs N % 512 == 0            utils.py:146-155
    (dict(tile_p=16), -5), (dict(bits=3, tile_p=64, N=1024), -5),  # utils.py:138-139
])
def test_cabi_validation_errors(kwargs, code):
    """Argument validation happens before any device is touched, so it is testable without a GPU."""
    rc = _call_qgemm(**kwargs)
    assert rc == code, (kwargs, rc, _lib.lib.flute_b200_last_error())
    assert _lib.lib.flute_b200_last_error() != b""
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_cabi_empty_batch_is_a_noop():
    assert _call_qgemm(M=0) == 0


def test_cabi_null_pointer_and_alignment():
    assert _call_qgemm(ptr=0) == -10
    assert _call_qgemm(ptr=0x1004) == -9      # TMA needs 16-byte aligned A / Q


def test_hadamard_size_validation():
    p = ctypes.c_void_p(0x1000)
    assert _lib.lib.flute_b200_hadamard(p, p, 4, 48, 0, 0, None) == -13          # not a power of two
    assert _lib.lib.flute_b200_hadamard(p, p, 4, 1 << 16, 0, 0, None) == -13     # > 2^15 (hadamard_transform.cpp:24-26)
    assert _lib.lib.flute_b200_hadamard(p, p, 0, 64, 0, 0, None) == 0            # no rows


# ---------------------------------------------------------------- templates
def test_template_table_matches_reference(golden):
    keys = [tuple(k) for k in golden["template_keys"].tolist()]
    assert len(keys) == len(templates.TEMPLATE_CONFIGS) == 216
    for k, tp, tm, tk, sm in zip(keys, golden["template_tileP"], golden["template_tileM"], golden["template_tileK"],
                                 golden["template_smsmul"]):
        cfg = templates.TEMPLATE_CONFIGS[k]
        assert (cfg["TileP"], cfg["TileM"], cfg["TileK"], cfg["SMs_Multiple"]) == (tp, tm, tk, sm), k
    assert flute.TEMPLATE_CONFIGS is templates.TEMPLATE_CONFIGS


def test_template_helpers():
    for bits in (2, 3, 4):
        tid = templates.default_template_id(bits)
        assert templates.tile_p_of(bits, tid) == 32
        assert tid in utils.get_template_ids(bits)
        cfg = utils.get_template_config(bits, tid, 148)
        assert set(cfg) == {"tileM", "tileK", "tileP", "blocks"} and cfg["blocks"] % 148 == 0
    assert len(utils.get_template_ids(4)) == 144 and len(utils.get_template_ids(3)) == 36
    with pytest.raises(RuntimeError):
        templates.tile_p_of(4, 999)
    assert utils.is_template_supported(1, 4096, 4096, 4, templates.default_template_id(4), 148)
    assert not utils.is_template_supported(1, 4096, 4096, 3, 0, 148)     # 3-bit template with TileP 64


# ---------------------------------------------------------------- packers / LUT^2
def test_torch_packers_match_reference(golden):
    from conftest import golden_cases
    for base, tok in list(golden_cases(golden, "pack_")) + list(golden_cases(golden, "packramp_")):
        bits, tp = int(tok[1][1:]), int(tok[2][2:])
        Q = utils.pack_tile_p(torch.from_numpy(golden[base + "_W"]), bits, tp)
        assert Q.dtype == torch.int16 and (Q.numpy() == golden[base + "_Q"]).all(), base


def test_pack_public_signature_and_legacy_call():
    W = torch.randint(0, 16, (64, 256), dtype=torch.uint8)
    tid = templates.default_template_id(4)
    a = utils.pack(W, 4, [tid], 148)
    b = utils.pack(W.to(torch.int64), num_bits=4, group_size=64)          # legacy: integrations/vllm_utils.py:314-317
    assert torch.equal(a, b)
    with pytest.raises(ValueError):
        utils.pack(W, 4, [0, tid], 148)                                   # mixed tile_P (utils.py:285-287)
    with pytest.raises(NotImplementedError):
        utils.pack(W.view(2, 32, 256), 4, [tid], 148)
    with pytest.raises(ValueError):
        utils.safe_cast(torch.tensor([1.5]), torch.uint8)


def test_make_qmap2_matches_reference(golden):
    for bits in (2, 3, 4):
        for name, dt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
            table = torch.from_numpy(golden[f"qmap2_b{bits}_{name}_table"].view(np.int16)).view(dt)
            got = utils.make_qmap2_from_qmap(table)
            assert got.dtype == torch.float32 and got.shape == (2 ** bits, 2 ** bits, 1)
            assert (got.numpy().view(np.uint32) == golden[f"qmap2_b{bits}_{name}_table2"].view(np.uint32)).all()
    with pytest.raises(TypeError):
        utils.make_qmap2_from_qmap(torch.zeros(16))
    with pytest.raises(ValueError):
        utils.make_qmap2_from_qmap(torch.zeros(4, 4, dtype=torch.float16))


# ---------------------------------------------------------------- torch op surface
def test_op_schema_is_the_references():
    s = str(torch.ops.flute.qgemm_raw_simple.default._schema)
    assert s == ("flute::qgemm_raw_simple(Tensor input, Tensor weight, Tensor scales, Tensor table, Tensor table2, "
                 "Tensor(a!) workspace, int num_bits, int group_size, int template_id, int num_sms) -> Tensor")
    s = str(torch.ops.flute.qgemm_raw_simple_hadamard.default._schema)
    assert "int hadamard_size, int template_id, int num_sms) -> Tensor" in s
    assert flute.qgemm is torch.ops.flute.qgemm_raw_simple
    assert callable(flute.qgemm_simple) and callable(flute.qgemm_hadamard)


def test_compiled_torch_binding_is_live():
    """INTEGRATION.md option A, built: TORCH_LIBRARY(flute) + TORCH_LIBRARY_IMPL(flute, CUDA) come from the compiled
    shim (csrc/torch_binding.cpp, counterpart of flute/csrc/qgemm.cpp:251-260) whenever it has been built."""
    import os
    from flute_b200 import ops
    shim = os.path.join(os.path.dirname(ops.__file__), "_flute_b200_torch.so")
    if not os.path.exists(shim) or os.environ.get("FLUTE_B200_PY_OPS") == "1":
        assert ops.BINDING == "python"
        pytest.skip("compiled binding not built in this tree (build.py --torch)")
    assert ops.BINDING == "cpp"
    for name in ("qgemm_raw_simple", "qgemm_raw_simple_hadamard", "hadamard_transform"):
        assert torch._C._dispatch_has_kernel_for_dispatch_key(f"flute::{name}", "CUDA"), name
    # the shim links the C-ABI library and nothing CUDA of its own
    import subprocess
    needed = subprocess.run(["readelf", "-d", shim], capture_output=True, text=True).stdout
    assert "libflute_b200.so" in needed


def _meta_args(M=(2, 3), N=256, K=128, bits=4, group=64, dtype=torch.float16):
    dev = "meta"
    return (torch.empty(M + (K,), dtype=dtype, device=dev), torch.empty((N // 16 * bits, K), dtype=torch.int16, device=dev),
            torch.empty((N, K // group), dtype=dtype, device=dev), torch.empty(2 ** bits, dtype=dtype, device=dev),
            torch.empty((2 ** bits, 2 ** bits, 1), dtype=torch.float32, device=dev),
            torch.empty(1024, dtype=torch.uint8, device=dev), bits, group, 0, 148)


def test_fake_impl_shapes_and_contract():
    """flute/ops.py:4-55: output is input.shape[:-1] + (N,); rank / dtype / shape violations raise."""
    out = flute.qgemm(*_meta_args())
    assert out.shape == (2, 3, 256) and out.dtype == torch.float16
    out = flute.qgemm_hadamard(*_meta_args()[:8], 64, 0, 148)
    assert out.shape == (2, 3, 256)
    args = list(_meta_args())
    bad = args.copy(); bad[0] = bad[0].float()
    with pytest.raises(TypeError):
        flute.qgemm(*bad)
    bad = args.copy(); bad[1] = bad[1][:-1]
    with pytest.raises(ValueError):
        flute.qgemm(*bad)
    bad = args.copy(); bad[0] = torch.empty(128, dtype=torch.float16, device="meta")      # input.ndim >= 2
    with pytest.raises(ValueError):
        flute.qgemm(*bad)
    bad = args.copy(); bad[4] = torch.empty((16, 16, 2), dtype=torch.float32, device="meta")
    with pytest.raises(ValueError):
        flute.qgemm(*bad)


def test_no_cpu_fallback():
    """The product path has no CPU implementation: CPU tensors are refused, not computed."""
    a = [t.to("cpu") if isinstance(t, torch.Tensor) else t for t in
         (torch.zeros(1, 128, dtype=torch.float16), torch.zeros(64, 128, dtype=torch.int16),
          torch.zeros(256, 2, dtype=torch.float16), torch.zeros(16, dtype=torch.float16),
          torch.zeros(16, 16, 1), torch.zeros(1024, dtype=torch.uint8))]
    with pytest.raises((NotImplementedError, RuntimeError)):
        flute.qgemm(*a, 4, 64, 0, 148)
    with pytest.raises(ValueError):
        utils.dequantize(a[1], a[2], a[4], 4, 64)


# ---------------------------------------------------------------- tune surface
def test_tune_metadata_roundtrip():
    m = tune.TuneMetaData(M=1, N=4096, K=4096, num_bits=4, group_size=64, num_sms=148, dtype=torch.bfloat16,
                          device=torch.device("cuda:0"), template_id=templates.default_template_id(4))
    d = m.to_dict()
    assert d["dtype"] == "torch.bfloat16" and d["device"] == "cuda:0"
    assert tune.TuneMetaData.from_dict(d) == m
    with pytest.raises(ValueError):
        tune.TuneMetaData.from_dict({**d, "dtype": "torch.int8"})


# ---------------------------------------------------------------- TP sharding (host logic)
@pytest.mark.parametrize("bits,tile_p,N", [(4, 32, 1024), (4, 64, 1024), (2, 32, 1024), (3, 32, 2048)])
def test_column_shard_is_a_row_slice(bits, tile_p, N):
    """Rank r's packed rows, dequantised alone, equal columns [r*N/tp, (r+1)*N/tp) of the full weight."""
    from oracle import flute_oracle as O
    K, group, world = 128, 64, 2
    rng = np.random.default_rng(bits)
    W = rng.integers(0, 1 << bits, size=(K, N), dtype=np.uint8)
    Q = torch.from_numpy(O.pack(W, bits, tile_p))
    S = torch.from_numpy(rng.standard_normal((N, K // group)).astype(np.float16))
    for r in range(world):
        Qr, Sr = parallel.shard_packed_linear(Q, S, bits, r, world, tile_p)
        assert Qr.shape == (Q.shape[0] // world, K) and Sr.shape == (N // world, K // group)
        assert (O.unpack(Qr.numpy(), bits, tile_p) == W[:, r * N // world:(r + 1) * N // world]).all()
    with pytest.raises(ValueError):
        parallel.shard_packed_linear(Q, S, bits, 0, 3, tile_p)
