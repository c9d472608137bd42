"""NF quantiser and FluteLinear host logic, CPU only.

flute_b200/nf_utils.py is pinned bit for bit to vectors the reference's own flute/nf_utils.py produced
(tests/golden/make_golden_nf.py); FluteLinear / prepare_model_flute are checked for the reference's module contract
(buffer names and shapes, extra state, state-dict round trip) and for consistency of the packed weight with the
quantiser's indices through the oracle."""
import os

import numpy as np
import pytest
import torch

from flute_b200 import nf_utils, utils
from flute_b200.integrations import FluteLinear, prepare_model_flute

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "nf_quant.npz"))


@pytest.mark.parametrize("bits", [2, 3, 4])
def test_code_books_match_reference(bits, gold):
    v, p = nf_utils.get_values_pivots(bits, False)
    assert np.array_equal(v.numpy(), gold[f"values_{bits}"]) and np.array_equal(p.numpy(), gold[f"pivots_{bits}"])
    assert v.numel() == 2 ** bits and p.numel() == 2 ** bits - 1 and float(v.abs().max()) == 1.0


def test_nf_quantize_matches_reference(gold):
    keys = sorted({k.rsplit("_", 1)[0] for k in gold.files if k.startswith("q_") and k.endswith("_W")})
    assert len(keys) == 4
    for key in keys:
        _, bits, group, N, K = key.split("_")
        bits, group = int(bits), int(group)
        W = torch.from_numpy(gold[key + "_W"])
        dq, idx, absmax, values = nf_utils.nf_quantize(W, bits, group)
        assert np.array_equal(idx.numpy().astype(np.int16), gold[key + "_idx"]), key
        assert np.array_equal(absmax.numpy(), gold[key + "_absmax"]), key
        assert np.array_equal(dq.numpy(), gold[key + "_dq"]), key
        for name, dt in (("bf16", torch.bfloat16), ("f16", torch.float16)):
            fq = nf_utils.nf_quantize_2(W.to(dt), bits, group, dt)
            assert np.array_equal(fq.contiguous().view(torch.int16).numpy(), gold[key + f"_fq2_{name}"]), (key, name)


def test_custom_scales_and_validation():
    W = torch.randn(8, 128)
    s = torch.full((8 * 2,), 3.0)
    dq, idx, absmax, _ = nf_utils.nf_quantize(W, 4, 64, custom_scales=s)
    assert torch.equal(absmax, s) and idx.max() < 16
    with pytest.raises(ValueError):
        nf_utils.nf_quantize(torch.randn(4, 100), 4, 64)


@pytest.mark.parametrize("bits,group", [(4, 64), (3, 64), (2, 128)])
def test_prepare_model_flute_module_contract(bits, group):
    """Buffer names / shapes / dtypes of flute/integrations/base.py:203-326, packed weight == pack(indices)."""
    from oracle import flute_oracle
    N, K = (1024, 256) if bits != 3 else (1024, 256)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(K, N, bias=True), torch.nn.GELU(), torch.nn.Linear(N, 512, bias=False)).to(torch.float16)
    W0 = model[0].weight.detach().clone()
    prepare_model_flute("model", model, bits, group)
    lin = model[0]
    assert isinstance(lin, FluteLinear) and isinstance(model[2], FluteLinear)
    assert lin.weight.shape == (N // 16 * bits, K) and lin.weight.dtype == torch.int16
    assert lin.scales.shape == (N, K // group) and lin.scales.dtype == torch.float16
    assert lin.tables.shape == (2 ** bits,) and lin.tables2.shape == (2 ** bits, 2 ** bits, 1) and lin.tables2.dtype == torch.float32
    assert lin.bias is not None and model[2].bias is None
    assert lin.get_extra_state() == {"num_bits": bits, "group_size": group, "template_id": lin.template_id}
    # the packed weight decodes (oracle) to the quantiser's dequantised weight in the kernel's arithmetic
    W_hat = flute_oracle.dequantize(lin.weight.numpy(), lin.scales.numpy(),
                                    lin.tables2.numpy(), bits, group, "float16")                   # [K, N] uint16 bits
    W_hat = torch.from_numpy(W_hat.view(np.int16)).view(torch.float16)
    _, idx, absmax, values = nf_utils.nf_quantize(W0, bits, group)
    expect = (values.to(torch.float16)[idx] * absmax.to(torch.float16).view(N, K // group).repeat_interleave(group, 1))
    assert torch.equal(W_hat.T.contiguous().view(torch.int16), expect.contiguous().view(torch.int16))
    # state dict round trip into a freshly constructed module (meta-free path the HF loader uses)
    sd = model.state_dict()
    fresh = FluteLinear(K, N, bits, group, lin.template_id, bias=True, device=torch.device("cpu"), dtype=torch.float16)
    fresh.load_state_dict({k[2:]: v for k, v in sd.items() if k.startswith("0.")})
    assert torch.equal(fresh.weight, lin.weight) and torch.equal(fresh.scales, lin.scales)
    bad = FluteLinear(K, N, bits, 256 if group != 256 else 64, lin.template_id, bias=True, device=torch.device("cpu"), dtype=torch.float16)
    with pytest.raises((ValueError, RuntimeError)):
        bad.load_state_dict({k[2:]: v for k, v in sd.items() if k.startswith("0.")})


def test_prepare_model_flute_fake_is_kernel_arithmetic():
    torch.manual_seed(1)
    m = torch.nn.Sequential(torch.nn.Linear(128, 128, bias=False)).to(torch.bfloat16)
    W = m[0].weight.detach().clone()
    prepare_model_flute("m", m, 4, 64, fake=True)
    assert isinstance(m[0], torch.nn.Linear)
    assert torch.equal(m[0].weight, nf_utils.nf_quantize_2(W, 4, 64, torch.bfloat16))
