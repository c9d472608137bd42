"""Golden vectors of the reference's NF quantiser (`flute/nf_utils.py`), generated FROM THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference).  The reference hard-codes `.cuda()` (nf_utils.py:32); it is
imported here with `torch.Tensor.cuda` patched to the identity so that its own code runs on the CPU unchanged.  Stored:
code books and pivots for 2/3/4 bits, and (indices, scales, fake-quantised weights of both flavours) for seeded fp32 /
bf16 / fp16 weights.  tests/test_nf_quant.py pins flute_b200/nf_utils.py to them bit for bit.

    python tests/golden/make_golden_nf.py
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/flute/nf_utils.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self          # the only change: no device move
    spec = importlib.util.spec_from_file_location("ref_nf_utils", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    for bits in (2, 3, 4):
        v, p = ref.get_values_pivots(bits, False)
        out[f"values_{bits}"] = v.numpy()
        out[f"pivots_{bits}"] = p.numpy()
    g = torch.Generator().manual_seed(20260923)
    for bits, group, (N, K) in ((4, 64, (96, 256)), (3, 128, (64, 256)), (2, 64, (32, 128)), (4, 128, (48, 384))):
        W = torch.randn((N, K), generator=g) * 0.05
        dq, idx, absmax, values = ref.nf_quantize(W, bits, group)
        key = f"q_{bits}_{group}_{N}_{K}"
        out[key + "_W"] = W.numpy()
        out[key + "_idx"] = idx.numpy().astype(np.int16)
        out[key + "_absmax"] = absmax.numpy()
        out[key + "_dq"] = dq.numpy()
        for name, dt in (("bf16", torch.bfloat16), ("f16", torch.float16)):
            fq = ref.nf_quantize_2(W.to(dt), bits, group, dt)
            out[key + f"_fq2_{name}"] = fq.contiguous().view(torch.int16).numpy()
    np.savez_compressed(os.path.join(HERE, "nf_quant.npz"), **out)
    print("wrote", os.path.join(HERE, "nf_quant.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
