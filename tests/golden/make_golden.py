"""Generate the golden wire-format vectors in this directory FROM THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference).  It imports the reference's
own CPU packers -- `flute/utils.py::_pack_{2,3,4}bit`, `make_qmap2_from_qmap` and
`flute/packbits_utils.py` -- through a 4-line stub package (the real
`flute/__init__.py` does `from . import _C`, which needs the CUDA build), exactly as
SURVEY.md appendix A describes, and stores small input/output pairs as `.npz`.

    python tests/golden/make_golden.py

The oracle (`oracle/flute_oracle.py`) and the C oracle are pinned against these files by
`tests/test_oracle_golden.py`; the CUDA path is then pinned against the oracle.
Nothing here is read on the GPU box except the committed `.npz` files.
"""
import importlib
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    tmp = tempfile.mkdtemp(prefix="flute_ref_stub_")
    pkg = os.path.join(tmp, "flute")
    os.makedirs(pkg)
    with open(os.path.join(pkg, "__init__.py"), "w") as f:
        f.write(
            "import torch\n"
            "qgemm = None\n"
            f"TEMPLATE_CONFIGS = torch.load('{REF}/flute/data/qgemm_kernel_raw_generated_configs.pth', weights_only=True)\n"
        )
    for name in ("utils.py", "packbits_utils.py"):
        shutil.copy(os.path.join(REF, "flute", name), os.path.join(pkg, name))
    sys.path.insert(0, tmp)
    flute = importlib.import_module("flute")
    utils = importlib.import_module("flute.utils")
    return flute, utils


def main():
    flute, utils = import_reference()
    rng = np.random.default_rng(20260923)
    out = {}

    # ---- packers: (bits, tile_P, K, N) ----
    cases = [
        (4, 32, 8, 512), (4, 64, 8, 512), (4, 32, 64, 128), (4, 32, 128, 256),
        (2, 32, 8, 1024), (2, 64, 8, 1024), (2, 32, 64, 256),
        (3, 32, 8, 1024), (3, 32, 64, 512), (3, 32, 128, 1536),
    ]
    for bits, tp, K, N in cases:
        W = rng.integers(0, 1 << bits, size=(K, N), dtype=np.uint8)
        Wt = torch.from_numpy(W.copy())
        fn = {4: utils._pack_4bit, 2: utils._pack_2bit, 3: utils._pack_3bit}[bits]
        Q = fn(Wt, tile_P=tp)
        assert Q.dtype == torch.int16 and Q.shape == (N // 16 * bits, K), (Q.dtype, Q.shape)
        key = f"pack_b{bits}_tp{tp}_k{K}_n{N}"
        out[key + "_W"] = W
        out[key + "_Q"] = Q.numpy().copy()

    # ---- structured patterns: one-hot indices make every field position visible ----
    for bits, tp, K, N in [(4, 32, 4, 128), (2, 32, 4, 256), (3, 32, 4, 512)]:
        W = np.zeros((K, N), dtype=np.uint8)
        W[np.arange(K)[:, None] % K, np.arange(N)[None, :]] = (np.arange(N)[None, :] + np.arange(K)[:, None]) % (1 << bits)
        fn = {4: utils._pack_4bit, 2: utils._pack_2bit, 3: utils._pack_3bit}[bits]
        Q = fn(torch.from_numpy(W.copy()), tile_P=tp)
        key = f"packramp_b{bits}_tp{tp}_k{K}_n{N}"
        out[key + "_W"] = W
        out[key + "_Q"] = Q.numpy().copy()

    # ---- table2 from table ----
    for bits in (2, 3, 4):
        for dt, name in ((torch.float16, "f16"), (torch.bfloat16, "bf16")):
            table = torch.from_numpy(rng.standard_normal(1 << bits).astype(np.float32)).to(dt)
            t2 = utils.make_qmap2_from_qmap(table)
            out[f"qmap2_b{bits}_{name}_table"] = table.view(torch.int16).numpy().view(np.uint16).copy()
            out[f"qmap2_b{bits}_{name}_table2"] = t2.numpy().copy()

    # ---- the reference tests' ground truth, evaluated with torch on CPU, literally as
    #      tests/kernel.py:68-71 / tune.py:332-335 write it:
    #          W_ = qmap[W]; S_ = repeat_interleave(S, group, dim=1).T; D_ = torch.mm(A, W_ * S_)
    #      and tests/higgs.py:7-17 (vector_dequantize_higgs) for the HIGGS pair grid.
    gt_cases = [
        # bits, group, dtype, M, K, N
        (4, 64, torch.float16, 3, 128, 256), (4, 64, torch.bfloat16, 5, 256, 128),
        (4, 128, torch.float16, 1, 256, 128), (4, 256, torch.bfloat16, 2, 512, 128),
        (3, 64, torch.float16, 3, 128, 512), (3, 128, torch.bfloat16, 1, 256, 512),
        (2, 64, torch.float16, 4, 128, 256), (2, 64, torch.bfloat16, 2, 192, 512),
    ]
    for idx, (bits, group, dt, M, K, N) in enumerate(gt_cases):
        torch.manual_seed(idx)
        A = torch.randn((M, K), dtype=dt) / 100.
        W = torch.randint(0, 2 ** bits, (K, N), dtype=torch.int64)
        S = torch.randn((N, K // group), dtype=dt)
        qmap = torch.randn(2 ** bits, dtype=dt)
        W_ = qmap[W]
        S_ = torch.repeat_interleave(S, group, dim=1).T
        What = W_ * S_
        D_ = torch.mm(A, What)
        fn = {4: utils._pack_4bit, 2: utils._pack_2bit, 3: utils._pack_3bit}[bits]
        Q = fn(W.to(torch.uint8), tile_P=32)
        name = "f16" if dt == torch.float16 else "bf16"
        key = f"gt{idx}_b{bits}_g{group}_{name}_m{M}_k{K}_n{N}"
        bits16 = lambda t: t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()
        out[key + "_A"] = bits16(A)
        out[key + "_W"] = W.to(torch.uint8).numpy().copy()
        out[key + "_Q"] = Q.numpy().copy()
        out[key + "_S"] = bits16(S)
        out[key + "_table"] = bits16(qmap)
        out[key + "_table2"] = utils.make_qmap2_from_qmap(qmap).numpy().copy()
        out[key + "_What"] = bits16(What)
        out[key + "_D"] = bits16(D_)

    for idx, (bits, dt) in enumerate([(4, torch.float16), (3, torch.bfloat16), (2, torch.float16)]):
        torch.manual_seed(100 + idx)
        N, K, group = 512, 128, 64
        num_codes = 2 ** (bits * 2)
        codes = torch.randint(0, num_codes, (N, K // 2), dtype=torch.uint8)
        scales = torch.randn((N, K // group), dtype=dt)
        grid = torch.randn((num_codes, 2), dtype=dt)
        w = grid[codes.int()]                                    # tests/higgs.py:12
        w = w.reshape(w.shape[0], -1, group) * scales[..., None]
        w = w.reshape(w.shape[0], -1)                            # [N, K]
        name = "f16" if dt == torch.float16 else "bf16"
        key = f"higgs{idx}_b{bits}_{name}"
        bits16 = lambda t: t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()
        out[key + "_codes"] = codes.numpy().copy()
        out[key + "_scales"] = bits16(scales)
        out[key + "_grid"] = bits16(grid)
        out[key + "_dense"] = bits16(w)

    # ---- template id -> TileP map (which packing a stored checkpoint uses, utils.py:302-309) ----
    cfgs = flute.TEMPLATE_CONFIGS
    keys = sorted(cfgs.keys())
    out["template_keys"] = np.array(keys, dtype=np.int32)
    out["template_tileP"] = np.array([cfgs[k]["TileP"] for k in keys], dtype=np.int32)
    out["template_tileM"] = np.array([cfgs[k]["TileM"] for k in keys], dtype=np.int32)
    out["template_tileK"] = np.array([cfgs[k]["TileK"] for k in keys], dtype=np.int32)
    out["template_smsmul"] = np.array([cfgs[k]["SMs_Multiple"] for k in keys], dtype=np.int32)

    path = os.path.join(HERE, "wire_format.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
