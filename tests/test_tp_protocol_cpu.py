"""Host side of the fused tensor-parallel exchange, on CPU (no GPU, no process group).

`flute_b200.parallel.FusedGather` fills one `flute_b200_tp` descriptor per call: peer pointers of the plain and the word
image, the sequence-number arithmetic (uses / calls per step, the step counter), the word-image address of activations
that live inside a gathered buffer (column sub-views included).  Here both ranks of a tp = 2 group live in ONE process,
their "symmetric" buffers are ordinary host tensors, and `flute_b200_qgemm_tp` is replaced by an executable restatement
of what the kernel does with the descriptor (include/flute_b200.h, csrc/qgemm_decode_sm100.cu: store_out / the
word-reading activation warp): read A from the word image and INSIST on the expected sequence number in every word,
compute this rank's column slice with the oracle, store {value, sequence} words into every rank's image.  A wrong
`uses` / `call` / offset / stride anywhere makes a sequence check fail; the gathered result of the sharded chain must
equal the unsharded oracle chain bit for bit, step after step (images are re-used, sequence numbers grow)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flute_b200 import _lib, parallel, utils          # noqa: E402
from helpers import bits16, from_bits16                # noqa: E402


def _u16(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_uint16 * n).from_address(ptr))


def _u32(ptr, n):
    return np.ctypeslib.as_array((ctypes.c_uint32 * n).from_address(ptr))


class _KernelModel:
    """Stands in for libflute_b200's tensor-parallel entry points; records what it was asked to do."""

    def __init__(self):
        self.calls = []
        self.published = []
        self.waits = []

    # --- the three tiny kernels -------------------------------------------------------------
    def flute_b200_tp_advance(self, epoch_ptr, device, stream):
        _u32(epoch_ptr, 1)[0] += 1
        return 0

    def flute_b200_tp_publish(self, flags, tp, device, stream):
        for r in range(tp):
            _u32(flags[r], 1)[0] += 1
        self.published.append(tp)
        return 0

    def flute_b200_tp_wait(self, flag, per_step, offset, epoch_ptr, device, stream):
        # on the GPU this spins until every rank's publish has arrived; the ranks of this model run one after the other, so
        # the condition is checked once all of them have issued their end_step (waits_satisfied)
        self.waits.append((flag, (int(_u32(epoch_ptr, 1)[0]) - 1) * per_step + offset))
        return 0

    def waits_satisfied(self):
        ok = all(int(_u32(flag, 1)[0]) >= expected for flag, expected in self.waits)
        self.waits.clear()
        return ok

    # --- the GEMM with the exchange in its epilogue ----------------------------------------------
    def flute_b200_qgemm_tp(self, A, Q, S, table, table2, ws, ws_bytes, M, N, K, bits, group, tile_p, code, flags, device,
                            stream, desc_ref):
        from oracle import c_oracle
        d = desc_ref._obj
        assert 1 <= d.tp <= 8 and 0 <= d.rank < d.tp and d.n_total == d.tp * N
        epoch = int(_u32(d.epoch, 1)[0])
        assert epoch >= 1, "begin_step (flute_b200_tp_advance) must run before the step's first call"
        if d.in_ll:
            assert d.in_ll % 16 == 0 and d.in_ll_stride % 2 == 0       # the launcher's alignment checks
            expected = (epoch - 1) * d.in_uses + d.in_call + 1
            a = np.empty((M, K), dtype=np.uint16)
            for m in range(M):
                words = _u32(d.in_ll + 8 * m * d.in_ll_stride, 2 * K).reshape(K, 2)
                assert (words[:, 1] == expected).all(), \
                    f"activation words carry sequence {set(words[:, 1].tolist())}, the consumer waits for {expected}"
                a[m] = words[:, 0].astype(np.uint16)
        else:
            a = _u16(A, M * K).reshape(M, K).copy()
        P = N // 16 * bits
        q = np.ctypeslib.as_array((ctypes.c_int16 * (P * K)).from_address(Q)).reshape(P, K)
        s = _u16(S, N * (K // group)).reshape(N, K // group)
        t2 = np.ctypeslib.as_array((ctypes.c_float * (4 ** bits)).from_address(table2)).reshape(2 ** bits, 2 ** bits, 1)
        out = c_oracle.qgemm(a, q, s, t2, bits, group, code == _lib.BF16, tile_p)          # [M, N] bit patterns
        seq = (epoch - 1) * d.out_uses + d.out_call + 1
        ll_delta = d.ll_peers[0] - d.out_peers[0]
        for r in range(d.tp):
            assert d.ll_peers[r] - d.out_peers[r] == ll_delta          # one layout on every rank
            words = _u32(d.out_peers[r] + ll_delta, 2 * M * d.n_total).reshape(M, d.n_total, 2)
            words[:, d.rank * N:(d.rank + 1) * N, 0] = out
            words[:, d.rank * N:(d.rank + 1) * N, 1] = seq
            if d.write_plain:
                _u16(d.out_peers[r], M * d.n_total).reshape(M, d.n_total)[:, d.rank * N:(d.rank + 1) * N] = out
        self.calls.append(dict(rank=d.rank, seq=seq, in_ll=bool(d.in_ll), plain=bool(d.write_plain)))
        return 0


class _Shim:
    """What FusedGather uses of flute_b200._lib."""
    TpDesc, BF16, F16 = _lib.TpDesc, _lib.BF16, _lib.F16

    def __init__(self):
        self.lib = _KernelModel()

    @staticmethod
    def check(rc):
        assert rc == 0


class _HostGather(parallel.FusedGather):
    """FusedGather over host memory: every rank's buffer lives in this process, `registry` is the peer mapping."""

    def __init__(self, registry, shim, rank, tp, outputs, dtype):
        self._registry = registry
        super().__init__(torch.device("cpu"), rank, tp, outputs, dtype, group="one process")
        self._lib = shim

    def _allocate(self, nbytes, group):
        if not self._registry:
            self._registry.extend(torch.zeros(nbytes, dtype=torch.uint8) for _ in range(self.tp))
        assert all(b.numel() == nbytes for b in self._registry)       # every rank computes the same layout
        return self._registry[self.rank], [b.data_ptr() for b in self._registry]

    def _stream(self):
        return 0


def _linear(g, N, K, bits, group, dtype):
    W = torch.randint(0, 2 ** bits, (K, N), generator=g, dtype=torch.int64).to(torch.uint8)
    S = (torch.randn((N, K // group), generator=g) / K ** 0.5).to(dtype)
    return utils.pack_tile_p(W, bits, 32), S


@pytest.mark.parametrize("M", [1, 3])
def test_fused_gather_descriptors_drive_a_correct_exchange(M):
    from oracle import c_oracle
    tp, bits, group, layers, steps, dtype = 2, 4, 64, 2, 3, torch.float16
    # `up` is wider than the K of `down`: down reads a column sub-view of the gathered buffer (offset / stride path)
    shapes = [("up", 1024, 512), ("down", 512, 768)]
    g = torch.Generator().manual_seed(5)
    table = torch.randn(2 ** bits, generator=g).to(dtype)
    table2 = utils.make_qmap2_from_qmap(table)
    weights = [{name: _linear(g, N, K, bits, group, dtype) for name, N, K in shapes} for _ in range(layers)]
    shards = [[{name: parallel.shard_packed_linear(*lin[name], bits, r, tp, 32) for name, _, _ in shapes} for lin in weights]
              for r in range(tp)]
    ws = torch.zeros(1024, dtype=torch.uint8)
    registry, shim = [], _Shim()
    fgs = [_HostGather(registry, shim, r, tp, [(name, M, N, layers) for name, N, K in shapes], dtype) for r in range(tp)]

    def reference(x0):
        x = x0
        for lin in weights:
            for name, N, K in shapes:
                Q, S = lin[name]
                x = from_bits16(c_oracle.qgemm(bits16(x[:, :K].contiguous()), Q.numpy(), bits16(S), table2.numpy(), bits, group,
                                               False, 32), dtype)
        return x

    for step in range(steps):
        x0 = (torch.randn((M, shapes[0][2]), generator=g) / 10).to(dtype)
        for fg in fgs:
            fg.begin_step()
        xs = [x0.clone() for _ in range(tp)]
        for li in range(layers):
            for name, N, K in shapes:
                last = li == layers - 1 and name == shapes[-1][0]
                nxt = []
                for r, fg in enumerate(fgs):          # every rank issues the same call; lock step
                    Q, S = shards[r][li][name]
                    nxt.append(fg.qgemm(xs[r][:, :K], Q, S, table, table2, ws, name, N // tp, K, bits, group, 0, plain=last))
                xs = nxt
        for fg in fgs:
            fg.end_step(shapes[-1][0])
        assert shim.lib.waits_satisfied(), "end_step would spin for ever: some rank's publish is missing"
        want = reference(x0)
        for r in range(tp):
            assert torch.equal(xs[r].view(torch.int16), want.view(torch.int16)), f"step {step}, rank {r}: gathered result differs"
    calls = shim.lib.calls
    assert len(calls) == steps * layers * len(shapes) * tp
    assert [c["in_ll"] for c in calls[:tp]] == [False] * tp and all(c["in_ll"] for c in calls[tp:2 * tp])   # x0 is local
    assert sum(c["plain"] for c in calls) == steps * tp          # only the step's final output keeps a plain image
    assert max(c["seq"] for c in calls) == steps * layers        # (epoch - 1) * uses + call + 1 after the last step


def test_fused_gather_rejects_misuse():
    registry, shim = [], _Shim()
    fg = _HostGather(registry, shim, 0, 1, [("a", 1, 512, 1), ("b", 1, 512, 2)], torch.float16)
    g = torch.Generator().manual_seed(1)
    Q, S = _linear(g, 512, 512, 4, 64, torch.float16)
    table = torch.randn(16, generator=g).to(torch.float16)
    table2 = utils.make_qmap2_from_qmap(table)
    ws = torch.zeros(64, dtype=torch.uint8)
    x0 = torch.randn((1, 512), generator=g).to(torch.float16)
    call = lambda x, name: fg.qgemm(x, Q, S, table, table2, ws, name, 512, 512, 4, 64, 0)
    fg.begin_step()
    with pytest.raises(ValueError, match="not written in this step"):
        call(fg.out["b"]["view"], "a")                    # reads a gathered buffer nobody has written yet
    ya = call(x0, "a")
    with pytest.raises(ValueError, match="more than its declared"):
        call(x0, "a")                                     # `a` was declared with one write per step
    yb = call(ya, "b")
    with pytest.raises(ValueError, match="both the activations and the output"):
        call(yb, "b")                                     # would overwrite the words it is still reading
    with pytest.raises(ValueError, match="shape mismatch"):
        fg.qgemm(x0, Q, S, table, table2, ws, "b", 256, 512, 4, 64, 0)
    fg.end_step("b")
    with pytest.raises(ValueError, match="already published"):
        fg.end_step("b")
