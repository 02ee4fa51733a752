"""CPU test of the autograd glue behind ``async_cache=True`` (TTMaskedLookupFunction): the ops it
calls are replaced by oracle-backed stand-ins with the shim's signatures, so argument order, the mask
convention (-1: TT cores, >= 0: cache row, -2: dropped) and the positions of the returned gradients are pinned
without a GPU.  The kernels behind the real ops are covered by tests/test_zz2_gpu_async_cache.py."""
import numpy as np
import pytest
import torch

from oracle import tt_oracle as O
from tests.helpers import make_cores, ragged_batch

P, Q, RANKS = [5, 6, 7], [2, 2, 4], [3, 5]
R = [1] + RANKS + [1]
D, B, C = 16, 12, 9


class _Ops:
    """tt_forward / cache_forward / *_backward with the signatures of fbtt_embedding_b200.tt_embeddings."""

    def __init__(self):
        self.calls = []

    @staticmethod
    def _tt_part(nnz, indices, rowidx, loc):
        keep = loc.numpy()[:nnz] == -1
        return indices.numpy()[:nnz][keep], rowidx.numpy()[:nnz][keep]

    def set_keep_plans(self, flag):
        return True

    def tt_forward(self, batch_count, num_tables, Bx, Dx, p, q, ranks, L, nnz, indices, rowidx, tableidx, cores,
                   cache_locations=None, keep_plan=True):
        self.calls.append("tt_forward")
        assert (batch_count, num_tables, Bx, Dx, list(p), list(q), list(ranks)) == (1000, 1, B, D, P, Q, R)
        idx, row = self._tt_part(nnz, indices, rowidx, cache_locations)
        out = O.tt_forward(1, B, D, P, Q, RANKS, O.make_L(P), len(idx), idx, row, np.zeros(len(idx), np.int64),
                           [c.detach().numpy() for c in cores])
        return torch.from_numpy(out)

    def cache_forward(self, Bx, nnz, loc, rowidx, cache_weight, output):
        self.calls.append("cache_forward")
        keep = loc.numpy()[:nnz] >= 0
        o = output.numpy()[0]
        O.cache_forward(loc.numpy()[:nnz][keep], rowidx.numpy()[:nnz][keep], cache_weight.detach().numpy(), o)

    def _dense(self, nnz, indices, rowidx, d_output, cores, loc):
        idx, row = self._tt_part(nnz, indices, rowidx, loc)
        return O.tt_backward_dense(D, P, Q, RANKS, O.make_L(P), len(idx), idx, row, np.zeros(len(idx), np.int64),
                                   d_output.numpy(), [c.detach().numpy() for c in cores])

    def tt_dense_backward(self, batch_count, Dx, p, q, ranks, L, nnz, indices, rowidx, tableidx, d_output, cores,
                          cache_locations=None):
        self.calls.append("tt_dense_backward")
        return [torch.from_numpy(g.astype(np.float32)) for g in self._dense(nnz, indices, rowidx, d_output, cores, cache_locations)]

    def tt_sgd_backward(self, batch_count, Dx, lr, p, q, ranks, L, nnz, indices, rowidx, tableidx, d_output, cores,
                        cache_locations=None):
        self.calls.append("tt_sgd_backward")
        g = self._dense(nnz, indices, rowidx, d_output, cores, cache_locations)
        with torch.no_grad():
            for c, gi in zip(cores, g):
                c -= lr * torch.from_numpy(gi.astype(np.float32))

    def cache_backward_sgd(self, nnz, d_output, loc, rowidx, lr, cache_weight):
        self.calls.append("cache_backward_sgd")
        keep = loc.numpy()[:nnz] >= 0
        O.cache_backward_sgd(d_output.numpy()[0], loc.numpy()[:nnz][keep], rowidx.numpy()[:nnz][keep], lr,
                             cache_weight.detach().numpy())

    def cache_backward_dense(self, nnz, d_output, loc, rowidx, lr, cache_weight):
        self.calls.append("cache_backward_dense")
        keep = loc.numpy()[:nnz] >= 0
        return torch.from_numpy(O.cache_backward_dense(d_output.numpy()[0], loc.numpy()[:nnz][keep],
                                                       rowidx.numpy()[:nnz][keep], cache_weight.detach().numpy()))


def _setup(monkeypatch, seed):
    from fbtt_embedding_b200 import tt_embeddings_ops as ops

    fake = _Ops()
    monkeypatch.setattr(ops, "tt_embeddings", fake)
    rng = np.random.RandomState(seed)
    cores = [torch.nn.Parameter(torch.from_numpy(c)) for c in make_cores(rng, 1, P, Q, RANKS)]
    cache_weight = torch.nn.Parameter(torch.from_numpy(rng.uniform(-1, 1, (C, D)).astype(np.float32)))
    idx, off = ragged_batch(rng, B, int(np.prod(P)), 4.0, 2.0)
    row, _ = O.compute_rowidx(off, 1)
    loc = rng.randint(-1, C, size=len(idx)).astype(np.int32)  # a mix of TT (-1) and cached lookups
    loc[::7] = -2                                             # and a few entries no half may touch
    d_out = torch.from_numpy(rng.uniform(-1, 1, (1, B, D)).astype(np.float32))
    args = dict(idx=torch.from_numpy(idx), row=torch.from_numpy(row), tbl=torch.zeros(len(idx), dtype=torch.int64),
                loc=torch.from_numpy(loc), d_out=d_out)
    return ops, fake, cores, cache_weight, args


def _want_forward(cores, cache_weight, a):
    idx, row, loc = a["idx"].numpy(), a["row"].numpy(), a["loc"].numpy()
    tt = loc == -1
    out = O.tt_forward(1, B, D, P, Q, RANKS, O.make_L(P), int(tt.sum()), idx[tt], row[tt], np.zeros(int(tt.sum()), np.int64),
                       [c.detach().numpy() for c in cores])
    O.cache_forward(loc[loc >= 0], row[loc >= 0], cache_weight.detach().numpy(), out[0])
    return out


def test_masked_lookup_dense_mode_gradient_positions(monkeypatch):
    ops, fake, cores, cw, a = _setup(monkeypatch, 0)
    want = _want_forward(cores, cw, a)
    out = ops.TTMaskedLookupFunction.apply(B, D, P, Q, R, torch.tensor(O.make_L(P)), a["idx"], a["row"], a["tbl"], a["loc"],
                                           ops.OptimType.SGD, 0.1, 1e-8, False, None, cw, [], *cores)
    np.testing.assert_allclose(out.detach().numpy(), want, rtol=1e-5, atol=1e-6)
    out.backward(a["d_out"])
    idx, row, loc = a["idx"].numpy(), a["row"].numpy(), a["loc"].numpy()
    tt = loc == -1
    g = O.tt_backward_dense(D, P, Q, RANKS, O.make_L(P), int(tt.sum()), idx[tt], row[tt], np.zeros(int(tt.sum()), np.int64),
                            a["d_out"].numpy(), [c.detach().numpy() for c in cores])
    for c, gi in zip(cores, g):
        np.testing.assert_allclose(c.grad.numpy(), gi, rtol=1e-5, atol=1e-6)
    gcw = O.cache_backward_dense(a["d_out"].numpy()[0], loc[loc >= 0], row[loc >= 0], cw.detach().numpy())
    np.testing.assert_allclose(cw.grad.numpy(), gcw, rtol=1e-6, atol=1e-7)
    assert fake.calls == ["tt_forward", "cache_forward", "tt_dense_backward", "cache_backward_dense"]


def test_masked_lookup_fused_sgd(monkeypatch):
    ops, fake, cores, cw, a = _setup(monkeypatch, 1)
    lr = 0.05
    c0 = [c.detach().numpy().copy() for c in cores]
    cw0 = cw.detach().numpy().copy()
    out = ops.TTMaskedLookupFunction.apply(B, D, P, Q, R, torch.tensor(O.make_L(P)), a["idx"], a["row"], a["tbl"], a["loc"],
                                           ops.OptimType.SGD, lr, 1e-8, True, None, cw, [], *cores)
    out.backward(a["d_out"])
    assert all(c.grad is None for c in cores) and cw.grad is None  # fused: updates happen inside the ops
    idx, row, loc = a["idx"].numpy(), a["row"].numpy(), a["loc"].numpy()
    tt = loc == -1
    g = O.tt_backward_dense(D, P, Q, RANKS, O.make_L(P), int(tt.sum()), idx[tt], row[tt], np.zeros(int(tt.sum()), np.int64),
                            a["d_out"].numpy(), c0)
    for c, w in zip(cores, O.sgd_step(c0, g, lr)):
        np.testing.assert_allclose(c.detach().numpy(), w, rtol=1e-5, atol=1e-6)
    O.cache_backward_sgd(a["d_out"].numpy()[0], loc[loc >= 0], row[loc >= 0], lr, cw0)
    np.testing.assert_allclose(cw.detach().numpy(), cw0, rtol=1e-6, atol=1e-7)
    assert fake.calls == ["tt_forward", "cache_forward", "tt_sgd_backward", "cache_backward_sgd"]


def test_shim_rejects_a_bad_mask():
    from fbtt_embedding_b200 import tt_embeddings as ext

    with pytest.raises(RuntimeError):
        ext._mask(torch.zeros(4, dtype=torch.int64), 4)  # not int32 / not CUDA
    assert ext._mask(None, 4) is None
