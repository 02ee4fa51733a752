"""CPU test of fbtt_embedding_b200/grouped.py's host logic: the item array it hands to
``ttb_group_*`` -- shapes, counts, every pointer into the shared COO / output / gradient / plan buffers -- is
executed here by a stand-in for libttb that walks the array exactly as include/ttb.h describes and computes
each item with the numpy oracle on the (host) memory the pointers name.  What the real library does with the
same array is covered on the GPU (tests/test_zz1_gpu_group.py); this pins the Python side where it can run."""
import contextlib
import ctypes

import numpy as np
import pytest
import torch
from torch import nn

from oracle import tt_oracle as O
from tests.helpers import ragged_batch

D = 16
SPECS = [
    dict(p=[5, 6, 7], q=[2, 2, 4], ranks=[3, 5]),
    dict(p=[20, 22, 25], q=[2, 2, 4], ranks=[4, 4]),
    dict(p=[9, 8], q=[4, 4], ranks=[6]),
    dict(p=[2, 3, 2, 3], q=[2, 2, 2, 2], ranks=[2, 3, 2]),
    dict(p=[6, 5, 4], q=[4, 1, 4], ranks=[16, 16]),  # warp-MMA family: this item carries a bucketing-plan workspace
    dict(p=[3, 2, 2], q=[4, 1, 4], ranks=[16, 16]),  # and a second one behind it in the same flat plan buffer
]


def _np(ptr, n, ctype, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,)).view(dtype)


class _FakeTable(nn.Module):
    """The attributes of TTEmbeddingBag that GroupedLookup reads, on the CPU."""

    def __init__(self, spec, optimizer, sparse, lr, eps, rng):
        super().__init__()
        from fbtt_embedding_b200.tt_embeddings_ops import _SGD_FAMILY

        p, q, ranks = spec["p"], spec["q"], spec["ranks"]
        R = [1] + ranks + [1]
        self.num_tables, self.use_cache, self.embedding_dim = 1, False, D
        self.sparse, self.optimizer, self.learning_rate, self.eps = sparse, optimizer, lr, eps
        self.tt_p_shapes, self.tt_q_shapes, self.tt_ranks = p, q, R
        self.tt_cores = nn.ParameterList(
            nn.Parameter(torch.from_numpy(rng.uniform(-0.5, 0.5, (1, p[t], R[t] * q[t] * R[t + 1])).astype(np.float32)))
            for t in range(len(p)))
        self.optimizer_state = [torch.zeros(c.shape if optimizer not in _SGD_FAMILY else (0,)) for c in self.tt_cores]


class _FakeLib:
    """ttb_group_* re-stated over host memory with the oracle; everything else is the real libttb."""

    def __init__(self, real):
        self.real = real
        self.calls = []

    def __getattr__(self, name):
        return getattr(self.real, name)

    @staticmethod
    def _dims(it):
        s = it.shape
        T = s.T
        return T, list(s.p)[:T], list(s.q)[:T], list(s.R)[1:T], s.B, s.D, list(s.L)[:T]

    def _cores(self, it, field):
        T, p, q, ranks, B, Dm, L = self._dims(it)
        R = [1] + ranks + [1]
        return [_np(getattr(it, field)[t], p[t] * R[t] * q[t] * R[t + 1], ctypes.c_float, np.float32).reshape(
            1, p[t], -1) for t in range(T)]

    def ttb_group_preprocess(self, n, items, stream):
        self.calls.append("preprocess")
        for i in range(n):
            it = items[i]
            if it.nnz == 0:
                continue
            off = _np(it.offsets, it.shape.num_tables * it.shape.B + 1, ctypes.c_int64, np.int64)
            row, tbl = O.compute_rowidx(off, it.shape.num_tables)
            _np(it.rowidx, it.nnz, ctypes.c_int64, np.int64)[:] = row
            _np(it.tableidx, it.nnz, ctypes.c_int64, np.int64)[:] = tbl
        return 0

    def ttb_group_forward(self, n, items, stream):
        self.calls.append("forward")
        for i in range(n):
            it = items[i]
            if it.nnz == 0:
                continue
            T, p, q, ranks, B, Dm, L = self._dims(it)
            assert L == list(O.make_L(p))
            assert it.plan_ready == 0
            if it.workspace_bytes:  # header contract: zero on entry
                assert it.workspace % 256 == 0
                hb = self.real.ttb_tt_workspace_header_bytes(ctypes.byref(it.shape), it.nnz)
                assert not _np(it.workspace, hb, ctypes.c_uint8, np.uint8).any()
            idx = _np(it.indices, it.nnz, ctypes.c_int64, np.int64)
            row = _np(it.rowidx, it.nnz, ctypes.c_int64, np.int64)
            tbl = _np(it.tableidx, it.nnz, ctypes.c_int64, np.int64)
            out = _np(it.output, B * Dm, ctypes.c_float, np.float32).reshape(1, B, Dm)
            assert not out.any(), "output must arrive zero-filled"
            out += O.tt_forward(1, B, Dm, p, q, ranks, L, it.nnz, idx, row, tbl, self._cores(it, "cores"))
        return 0

    def ttb_group_backward(self, n, items, optim, lr, eps, stream):
        self.calls.append(("backward", optim))
        for i in range(n):
            it = items[i]
            if it.nnz == 0:
                continue
            T, p, q, ranks, B, Dm, L = self._dims(it)
            assert it.plan_ready == (1 if it.workspace_bytes else 0)
            idx = _np(it.indices, it.nnz, ctypes.c_int64, np.int64)
            row = _np(it.rowidx, it.nnz, ctypes.c_int64, np.int64)
            tbl = _np(it.tableidx, it.nnz, ctypes.c_int64, np.int64)
            d_out = _np(it.d_output, B * Dm, ctypes.c_float, np.float32).reshape(1, B, Dm)
            cores = self._cores(it, "cores")
            grads = self._cores(it, "grads")
            for g in grads:
                assert not g.any(), "gradient buffers must arrive zero"
            g = O.tt_backward_dense(Dm, p, q, ranks, L, it.nnz, idx, row, tbl, d_out, cores)
            if optim == 2:  # TTB_OPTIM_DENSE
                for dst, src in zip(grads, g):
                    dst[...] = src
            elif optim == 0:
                for c, new in zip(cores, O.sgd_step(cores, g, lr)):
                    c[...] = new
            else:
                state = self._cores(it, "opt_state")
                new_c, new_s = O.adagrad_step(cores, state, g, lr, eps)
                for c, s_, nc, ns in zip(cores, state, new_c, new_s):
                    c[...] = nc
                    s_[...] = ns
        return 0


@pytest.fixture
def cpu_ext(monkeypatch):
    from fbtt_embedding_b200 import tt_embeddings as ext

    fake = _FakeLib(ext._lib)
    monkeypatch.setattr(ext, "_lib", fake)
    monkeypatch.setattr(ext, "_i64c", lambda t, what: t.contiguous())
    monkeypatch.setattr(ext, "_f32c", lambda t, what: t.contiguous())
    monkeypatch.setattr(ext, "_cores_inplace", lambda cores, what="tt_cores": [c.data for c in cores])
    monkeypatch.setattr(ext, "_DeviceGuard", lambda t: contextlib.nullcontext())
    monkeypatch.setattr(ext, "_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    return ext, fake


def _group(optimizer_name, sparse, seed=0, lr=0.1, eps=1e-3):
    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200.grouped import GroupedLookup

    rng = np.random.RandomState(seed)
    tables = [_FakeTable(s, getattr(OptimType, optimizer_name), sparse, lr, eps, rng) for s in SPECS]
    return GroupedLookup(tables), tables


def _inputs(rng, B, empty_table=None):
    idx, off = [], []
    for n, s in enumerate(SPECS):
        e = n == empty_table
        i, o = ragged_batch(rng, B, int(np.prod(s["p"])), 0.0 if e else 4.0, 0.0 if e else 2.0)
        idx.append(torch.from_numpy(i))
        off.append(torch.from_numpy(o))
    return idx, off


def _oracle_table(spec, cores, idx, off, B):
    row, tbl = O.compute_rowidx(off.numpy(), 1)
    i = idx.numpy()
    return O.tt_forward(1, B, D, spec["p"], spec["q"], spec["ranks"], O.make_L(spec["p"]), len(i), i, row, tbl, cores)[0], (i, row, tbl)


@pytest.mark.parametrize("optimizer", ["SGD", "EXACT_ADAGRAD"])
def test_group_item_array_drives_a_fused_step(cpu_ext, optimizer):
    ext, fake = cpu_ext
    group, tables = _group(optimizer, sparse=True)
    rng = np.random.RandomState(5)
    B = 24
    for step in range(3):  # several steps: pooled plan buffers and the zero-on-exit scratch are reused
        idx, off = _inputs(rng, B, empty_table=step)
        before = [[c.detach().numpy().copy() for c in t.tt_cores] for t in tables]
        state0 = [[s.numpy().copy() for s in t.optimizer_state] for t in tables]
        d_out = torch.from_numpy(rng.uniform(-1, 1, (len(SPECS), B, D)).astype(np.float32))
        out = group.lookup(idx, off)
        assert out.shape == (len(SPECS), B, D) and out.requires_grad
        out.backward(d_out)
        for n, spec in enumerate(SPECS):
            want, (i, row, tbl) = _oracle_table(spec, before[n], idx[n], off[n], B)
            np.testing.assert_allclose(out[n].detach().numpy(), want, rtol=1e-5, atol=1e-6)
            g = O.tt_backward_dense(D, spec["p"], spec["q"], spec["ranks"], O.make_L(spec["p"]), len(i), i, row, tbl,
                                    d_out[n].numpy()[None], before[n])
            if optimizer == "SGD":
                new_c = O.sgd_step(before[n], g, 0.1)
            else:
                new_c, new_s = O.adagrad_step(before[n], state0[n], g, 0.1, 1e-3)
                for a, b in zip(tables[n].optimizer_state, new_s):
                    np.testing.assert_allclose(a.numpy(), b, rtol=1e-6, atol=1e-7)
            for a, b in zip(tables[n].tt_cores, new_c):
                np.testing.assert_allclose(a.detach().numpy(), b, rtol=1e-6, atol=1e-7)
    assert fake.calls[:3] == ["preprocess", "forward", ("backward", 0 if optimizer == "SGD" else 1)]
    assert len(fake.calls) == 9  # three host calls per step for the whole group


def test_group_dense_mode_returns_core_gradients(cpu_ext):
    ext, fake = cpu_ext
    group, tables = _group("SGD", sparse=False, seed=1)
    rng = np.random.RandomState(6)
    B = 16
    idx, off = _inputs(rng, B, empty_table=2)
    d_out = torch.from_numpy(rng.uniform(-1, 1, (len(SPECS), B, D)).astype(np.float32))
    group.lookup(idx, off).backward(d_out)
    assert fake.calls[-1] == ("backward", 2)
    for n, spec in enumerate(SPECS):
        cores = [c.detach().numpy() for c in tables[n].tt_cores]
        _, (i, row, tbl) = _oracle_table(spec, cores, idx[n], off[n], B)
        g = O.tt_backward_dense(D, spec["p"], spec["q"], spec["ranks"], O.make_L(spec["p"]), len(i), i, row, tbl,
                                d_out[n].numpy()[None], cores)
        for c, want in zip(tables[n].tt_cores, g):
            assert c.grad is not None and c.grad.shape == c.shape
            np.testing.assert_allclose(c.grad.numpy(), want, rtol=1e-5, atol=1e-6)


def test_group_plan_buffers_are_pooled_and_inference_returns_them(cpu_ext):
    ext, fake = cpu_ext
    group, tables = _group("SGD", sparse=True, seed=2)
    rng = np.random.RandomState(7)
    B = 8
    idx, off = _inputs(rng, B)
    group.lookup(idx, off).backward(torch.zeros(len(SPECS), B, D))
    pooled = [b for free in group._plan_free.values() for b in free]
    assert len(pooled) == 1, "the step's backward parks the plan buffer (two items of SPECS take a bucketed path)"
    with torch.no_grad():
        group.lookup(idx, off)
    again = [b for free in group._plan_free.values() for b in free]
    assert len(again) == 1 and again[0].data_ptr() == pooled[0].data_ptr(), \
        "an inference forward draws the pooled buffer and hands it straight back"
    out1 = group.lookup(idx, off)  # two forwards in flight: the second one must not share the first one's plan
    out2 = group.lookup(idx, off)
    assert not [b for free in group._plan_free.values() for b in free]
    out1.backward(torch.zeros(len(SPECS), B, D))
    out2.backward(torch.zeros(len(SPECS), B, D))
    assert len([b for free in group._plan_free.values() for b in free]) == 2


def test_group_input_validation(cpu_ext):
    ext, fake = cpu_ext
    group, tables = _group("SGD", sparse=True, seed=3)
    rng = np.random.RandomState(8)
    idx, off = _inputs(rng, 8)
    with pytest.raises(RuntimeError):
        group.lookup(idx[:-1], off[:-1])
    bad = list(off)
    bad[1] = bad[1][:-1]
    with pytest.raises(RuntimeError):
        group.lookup(idx, bad)
    out = group.lookup(idx, off)
    with pytest.raises(RuntimeError):
        out.backward(torch.zeros(len(SPECS), 8, D + 4))  # autograd itself rejects the wrong d_output shape
