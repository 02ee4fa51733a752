"""CPU tests of the fused heterogeneous batch (include/ttb.h: ttb_het_*; fbtt_embedding_b200/fused.py).

1. The index decomposition the het kernels perform is one inline function compiled for host and device;
   ``ttb_het_digits`` runs it on the host, so it is pinned here against the oracle's per-table digits.
2. The module's host logic -- descriptor upload, concatenated parameters, table-major packing, autograd dispatch
   on (sparse, optimizer), plan-buffer handling -- is executed against a stand-in for the three device entry
   points that reads the SAME descriptors and concatenated cores through the pointers it is handed and computes
   every table with the numpy oracle.  What the real kernels do with those arguments is covered on the GPU
   (tests/test_zz4_gpu_fused.py)."""
import contextlib
import ctypes

import numpy as np
import pytest
import torch

from oracle import tt_oracle as O
from tests.helpers import ragged_batch

D, Q, RANKS = 16, [2, 2, 4], [3, 5]
P_SHAPES = [[5, 6, 7], [1, 2, 3], [20, 22, 25], [1, 1, 3], [4, 9, 2]]
E = [200, 5, 10900, 3, 72]  # num_embeddings <= prod(p): the tail rows of a table exist but are never looked up


def _np(ptr, n, ctype, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(n,)).view(dtype)


def test_het_describe_and_digits_match_the_oracle():
    from fbtt_embedding_b200 import tt_embeddings as ext

    lay = ext.HetLayout(P_SHAPES)
    assert lay.P == [sum(p[t] for p in P_SHAPES) for t in range(3)]
    rng = np.random.RandomState(0)
    for k, p in enumerate(P_SHAPES):
        assert lay.rows[k] == int(np.prod(p))
        assert list(lay.host[k].L)[:3] == list(O.make_L(p))
        assert lay.off[k] == [sum(q[t] for q in P_SHAPES[:k]) for t in range(3)]
        rows = lay.rows[k]
        probe = np.unique(np.concatenate([[0, rows - 1], rng.randint(0, rows, 50)]))
        L = O.make_L(p)
        for idx in probe:
            rem, want = int(idx), []
            for t in range(3):  # tt_embeddings_cuda.cu:795-799
                want.append(lay.off[k][t] + rem // int(L[t]))
                rem %= int(L[t])
            assert lay.digits(k, int(idx)) == want
        assert lay.digits(k, rows) is None and lay.digits(k, -1) is None
    assert lay.digits(len(P_SHAPES), 0) is None and lay.digits(-1, 0) is None


def test_het_digits_64bit_tables():
    from fbtt_embedding_b200 import tt_embeddings as ext

    p = [[2000, 2000, 2000], [3, 3, 3]]  # 8e9 rows: the 64-bit division branch
    lay = ext.HetLayout(p)
    for idx in (0, 7_999_999_999, 4_000_000_001, (1 << 32) + 5, (1 << 31) - 1):
        assert lay.digits(0, idx) == [idx // 4_000_000, (idx // 2000) % 2000, idx % 2000]
    assert lay.digits(0, 8_000_000_000) is None
    assert lay.digits(1, 26) == [2002, 2002, 2002]


def test_het_describe_rejects_bad_shapes():
    from fbtt_embedding_b200 import tt_embeddings as ext

    with pytest.raises(RuntimeError):
        ext.HetLayout([[3, 0, 5]])
    with pytest.raises(RuntimeError):
        ext.HetLayout([[3, 4, 5], [3, 4]])
    with pytest.raises(RuntimeError):
        ext.HetLayout([])
    with pytest.raises(RuntimeError):
        ext.HetLayout([[1 << 30, 1 << 30, 1 << 30]])  # prod(p) overflows int64... 2^90


class _FakeLib:
    """ttb_preprocess_rowidx / ttb_tt_forward_het / ttb_tt_backward_het over host memory with the oracle."""

    def __init__(self, real, ext):
        self.real, self.ext = real, ext
        self.calls = []

    def __getattr__(self, name):
        return getattr(self.real, name)

    def ttb_preprocess_rowidx(self, nnz, num_bags, B, offsets, rowidx, tableidx, stream):
        self.calls.append("rowidx")
        off = _np(offsets, num_bags + 1, ctypes.c_int64, np.int64)
        row, tbl = O.compute_rowidx(off, num_bags // B)
        _np(rowidx, nnz, ctypes.c_int64, np.int64)[:] = row
        _np(tableidx, nnz, ctypes.c_int64, np.int64)[:] = tbl
        return 0

    def _decode(self, shape_ref, n_tables, tables_ptr, cores_arr):
        s = shape_ref._obj
        T = s.T
        assert s.num_tables == 1
        q, R = list(s.q)[:T], list(s.R)[:T + 1]
        tabs = (self.ext._HetTable * n_tables).from_address(tables_ptr)
        P = list(s.p)[:T]
        assert list(s.L)[:T] == list(O.make_L(P))
        cat = [_np(cores_arr[t], P[t] * R[t] * q[t] * R[t + 1], ctypes.c_float, np.float32).reshape(1, P[t], -1)
               for t in range(T)]
        per_table = []
        for k in range(n_tables):
            h = tabs[k]
            p, off = list(h.p)[:T], list(h.off)[:T]
            assert list(h.L)[:T] == list(O.make_L(p)) and h.rows == int(np.prod(p))
            per_table.append((p, [c[:, off[t]:off[t] + p[t]] for t, c in enumerate(cat)]))
        return s, q, R[1:T], per_table, cat

    @staticmethod
    def _row_ptr(base, row_map_ref, D, k, b):
        """address of (table k, bag b)'s pooled row: include/ttb.h's row-map formula over the HOST arrays the
        (CPU) "device" pointers name"""
        m = row_map_ref._obj
        off = _np(m.peer_offset, m.world, ctypes.c_int64, np.int64)
        gid = _np(m.table_gid, max(k + 1, 1), ctypes.c_int32, np.int32)
        w, r = divmod(int(b), m.rows_per_rank)
        return base + 4 * (int(off[w]) + (r * m.tables_total + int(gid[k])) * D)

    def ttb_tt_forward_het(self, shape_ref, n_tables, tables, row_map, nnz, indices, rowidx, tableidx, cores, out, ws,
                           wsb, plan_ready, stream):
        self.calls.append("forward")
        s, q, ranks, per_table, _ = self._decode(shape_ref, n_tables, tables, cores)
        assert plan_ready == 0
        idx = _np(indices, nnz, ctypes.c_int64, np.int64)
        row = _np(rowidx, nnz, ctypes.c_int64, np.int64)
        tbl = _np(tableidx, nnz, ctypes.c_int64, np.int64)
        if row_map is None:
            o = _np(out, n_tables * s.B * s.D, ctypes.c_float, np.float32).reshape(n_tables, s.B, s.D)
            assert not o.any(), "output must arrive zero-filled"
        for k, (p, cores_k) in enumerate(per_table):
            m = tbl == k
            if m.any():
                n = int(m.sum())
                pooled = O.tt_forward(1, s.B, s.D, p, q, ranks, O.make_L(p), n, idx[m], row[m], np.zeros(n, np.int64),
                                      cores_k)[0]
                if row_map is None:
                    o[k] += pooled
                else:  # fused exchange: ADD every bag's row into the buffer of the rank that owns its batch slice
                    for b in np.unique(row[m]):
                        _np(self._row_ptr(out, row_map, s.D, k, b), s.D, ctypes.c_float, np.float32)[:] += pooled[b]
        return 0

    def ttb_tt_backward_het(self, shape_ref, n_tables, tables, row_map, optim, lr, eps, nnz, indices, rowidx, tableidx,
                            d_output, cores, grads, opt_state, ws, wsb, plan_ready, stream):
        self.calls.append(("backward", optim, plan_ready))
        s, q, ranks, per_table, cat = self._decode(shape_ref, n_tables, tables, cores)
        _, _, _, per_table_g, cat_g = self._decode(shape_ref, n_tables, tables, grads)
        for g in cat_g:
            assert not g.any(), "gradient buffers must arrive zero"
        idx = _np(indices, nnz, ctypes.c_int64, np.int64)
        row = _np(rowidx, nnz, ctypes.c_int64, np.int64)
        tbl = _np(tableidx, nnz, ctypes.c_int64, np.int64)
        if row_map is None:
            d_out = _np(d_output, n_tables * s.B * s.D, ctypes.c_float, np.float32).reshape(n_tables, s.B, s.D)
        else:  # gather every (table, bag) gradient row from the rank that owns the bag's batch slice
            d_out = np.zeros((n_tables, s.B, s.D), np.float32)
            for k in range(n_tables):
                for b in range(s.B):
                    d_out[k, b] = _np(self._row_ptr(d_output, row_map, s.D, k, b), s.D, ctypes.c_float, np.float32)
        for k, (p, cores_k) in enumerate(per_table):
            m = tbl == k
            if not m.any():
                continue
            n = int(m.sum())
            g = O.tt_backward_dense(s.D, p, q, ranks, O.make_L(p), n, idx[m], row[m], np.zeros(n, np.int64),
                                    d_out[k][None], [c.copy() for c in cores_k])
            for dst, src in zip(per_table_g[k][1], g):
                dst[...] = src
        if optim == 2:
            return 0
        if optim == 0:
            for c, new in zip(cat, O.sgd_step([c.copy() for c in cat], cat_g, lr)):
                c[...] = new
        else:
            _, _, _, _, cat_s = self._decode(shape_ref, n_tables, tables, opt_state)
            new_c, new_s = O.adagrad_step([c.copy() for c in cat], [x.copy() for x in cat_s], cat_g, lr, eps)
            for c, s_, nc, ns in zip(cat, cat_s, new_c, new_s):
                c[...] = nc
                s_[...] = ns
        for g in cat_g:  # fused modes leave the scratch zero (include/ttb.h)
            g[...] = 0
        return 0


@pytest.fixture
def cpu_ext(monkeypatch):
    from fbtt_embedding_b200 import tt_embeddings as ext

    fake = _FakeLib(ext._lib, ext)
    monkeypatch.setattr(ext, "_lib", fake)
    monkeypatch.setattr(ext, "_i64c", lambda t, what: t.contiguous())
    monkeypatch.setattr(ext, "_f32c", lambda t, what: t.contiguous())
    monkeypatch.setattr(ext, "_cores_inplace", lambda cores, what="tt_cores": [c.data for c in cores])
    monkeypatch.setattr(ext, "_DeviceGuard", lambda t: contextlib.nullcontext())
    monkeypatch.setattr(ext, "_stream", lambda: 0)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    ext._core_cache.clear()
    ext._grad_cache.clear()
    yield ext, fake
    ext._core_cache.clear()
    ext._grad_cache.clear()
    ext._plan_cache.clear()
    ext._plan_free.clear()


def _module(optimizer_name, sparse, lr=0.1, eps=1e-3):
    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200.fused import FusedTTEmbeddingBag

    torch.manual_seed(0)
    return FusedTTEmbeddingBag(E, D, RANKS, P_SHAPES, Q, optimizer=getattr(OptimType, optimizer_name),
                               learning_rate=lr, eps=eps, sparse=sparse, weight_dist="uniform", device="cpu")


def _inputs(rng, B, empty_table=None):
    idx, off = [], []
    for k in range(len(E)):
        e = k == empty_table
        i, o = ragged_batch(rng, B, E[k], 0.0 if e else 3.0, 0.0 if e else 2.0)
        idx.append(torch.from_numpy(i))
        off.append(torch.from_numpy(o))
    return idx, off


def _oracle_step(mod, before, idx, off, B, d_out):
    """per table: pooled rows and dense core gradients from that table's own slices"""
    outs, grads = [], []
    for k, p in enumerate(P_SHAPES):
        o_ = mod.layout.off[k]
        cores = [before[t][:, o_[t]:o_[t] + p[t]] for t in range(3)]
        row, tbl = O.compute_rowidx(off[k].numpy(), 1)
        i = idx[k].numpy()
        outs.append(O.tt_forward(1, B, D, p, Q, RANKS, O.make_L(p), len(i), i, row, tbl, cores)[0])
        grads.append(O.tt_backward_dense(D, p, Q, RANKS, O.make_L(p), len(i), i, row, tbl, d_out[k].numpy()[None], cores))
    cat_g = [np.concatenate([grads[k][t] for k in range(len(E))], axis=1) for t in range(3)]
    return np.stack(outs), cat_g


def test_pack_table_major():
    from fbtt_embedding_b200.fused import pack_table_major

    rng = np.random.RandomState(3)
    idx, off = _inputs(rng, 6, empty_table=1)
    ci, co = pack_table_major(idx, off)
    assert ci.numel() == sum(i.numel() for i in idx) and co.numel() == len(E) * 6 + 1
    row, tbl = O.compute_rowidx(co.numpy(), len(E))
    pos = 0
    for k in range(len(E)):
        n = idx[k].numel()
        r1, _ = O.compute_rowidx(off[k].numpy(), 1)
        assert (tbl[pos:pos + n] == k).all() and (row[pos:pos + n] == r1).all()
        assert (ci[pos:pos + n] == idx[k]).all()
        pos += n
    with pytest.raises(RuntimeError):
        pack_table_major(idx, off[:-1])


@pytest.mark.parametrize("optimizer", ["SGD", "EXACT_ADAGRAD"])
def test_fused_module_step_matches_per_table_oracle(cpu_ext, optimizer):
    ext, fake = cpu_ext
    mod = _module(optimizer, sparse=True)
    assert [tuple(c.shape) for c in mod.tt_cores] == [(1, 31, 6), (1, 40, 30), (1, 40, 20)]
    rng = np.random.RandomState(5)
    B = 12
    for step in range(3):
        idx, off = _inputs(rng, B, empty_table=step)
        before = [c.detach().numpy().copy() for c in mod.tt_cores]
        state0 = [s.numpy().copy() for s in mod.optimizer_state]
        d_out = torch.from_numpy(rng.uniform(-1, 1, (len(E), B, D)).astype(np.float32))
        out = mod(idx, off)
        assert out.shape == (len(E), B, D) and out.requires_grad
        out.backward(d_out)
        want_out, cat_g = _oracle_step(mod, before, idx, off, B, d_out)
        np.testing.assert_allclose(out.detach().numpy(), want_out, rtol=1e-5, atol=1e-6)
        if optimizer == "SGD":
            new_c = O.sgd_step(before, cat_g, 0.1)
        else:
            new_c, new_s = O.adagrad_step(before, state0, cat_g, 0.1, 1e-3)
            for a, b in zip(mod.optimizer_state, new_s):
                np.testing.assert_allclose(a.numpy(), b, rtol=1e-6, atol=1e-7)
        for a, b in zip(mod.tt_cores, new_c):
            np.testing.assert_allclose(a.detach().numpy(), b, rtol=1e-6, atol=1e-7)
    kinds = [c if isinstance(c, str) else c[0] for c in fake.calls]
    assert kinds == ["rowidx", "forward", "backward"] * 3  # three device entry points per step for ALL tables
    assert all(c[1] == (0 if optimizer == "SGD" else 1) for c in fake.calls if not isinstance(c, str))


def test_fused_module_dense_mode_and_table_major_tensor_input(cpu_ext):
    from fbtt_embedding_b200.fused import pack_table_major

    ext, fake = cpu_ext
    mod = _module("SGD", sparse=False)
    rng = np.random.RandomState(6)
    B = 10
    idx, off = _inputs(rng, B, empty_table=3)
    ci, co = pack_table_major(idx, off)
    d_out = torch.from_numpy(rng.uniform(-1, 1, (len(E), B, D)).astype(np.float32))
    before = [c.detach().numpy().copy() for c in mod.tt_cores]
    out = mod(ci, co)  # the TableBatchedTTEmbeddingBag calling convention
    out.backward(d_out)
    assert fake.calls[-1][:2] == ("backward", 2)
    want_out, cat_g = _oracle_step(mod, before, idx, off, B, d_out)
    np.testing.assert_allclose(out.detach().numpy(), want_out, rtol=1e-5, atol=1e-6)
    for c, want in zip(mod.tt_cores, cat_g):
        assert c.grad is not None and c.grad.shape == c.shape
        np.testing.assert_allclose(c.grad.numpy(), want, rtol=1e-5, atol=1e-6)
    for a, b in zip(mod.tt_cores, before):  # dense mode leaves the update to the caller's optimizer
        assert (a.detach().numpy() == b).all()


def test_fused_module_table_views_and_init_scale(cpu_ext):
    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200.fused import FusedTTEmbeddingBag

    mod = _module("SGD", sparse=True)
    for k, p in enumerate(P_SHAPES):
        views = mod.table_cores(k)
        assert [v.shape[1] for v in views] == p
        # "uniform" draws from [0, hi(E_k)): every table is initialised with its OWN num_embeddings
        sigma = np.sqrt(2.0 / (E[k] + D))
        hi = sigma ** (1.0 / 3) * float(np.prod(np.asarray([1] + RANKS + [1], np.float64) ** (-1.0 / 6)))
        for v in views:
            assert float(v.min()) >= 0.0 and float(v.max()) <= hi * (1 + 1e-6)
        assert max(float(v.max()) for v in views) > 0.5 * hi
    src = [torch.full((1, P_SHAPES[2][t], mod.tt_cores[t].shape[2]), float(t + 1)) for t in range(3)]
    mod.load_table(2, src)
    for t, v in enumerate(mod.table_cores(2)):
        assert (v == float(t + 1)).all()
    assert not (mod.table_cores(1)[0] == 1.0).any()
    with pytest.raises(AssertionError):
        FusedTTEmbeddingBag([100], D, RANKS, [[2, 2, 2]], Q, device="cpu")  # prod(p) < num_embeddings
    assert OptimType.SGD in (mod.optimizer,)


def test_fused_module_input_validation(cpu_ext):
    ext, fake = cpu_ext
    mod = _module("SGD", sparse=True)
    rng = np.random.RandomState(8)
    idx, off = _inputs(rng, 8)
    with pytest.raises(RuntimeError):
        mod(idx[:-1], off)
    with pytest.raises(RuntimeError):
        mod(torch.cat(idx), torch.arange(0, 8))  # 7 bags: not a multiple of the table count
    out = mod(idx, off)
    with pytest.raises(RuntimeError):
        ext.tt_backward_het(mod.layout, ext.OPTIM_SGD, D, 0.1, 0.0, Q, mod.tt_ranks, 1, idx[0][:1], idx[0][:1],
                            idx[0][:1], torch.zeros(len(E) + 1, 8, D), list(mod.tt_cores))
    del out


# ---- fused exchange (ttb_row_map_t; fbtt_embedding_b200/sharded.py exchange="peer") ----------------------------
def test_row_map_offset_is_the_documented_formula():
    from fbtt_embedding_b200 import tt_embeddings as ext

    world, bw, tt, Dm = 3, 4, 7, 16
    off = [0, 123456, -98764]  # peers may sit below the local buffer
    gid = [5, 0, 6]
    m = ext.RowMap(world, bw, tt, off, gid, "cpu")
    for k in range(3):
        for b in range(world * bw):
            w, r = divmod(b, bw)
            assert m.offset(Dm, k, b) == off[w] + (r * tt + gid[k]) * Dm
    with pytest.raises(RuntimeError):
        m.offset(Dm, 0, world * bw)  # row outside the batch
    with pytest.raises(RuntimeError):
        ext.RowMap(world, bw, tt, off[:2], gid, "cpu")
    with pytest.raises(RuntimeError):
        ext.RowMap(world, bw, tt, off, [7], "cpu")  # table_gid >= tables_total


def _sharded(world, rank, optimizer_name="SGD", sparse=True):
    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200.sharded import TableShardedTTEmbeddingBag

    specs = [dict(num_embeddings=E[k], embedding_dim=D, tt_ranks=RANKS, tt_p_shapes=P_SHAPES[k], tt_q_shapes=Q)
             for k in range(len(E))]
    torch.manual_seed(100 + rank)
    return TableShardedTTEmbeddingBag(specs, None, fused=True, exchange="peer", world_size=world, rank=rank,
                                      optimizer=getattr(OptimType, optimizer_name), learning_rate=0.1, eps=1e-3,
                                      sparse=sparse, weight_dist="uniform", device="cpu")


@pytest.mark.parametrize("optimizer", ["SGD", "EXACT_ADAGRAD"])
def test_peer_exchange_three_ranks_in_one_process(cpu_ext, optimizer):
    """Every rank scatters the pooled rows of ITS tables into the batch-slice buffers of ALL ranks and gathers its
    gradients from there: after the scatter rank w holds out[w*bw:(w+1)*bw] of all tables, and each rank's fused
    update equals the oracle's step of its tables on the whole batch."""
    from fbtt_embedding_b200.sharded import LocalPeers

    ext, fake = cpu_ext
    W, B = 3, 12
    bw, T = B // W, len(E)
    mods = [_sharded(W, r, optimizer) for r in range(W)]
    assert sorted(t for m in mods for t in m.local_tables) == list(range(T))
    peers = LocalPeers(W, bw, T, D, "cpu")
    views = [peers.view(r) for r in range(W)]
    for m, v in zip(mods, views):
        m._peer_setup(v, B)
    rng = np.random.RandomState(9)
    idx, off = _inputs(rng, B, empty_table=4)
    before = [[c.detach().numpy().copy() for c in m.fused.tt_cores] for m in mods]
    state0 = [[s.numpy().copy() for s in m.fused.optimizer_state] for m in mods]
    states = [m._phase_forward(v, [idx[t] for t in m.local_tables], [off[t] for t in m.local_tables])
              for m, v in zip(mods, views)]
    # what every rank must now hold: its batch slice of every table
    pooled = np.zeros((T, B, D), np.float32)
    for r, m in enumerate(mods):
        for k, t in enumerate(m.local_tables):
            o_ = m.fused.layout.off[k]
            cores = [before[r][c][:, o_[c]:o_[c] + P_SHAPES[t][c]] for c in range(3)]
            row, tbl = O.compute_rowidx(off[t].numpy(), 1)
            i = idx[t].numpy()
            pooled[t] = O.tt_forward(1, B, D, P_SHAPES[t], Q, RANKS, O.make_L(P_SHAPES[t]), len(i), i, row, tbl, cores)[0]
    for w in range(W):
        want = pooled[:, w * bw:(w + 1) * bw].transpose(1, 0, 2)  # [bw, T, D]
        np.testing.assert_allclose(views[w].x.numpy(), want, rtol=1e-5, atol=1e-6)
    # backward: every rank's upstream gradient for its batch slice, all tables
    d_x = [rng.uniform(-1, 1, (bw, T, D)).astype(np.float32) for _ in range(W)]
    for w in range(W):
        views[w].dx.copy_(torch.from_numpy(d_x[w]))
    d_pooled = np.concatenate([d.transpose(1, 0, 2) for d in d_x], axis=1)  # [T, B, D]
    for m, v, st in zip(mods, views, states):
        assert m._phase_backward(v, st) is None
    for r, m in enumerate(mods):
        grads = []
        for k, t in enumerate(m.local_tables):
            o_ = m.fused.layout.off[k]
            cores = [before[r][c][:, o_[c]:o_[c] + P_SHAPES[t][c]] for c in range(3)]
            row, tbl = O.compute_rowidx(off[t].numpy(), 1)
            i = idx[t].numpy()
            grads.append(O.tt_backward_dense(D, P_SHAPES[t], Q, RANKS, O.make_L(P_SHAPES[t]), len(i), i, row, tbl,
                                             d_pooled[t][None], cores))
        cat_g = [np.concatenate([g[c] for g in grads], axis=1) for c in range(3)]
        if optimizer == "SGD":
            new_c = O.sgd_step(before[r], cat_g, 0.1)
        else:
            new_c, new_s = O.adagrad_step(before[r], state0[r], cat_g, 0.1, 1e-3)
            for a, b in zip(m.fused.optimizer_state, new_s):
                np.testing.assert_allclose(a.numpy(), b, rtol=1e-6, atol=1e-7)
        for a, b in zip(m.fused.tt_cores, new_c):
            np.testing.assert_allclose(a.detach().numpy(), b, rtol=1e-6, atol=1e-7)


def test_peer_exchange_autograd_single_rank_equals_plain_fused_module(cpu_ext):
    """World size 1 through the autograd node (zero -> barrier -> scatter -> barrier -> clone; copy -> barrier ->
    gather): the result is the plain fused module's, transposed to [B, T, D]; dense mode returns core gradients."""
    from fbtt_embedding_b200.sharded import LocalPeers, _PeerLookup

    ext, fake = cpu_ext
    B = 8
    mod = _sharded(1, 0, sparse=False)
    barriers = []
    view = LocalPeers(1, B, len(E), D, "cpu").view(0)
    view.barrier = lambda: barriers.append(len(fake.calls))
    mod._peer_setup(view, B)
    rng = np.random.RandomState(10)
    idx, off = _inputs(rng, B)
    plain = mod.fused(idx, off)  # [T, B, D] through the ordinary entry point
    d_out = torch.from_numpy(rng.uniform(-1, 1, (B, len(E), D)).astype(np.float32))
    plain.backward(d_out.permute(1, 0, 2).contiguous())
    want_grads = [c.grad.clone() for c in mod.fused.tt_cores]
    for c in mod.fused.tt_cores:
        c.grad = None
    out = _PeerLookup.apply(mod, view, tuple(idx), tuple(off), *mod.fused.tt_cores)
    assert out.shape == (B, len(E), D) and out.requires_grad
    np.testing.assert_allclose(out.detach().numpy(), plain.detach().permute(1, 0, 2).numpy(), rtol=1e-6, atol=1e-7)
    assert out.data_ptr() != view.x.data_ptr(), "the exchange buffer is recycled: callers get a copy"
    out.backward(d_out)
    for c, g in zip(mod.fused.tt_cores, want_grads):
        np.testing.assert_allclose(c.grad.numpy(), g.numpy(), rtol=1e-6, atol=1e-7)
    assert len(barriers) == 2, "rows landed -> read, gradients in place -> gather"
    # later steps alternate between the two X regions and always start from a zeroed one (rows are ADDED)
    used = {view.x.data_ptr()}
    for _ in range(3):
        out2 = _PeerLookup.apply(mod, view, tuple(idx), tuple(off), *mod.fused.tt_cores)
        used.add(view.x.data_ptr())
        np.testing.assert_allclose(out2.detach().numpy(), out.detach().numpy(), rtol=1e-6, atol=1e-7)
    assert len(used) == 2
    # the table-major (keyed-jagged) pair goes through unchanged -- no per-step packing
    from fbtt_embedding_b200.fused import pack_table_major

    ci, co = pack_table_major(idx, off)
    out3 = _PeerLookup.apply(mod, view, ci, co, *mod.fused.tt_cores)
    np.testing.assert_allclose(out3.detach().numpy(), out.detach().numpy(), rtol=1e-6, atol=1e-7)


def test_fused_module_exchanges_tables_with_single_table_checkpoints(cpu_ext):
    """table_state_dict / load_table_state_dict speak the reference's per-module key names."""
    a = _module("EXACT_ADAGRAD", sparse=True)
    b = _module("EXACT_ADAGRAD", sparse=True)
    with torch.no_grad():
        for c in b.tt_cores:
            c.zero_()
        for s_ in a.optimizer_state:
            s_.uniform_(0.0, 1.0)
    for k in range(len(E)):
        sd = a.table_state_dict(k)
        assert set(sd) == {"tt_cores.0", "tt_cores.1", "tt_cores.2", "optimizer_state.optimizer_state0",
                           "optimizer_state.optimizer_state1", "optimizer_state.optimizer_state2", "L", "hashtbl",
                           "cache_state"}
        assert sd["L"].tolist() == list(O.make_L(P_SHAPES[k]))
        assert [tuple(sd[f"tt_cores.{t}"].shape) for t in range(3)] == [
            (1, P_SHAPES[k][t], a.tt_cores[t].shape[2]) for t in range(3)]
        b.load_table_state_dict(k, sd)
    for x, y in zip(a.tt_cores, b.tt_cores):
        assert torch.equal(x, y)
    for x, y in zip(a.optimizer_state, b.optimizer_state):
        assert torch.equal(x, y)
    bad = a.table_state_dict(0)
    with pytest.raises(RuntimeError):
        b.load_table_state_dict(2, bad)  # another table's p-shape
    sgd = _module("SGD", sparse=True)  # no Adagrad state on this side: cores only
    sgd.load_table_state_dict(1, a.table_state_dict(1))
    for x, y in zip(sgd.table_cores(1), a.table_cores(1)):
        assert torch.equal(x, y)
