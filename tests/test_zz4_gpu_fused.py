"""GPU tests of the fused heterogeneous batch (include/ttb.h: ttb_tt_forward_het / ttb_tt_backward_het,
fbtt_embedding_b200/fused.py): tables of different sizes that share q-shapes and ranks, cores concatenated along
the slice dimension, ONE plan / forward / backward / sweep launch for all of them.  Must equal, table by table, the
oracle and the single-table modules holding the same weights -- on the exact FFMA path and on both tensor-core
families.  (Sorts last on purpose: newest path, written after the round's GPU budget was spent.)"""
import numpy as np
import pytest
import torch

from oracle import tt_oracle as O
from tests.helpers import ragged_batch, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (q, ranks, D) per kernel family; p-shapes include tables with a single slice in a core and a 1-row table
FAMILIES = {
    "tcgen05": dict(q=[4, 4, 4], ranks=[32, 32]),
    "tcgen05_q2_8_r1_16": dict(q=[4, 4, 8], ranks=[16, 32]),
    "warp_mma": dict(q=[4, 4, 8], ranks=[64, 64]),
    "warp_mma16": dict(q=[4, 2, 4], ranks=[16, 16]),
    "generic": dict(q=[2, 4, 4], ranks=[5, 7]),
}
P3 = [[20, 22, 25], [1, 1, 3], [7, 9, 11], [2, 30, 2], [1, 1, 1], [6, 5, 4]]
P2 = [[30, 40], [1, 5], [9, 9]]
P4 = [[3, 4, 5, 6], [1, 2, 1, 2], [2, 2, 2, 2]]


@pytest.fixture(scope="module")
def ext():
    from fbtt_embedding_b200 import tt_embeddings as e

    yield e
    e.set_path(e.PATH_AUTO)


def t(x):
    return torch.as_tensor(np.ascontiguousarray(x), device=DEV)


def _module(p_shapes, q, ranks, optimizer="SGD", sparse=True, lr=0.1, eps=1e-4):
    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200.fused import FusedTTEmbeddingBag

    torch.manual_seed(7)
    E = [int(np.prod(p)) for p in p_shapes]
    return FusedTTEmbeddingBag(E, int(np.prod(q)), ranks, p_shapes, q, optimizer=getattr(OptimType, optimizer),
                               learning_rate=lr, eps=eps, sparse=sparse, weight_dist="uniform"), E


def _batches(rng, E, B, empty_table=None, mean=5.0):
    idx, off = [], []
    for k, e in enumerate(E):
        empty = k == empty_table
        i, o = ragged_batch(rng, B, e, 0.0 if empty else mean, 0.0 if empty else 3.0)
        idx.append(t(i))
        off.append(t(o))
    return idx, off


def _oracle(mod, p_shapes, q, ranks, before, idx, off, B, d_out):
    D = int(np.prod(q))
    outs, grads = [], []
    for k, p in enumerate(p_shapes):
        o_ = mod.layout.off[k]
        cores = [before[c][:, o_[c]:o_[c] + p[c]] for c in range(len(p))]
        row, tbl = O.compute_rowidx(off[k].cpu().numpy(), 1)
        i = idx[k].cpu().numpy()
        outs.append(O.tt_forward(1, B, D, p, q, ranks, O.make_L(p), len(i), i, row, tbl, cores)[0])
        if d_out is not None:
            grads.append(O.tt_backward_dense(D, p, q, ranks, O.make_L(p), len(i), i, row, tbl,
                                             d_out[k].cpu().numpy()[None], cores))
    cat_g = None
    if d_out is not None:
        cat_g = [np.concatenate([grads[k][c] for k in range(len(p_shapes))], axis=1) for c in range(len(p_shapes[0]))]
    return np.stack(outs), cat_g


def _tols(path, family):
    exact = path == "generic" or family == "generic"
    return (1e-5, 1e-4) if exact else (1e-3, 1e-2)  # (forward, fused state): north-star tolerances on tf32 paths


@pytest.mark.parametrize("family", list(FAMILIES))
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_fused_forward_matches_per_table_oracle(ext, path, family):
    ext.set_path(ext.PATH_GENERIC if path == "generic" else ext.PATH_AUTO)
    q, ranks = FAMILIES[family]["q"], FAMILIES[family]["ranks"]
    mod, E = _module(P3, q, ranks)
    rng = np.random.RandomState(21)
    B = 80
    idx, off = _batches(rng, E, B, empty_table=2)
    before = [c.detach().cpu().numpy().copy() for c in mod.tt_cores]
    with torch.no_grad():
        got = mod(idx, off)
    want, _ = _oracle(mod, P3, q, ranks, before, idx, off, B, None)
    assert got.shape == (len(P3), B, int(np.prod(q)))
    assert int(got[2].count_nonzero()) == 0  # the table without lookups pools nothing
    for k in range(len(P3)):
        if k != 2:
            assert rel_err(got[k].cpu().numpy(), want[k]) < _tols(path, family)[0], k


@pytest.mark.parametrize("p_shapes,q,ranks", [(P2, [8, 8], [12]), (P4, [2, 4, 2, 4], [3, 5, 4])])
def test_fused_forward_other_core_counts(ext, p_shapes, q, ranks):
    ext.set_path(ext.PATH_AUTO)  # T = 2 / 4 always take the generic kernels
    mod, E = _module(p_shapes, q, ranks)
    rng = np.random.RandomState(22)
    B = 40
    idx, off = _batches(rng, E, B)
    before = [c.detach().cpu().numpy().copy() for c in mod.tt_cores]
    d_out = torch.rand(len(p_shapes), B, int(np.prod(q)), device=DEV) * 0.1
    out = mod(idx, off)
    out.backward(d_out)
    want, cat_g = _oracle(mod, p_shapes, q, ranks, before, idx, off, B, d_out)
    np.testing.assert_allclose(out.detach().cpu().numpy(), want, rtol=1e-4, atol=1e-6)
    for a, b in zip(mod.tt_cores, O.sgd_step(before, cat_g, 0.1)):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("family", list(FAMILIES))
@pytest.mark.parametrize("optimizer", ["SGD", "EXACT_ADAGRAD"])
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_fused_step_matches_per_table_oracle(ext, path, optimizer, family):
    ext.set_path(ext.PATH_GENERIC if path == "generic" else ext.PATH_AUTO)
    q, ranks = FAMILIES[family]["q"], FAMILIES[family]["ranks"]
    lr, eps = (0.1, 0.0) if optimizer == "SGD" else (0.02, 1e-4)
    mod, E = _module(P3, q, ranks, optimizer, lr=lr, eps=eps)
    rng = np.random.RandomState(23)
    B = 64
    ftol, stol = _tols(path, family)
    for step in range(2):  # the second step reuses the pooled plan buffer and the zero-on-exit gradient scratch
        idx, off = _batches(rng, E, B, empty_table=step)
        before = [c.detach().cpu().numpy().copy() for c in mod.tt_cores]
        state0 = [s.cpu().numpy().copy() for s in mod.optimizer_state]
        d_out = torch.rand(len(P3), B, int(np.prod(q)), device=DEV) * 0.1
        out = mod(idx, off)
        out.backward(d_out)
        torch.cuda.synchronize()
        want, cat_g = _oracle(mod, P3, q, ranks, before, idx, off, B, d_out)
        assert rel_err(out.detach().cpu().numpy(), want) < ftol
        if optimizer == "SGD":
            for c, (a, b) in enumerate(zip(mod.tt_cores, O.sgd_step(before, cat_g, lr))):
                assert rel_err(a.detach().cpu().numpy(), b) < stol, (step, c)
        else:
            # Adagrad state = sum of squared gradients: compared directly; the weight update divides by
            # sqrt(state) + eps and amplifies a rounding difference where the state is ~0, so the update
            # arithmetic is checked from the path's OWN state
            new_c, new_s = O.adagrad_step(before, state0, cat_g, lr, eps)
            for c, (a, b) in enumerate(zip(mod.optimizer_state, new_s)):
                assert rel_err(a.cpu().numpy(), b) < stol, (step, c)
            if ftol < 1e-4:
                for c, (a, b) in enumerate(zip(mod.tt_cores, new_c)):
                    assert rel_err(a.detach().cpu().numpy(), b) < 1e-3, (step, c)
    flat, _ = ext.grad_scratch([c.data for c in mod.tt_cores])
    assert int(flat.count_nonzero()) == 0, "the fused backward must leave its gradient scratch zero"


@pytest.mark.parametrize("family", ["tcgen05", "warp_mma", "generic"])
def test_fused_equals_single_table_modules(ext, family):
    """Same weights loaded into one TTEmbeddingBag per table: pooled rows, dense gradients."""
    from fbtt_embedding_b200 import OptimType, TTEmbeddingBag

    ext.set_path(ext.PATH_AUTO)
    q, ranks = FAMILIES[family]["q"], FAMILIES[family]["ranks"]
    D = int(np.prod(q))
    mod, E = _module(P3, q, ranks, sparse=False)
    solo = [TTEmbeddingBag(E[k], D, ranks, P3[k], q, optimizer=OptimType.SGD, sparse=False, use_cache=False,
                           weight_dist="uniform") for k in range(len(P3))]
    with torch.no_grad():
        for k, m in enumerate(solo):
            for dst, src in zip(m.tt_cores, mod.table_cores(k)):
                dst.copy_(src)
    rng = np.random.RandomState(24)
    B = 72
    idx, off = _batches(rng, E, B, empty_table=4)
    d_out = torch.rand(len(P3), B, D, device=DEV) * 0.1
    out = mod(idx, off)
    out.backward(d_out)
    tol = 2e-5 if family == "generic" else 2e-3
    for k, m in enumerate(solo):
        o = m(idx[k], off[k])
        o.backward(d_out[k])
        assert rel_err(out[k].detach().cpu().numpy(), o.detach().cpu().numpy()) < tol or idx[k].numel() == 0
        off_k = mod.layout.off[k]
        for c, core in enumerate(m.tt_cores):
            mine = mod.tt_cores[c].grad[:, off_k[c]:off_k[c] + P3[k][c]]
            if idx[k].numel() == 0:
                assert int(mine.count_nonzero()) == 0
            else:
                assert rel_err(mine.cpu().numpy(), core.grad.cpu().numpy()) < tol, (k, c)


def test_fused_out_of_range_lookups_contribute_nothing(ext):
    """An index >= prod(p) of ITS table must not read the next table's slices."""
    for path in (ext.PATH_GENERIC, ext.PATH_AUTO):
        ext.set_path(path)
        q, ranks = FAMILIES["tcgen05"]["q"], FAMILIES["tcgen05"]["ranks"]
        mod, E = _module(P3, q, ranks)
        B = 4
        idx = [t(np.array([0, E[k] - 1, E[k], E[k] + 5], np.int64)) for k in range(len(P3))]
        off = [t(np.arange(0, 5, dtype=np.int64))] * len(P3)
        with torch.no_grad():
            got = mod(idx, off)
        assert int(got[:, 2:].count_nonzero()) == 0
        assert int(got[:, :2].count_nonzero()) > 0


def test_fused_step_is_five_launches_and_graph_capturable(ext):
    ext.set_path(ext.PATH_AUTO)
    q, ranks = FAMILIES["tcgen05"]["q"], FAMILIES["tcgen05"]["ranks"]
    mod, E = _module(P3, q, ranks)
    rng = np.random.RandomState(25)
    B = 64
    idx, off = _batches(rng, E, B)
    from fbtt_embedding_b200.fused import pack_table_major

    ci, co = pack_table_major(idx, off)
    d_out = torch.rand(len(P3), B, 64, device=DEV) * 0.1
    mod(ci, co).backward(d_out)  # warm-up: attributes, plan buffer
    torch.cuda.synchronize()
    n0 = ext.launch_count()
    mod(ci, co).backward(d_out)
    torch.cuda.synchronize()
    assert 0 < ext.launch_count() - n0 <= 5, "CSR->COO, plan, forward, backward, sweep -- for ALL tables"
    ref = [c.detach().clone() for c in mod.tt_cores]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        mod(ci, co).backward(d_out)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        mod(ci, co).backward(d_out)
    graph.replay()
    torch.cuda.synchronize()
    # two more SGD steps with the same batch and gradient happened since `ref` (stream warm-up + replay; the
    # capture itself executes nothing): compare with two eager steps on a twin
    twin, _ = _module(P3, q, ranks)
    with torch.no_grad():
        for a, b in zip(twin.tt_cores, ref):
            a.copy_(b)
    for _ in range(2):
        twin(ci, co).backward(d_out)
    torch.cuda.synchronize()
    for a, b in zip(mod.tt_cores, twin.tt_cores):
        assert rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy()) < 2e-3


# ---- fused exchange: the row map on real kernels, all "ranks" on this one GPU (LocalPeers) -----------------------
@pytest.mark.parametrize("family", ["tcgen05", "warp_mma", "generic"])
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_peer_exchange_three_ranks_on_one_gpu(ext, path, family):
    """Rank s adds the pooled rows of its tables into the batch-slice buffer of the rank that owns the row and reads
    its gradients from there (include/ttb.h, ttb_row_map_t).  Here the three ranks' buffers are three allocations
    on one device -- one address space, as peer-mapped NVLink memory is -- so the mapped plan / generic kernels
    are validated without a second GPU; the phases of all ranks are driven in turn on one stream."""
    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200.sharded import LocalPeers, TableShardedTTEmbeddingBag

    ext.set_path(ext.PATH_GENERIC if path == "generic" else ext.PATH_AUTO)
    q, ranks = FAMILIES[family]["q"], FAMILIES[family]["ranks"]
    D = int(np.prod(q))
    W, B, T = 3, 48, len(P3)
    bw = B // W
    E = [int(np.prod(p)) for p in P3]
    specs = [dict(num_embeddings=E[k], embedding_dim=D, tt_ranks=ranks, tt_p_shapes=P3[k], tt_q_shapes=q) for k in range(T)]
    torch.manual_seed(31)
    mods = [TableShardedTTEmbeddingBag(specs, None, fused=True, exchange="peer", world_size=W, rank=r,
                                       optimizer=OptimType.SGD, learning_rate=0.1, sparse=True, weight_dist="uniform")
            for r in range(W)]
    peers = LocalPeers(W, bw, T, D, DEV)
    views = [peers.view(r) for r in range(W)]
    for m, v in zip(mods, views):
        m._peer_setup(v, B)
    rng = np.random.RandomState(32)
    idx, off = _batches(rng, E, B, empty_table=1)
    before = [[c.detach().cpu().numpy().copy() for c in m.fused.tt_cores] for m in mods]
    states = [m._phase_forward(v, [idx[t] for t in m.local_tables], [off[t] for t in m.local_tables])
              for m, v in zip(mods, views)]
    torch.cuda.synchronize()
    ftol, stol = _tols(path, family)
    pooled = np.zeros((T, B, D), np.float32)
    for r, m in enumerate(mods):
        sub = [P3[t] for t in m.local_tables]
        o, _ = _oracle(m.fused, sub, q, ranks, before[r], [idx[t] for t in m.local_tables],
                       [off[t] for t in m.local_tables], B, None)
        for k, t in enumerate(m.local_tables):
            pooled[t] = o[k]
    for w in range(W):
        want = pooled[:, w * bw:(w + 1) * bw].transpose(1, 0, 2)
        assert rel_err(views[w].x.cpu().numpy(), want) < ftol, w
    d_x = [torch.rand(bw, T, D, device=DEV) * 0.1 for _ in range(W)]
    for w in range(W):
        views[w].dx.copy_(d_x[w])
    d_pooled = torch.cat([d.permute(1, 0, 2) for d in d_x], dim=1)  # [T, B, D]
    for m, v, st in zip(mods, views, states):
        m._phase_backward(v, st)
    torch.cuda.synchronize()
    for r, m in enumerate(mods):
        sub = [P3[t] for t in m.local_tables]
        _, cat_g = _oracle(m.fused, sub, q, ranks, before[r], [idx[t] for t in m.local_tables],
                           [off[t] for t in m.local_tables], B, d_pooled[m.local_tables].contiguous())
        for c, (a, b) in enumerate(zip(m.fused.tt_cores, O.sgd_step(before[r], cat_g, 0.1))):
            assert rel_err(a.detach().cpu().numpy(), b) < stol, (r, c)
