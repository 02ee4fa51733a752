"""GPU tests of table groups (fbtt_embedding_b200/grouped.py over `ttb_group_*`): a group of
differently-shaped tables must give what the per-table modules give -- and what the oracle gives -- for the
forward, the fused SGD / Adagrad step and the dense gradients, on one lane and on several, eagerly and inside
a CUDA graph.  (File name sorts last on purpose: this is the newest, least exercised path.)"""
import numpy as np
import pytest
import torch

from oracle import tt_oracle as O
from tests.helpers import ragged_batch, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
D = 64

# one table per kernel family: tcgen05 (r=32), warp-MMA (r=16), generic FFMA (T=2, T=4, odd ranks), and a tiny
# table whose lookups all share one bucket
SPECS = [
    dict(tt_p_shapes=[20, 22, 25], tt_q_shapes=[4, 4, 4], tt_ranks=[32, 32]),
    dict(tt_p_shapes=[7, 9, 11], tt_q_shapes=[4, 4, 4], tt_ranks=[16, 16]),
    dict(tt_p_shapes=[30, 40], tt_q_shapes=[8, 8], tt_ranks=[12]),
    dict(tt_p_shapes=[3, 4, 5, 6], tt_q_shapes=[2, 4, 2, 4], tt_ranks=[3, 5, 4]),
    dict(tt_p_shapes=[1, 1, 3], tt_q_shapes=[4, 4, 4], tt_ranks=[32, 32]),
    dict(tt_p_shapes=[6, 5, 4], tt_q_shapes=[4, 2, 8], tt_ranks=[5, 7]),
]


def _specs():
    return [dict(num_embeddings=int(np.prod(s["tt_p_shapes"])), embedding_dim=D, **s) for s in SPECS]


@pytest.fixture(scope="module")
def ext():
    from fbtt_embedding_b200 import tt_embeddings as e

    yield e
    e.set_path(e.PATH_AUTO)
    e.group_set_streams(1)


def t(x):
    return torch.as_tensor(np.ascontiguousarray(x), device=DEV)


def _batches(rng, B, empty_table=None):
    idx, off = [], []
    for n, s in enumerate(_specs()):
        empty = n == empty_table  # mean 0, std 0: every bag of that table is empty
        i, o = ragged_batch(rng, B, s["num_embeddings"], 0.0 if empty else 5.0, 0.0 if empty else 3.0)
        idx.append(t(i))
        off.append(t(o))
    return idx, off


def _pair(optimizer, sparse=True, eps=1e-4):
    """(group, per-table modules) with identical weights."""
    lr = 0.1 if optimizer == "SGD" else 0.02
    from fbtt_embedding_b200 import OptimType, TTEmbeddingBag, TTEmbeddingBagGroup

    torch.manual_seed(3)
    opt = getattr(OptimType, optimizer)
    grp = TTEmbeddingBagGroup(_specs(), optimizer=opt, learning_rate=lr, eps=eps, sparse=sparse, weight_dist="uniform")
    solo = [TTEmbeddingBag(**s, optimizer=opt, learning_rate=lr, eps=eps, sparse=sparse, use_cache=False,
                           weight_dist="uniform") for s in _specs()]
    with torch.no_grad():
        for a, b in zip(solo, grp.tables):
            for ca, cb in zip(a.tt_cores, b.tt_cores):
                ca.copy_(cb)
    return grp, solo


def _tol(path):
    # exact path: same kernels, only the order of fp32 atomics differs; auto: tf32 operands on the tensor-core
    # families, and the bucket plan orders lookups by an atomic counter (two runs tile differently)
    return 2e-5 if path == "generic" else 2e-3


@pytest.mark.parametrize("lanes", [1, 3])
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_group_forward_equals_per_table_modules_and_oracle(ext, path, lanes):
    ext.set_path(ext.PATH_GENERIC if path == "generic" else ext.PATH_AUTO)
    ext.group_set_streams(lanes)
    grp, solo = _pair("SGD")
    rng = np.random.RandomState(11)
    B = 96
    idx, off = _batches(rng, B, empty_table=3)
    with torch.no_grad():
        got = grp(idx, off)
        want = torch.stack([m(i, o) for m, i, o in zip(solo, idx, off)])
    assert got.shape == (len(SPECS), B, D)
    assert rel_err(got.cpu().numpy(), want.cpu().numpy()) < _tol(path)
    assert int(got[3].count_nonzero()) == 0  # the table without lookups pools nothing
    for n, s in enumerate(_specs()):  # and straight against the oracle, table by table
        p, q, ranks = s["tt_p_shapes"], s["tt_q_shapes"], s["tt_ranks"]
        cores = [c.detach().cpu().numpy() for c in grp.tables[n].tt_cores]
        rowidx, tableidx = O.compute_rowidx(off[n].cpu().numpy(), 1)
        i_np = idx[n].cpu().numpy()
        ref = O.tt_forward(1, B, D, p, q, ranks, O.make_L(p), len(i_np), i_np, rowidx, tableidx, cores)[0]
        assert rel_err(got[n].cpu().numpy(), ref) < (1e-5 if path == "generic" else 1e-3) or len(i_np) == 0


@pytest.mark.parametrize("lanes", [1, 4])
@pytest.mark.parametrize("optimizer", ["SGD", "EXACT_ADAGRAD"])
@pytest.mark.parametrize("path", ["generic", "auto"])
def test_group_fused_step_equals_per_table_modules(ext, path, optimizer, lanes):
    ext.set_path(ext.PATH_GENERIC if path == "generic" else ext.PATH_AUTO)
    ext.group_set_streams(lanes)
    grp, solo = _pair(optimizer)
    rng = np.random.RandomState(12)
    B = 64
    for step in range(3):
        idx, off = _batches(rng, B, empty_table=step % len(SPECS))
        d_out = torch.rand(len(SPECS), B, D, device=DEV) * 0.1
        out = grp(idx, off)
        out.backward(d_out)
        for n, m in enumerate(solo):
            m(idx[n], off[n]).backward(d_out[n])
    torch.cuda.synchronize()
    for n, m in enumerate(solo):
        for k, (a, b) in enumerate(zip(grp.tables[n].tt_cores, m.tt_cores)):
            # Adagrad divides by sqrt(state) + eps: where the state is ~0 a gradient rounding difference is
            # amplified, so compare at the looser north-star bound for fused state (1e-2) on that optimizer
            tol = _tol(path) if optimizer == "SGD" else max(_tol(path), 1e-2 if path == "auto" else 1e-4)
            assert rel_err(a.detach().cpu().numpy(), b.detach().cpu().numpy()) < tol, (n, k)
        if optimizer != "SGD":
            for a, b in zip(grp.tables[n].optimizer_state, m.optimizer_state):
                assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < max(_tol(path), 5e-3 if path == "auto" else 1e-5)
    assert int(grp._group._grad_flat.count_nonzero()) == 0, "the group's gradient scratch must come back zero"


def test_group_fused_sgd_matches_oracle(ext):
    ext.set_path(ext.PATH_GENERIC)
    ext.group_set_streams(2)
    grp, _ = _pair("SGD")
    rng = np.random.RandomState(13)
    B, lr = 48, 0.1
    idx, off = _batches(rng, B)
    before = [[c.detach().cpu().numpy().copy() for c in tbl.tt_cores] for tbl in grp.tables]
    d_out = torch.rand(len(SPECS), B, D, device=DEV) * 0.1
    grp(idx, off).backward(d_out)
    torch.cuda.synchronize()
    for n, s in enumerate(_specs()):
        p, q, ranks = s["tt_p_shapes"], s["tt_q_shapes"], s["tt_ranks"]
        rowidx, tableidx = O.compute_rowidx(off[n].cpu().numpy(), 1)
        i_np = idx[n].cpu().numpy()
        g = O.tt_backward_dense(D, p, q, ranks, O.make_L(p), len(i_np), i_np, rowidx, tableidx,
                                d_out[n].cpu().numpy()[None], before[n])
        want = O.sgd_step(before[n], g, lr)
        for k in range(len(p)):
            np.testing.assert_allclose(grp.tables[n].tt_cores[k].detach().cpu().numpy(), want[k], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_group_dense_gradients_equal_per_table_modules(ext, path):
    ext.set_path(ext.PATH_GENERIC if path == "generic" else ext.PATH_AUTO)
    ext.group_set_streams(2)
    grp, solo = _pair("SGD", sparse=False)
    rng = np.random.RandomState(14)
    B = 64
    idx, off = _batches(rng, B, empty_table=1)
    d_out = torch.rand(len(SPECS), B, D, device=DEV) * 0.1
    grp(idx, off).backward(d_out)
    for n, m in enumerate(solo):
        m(idx[n], off[n]).backward(d_out[n])
        for a, b in zip(grp.tables[n].tt_cores, m.tt_cores):
            assert a.grad is not None and a.grad.shape == a.shape
            if n == 1:
                assert int(a.grad.count_nonzero()) == 0
            else:
                assert rel_err(a.grad.cpu().numpy(), b.grad.cpu().numpy()) < _tol(path)


@pytest.mark.parametrize("lanes", [1, 3])
def test_group_step_inside_a_cuda_graph(ext, lanes):
    """Fork / join of the lanes is event record + wait, which a stream capture follows."""
    ext.set_path(ext.PATH_AUTO)
    ext.group_set_streams(lanes)
    grp, solo = _pair("SGD")
    rng = np.random.RandomState(15)
    B = 64
    idx, off = _batches(rng, B)
    d_out = torch.rand(len(SPECS), B, D, device=DEV) * 0.1
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        grp(idx, off).backward(d_out)  # warm-up outside the capture (lazy stream / attribute creation)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = grp(idx, off)
        out.backward(d_out)
    graph.replay()
    torch.cuda.synchronize()
    for n, m in enumerate(solo):  # warm-up step + replayed step == two eager steps of the per-table modules
        for _ in range(2):
            m(idx[n], off[n]).backward(d_out[n])
    torch.cuda.synchronize()
    want = torch.stack([m(i, o) for m, i, o in zip(solo, idx, off)])
    got = grp(idx, off)
    assert rel_err(got.detach().cpu().numpy(), want.detach().cpu().numpy()) < 5e-3


def test_group_rejects_bad_input(ext):
    ext.group_set_streams(1)
    grp, _ = _pair("SGD")
    rng = np.random.RandomState(16)
    idx, off = _batches(rng, 32)
    with pytest.raises(RuntimeError):
        grp(idx[:-1], off[:-1])
    off2 = list(off)
    off2[2] = off2[2][:-1]
    with pytest.raises(RuntimeError):
        grp(idx, off2)
    with pytest.raises(RuntimeError):
        grp([i.cpu() for i in idx], off)
    out = grp(idx, off)  # the group still works after the rejected calls
    assert out.shape == (len(SPECS), 32, D)


def test_table_sharded_grouped_single_rank(ext):
    """TableShardedTTEmbeddingBag(grouped=True) on a 1-rank group == the per-table path (exchange is the identity)."""
    import os

    import torch.distributed as dist

    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200.sharded import TableShardedTTEmbeddingBag

    ext.set_path(ext.PATH_GENERIC)
    created = False
    if not dist.is_initialized():
        import socket

        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
        created = True
    try:
        kw = dict(optimizer=OptimType.SGD, learning_rate=0.1, sparse=True, weight_dist="uniform")
        a = TableShardedTTEmbeddingBag(_specs(), grouped=True, **kw)
        b = TableShardedTTEmbeddingBag(_specs(), grouped=False, **kw)
        b.load_state_dict(a.state_dict())
        rng = np.random.RandomState(17)
        B = 32
        idx, off = _batches(rng, B)
        g = torch.rand(B, len(SPECS), D, device=DEV) * 0.1
        oa, ob = a(idx, off), b(idx, off)
        assert rel_err(oa.detach().cpu().numpy(), ob.detach().cpu().numpy()) < 2e-5
        oa.backward(g)
        ob.backward(g)
        for x, y in zip(a.parameters(), b.parameters()):
            assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 2e-5
    finally:
        if created:
            dist.destroy_process_group()
