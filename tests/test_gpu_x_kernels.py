"""The tcgen05 / bf16-operand kernel family (csrc/ttb_tt_x.cuh: ranks 16 / 32 / 64 / 128, split-precision operands) and the
CSR entry points (plan kernel derives each lookup's bag; optimizer applied inside the backward kernel).

Because every fp32 operand is split hi + lo and accumulated in three terms, this path is held to fp32-GRADE bounds
(1e-5-ish, element-wise as well as max-norm), far inside the north-star tolerances (1e-3 forward, 1e-2 backward state);
a regression to single-term bf16 (2e-3) or tf32 (2.5e-4) products fails these tests."""
import numpy as np
import pytest
import torch

from oracle import tt_oracle as O
from tests.helpers import elem_close, make_cores, ragged_batch, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ext():
    from fbtt_embedding_b200 import tt_embeddings as e

    e.set_path(e.PATH_AUTO)
    yield e
    e.set_path(e.PATH_AUTO)


def t(x):
    return torch.as_tensor(np.ascontiguousarray(x), device=DEV)


SHAPES = [
    dict(p=[20, 22, 25], q=[4, 4, 4], ranks=[32, 32]),   # README family
    dict(p=[10, 12, 14], q=[4, 4, 8], ranks=[32, 32]),   # config 5, r = 32
    dict(p=[8, 10, 12], q=[4, 4, 8], ranks=[64, 64]),    # config 4 / 5, r = 64: two column blocks
    dict(p=[6, 8, 10], q=[4, 2, 4], ranks=[64, 64]),     # q1 = 2: one column block of two j1 groups
    dict(p=[4, 6, 8], q=[4, 4, 8], ranks=[128, 128]),    # r = 128: four column blocks, 512 TMEM columns
    dict(p=[5, 3, 7], q=[4, 1, 4], ranks=[128, 128]),    # q1 = 1
    dict(p=[9, 11, 13], q=[4, 4, 8], ranks=[16, 16]),    # config 5, r = 16: ONE 64-column block (M = 64 dB1 accumulator)
    dict(p=[7, 9, 11], q=[4, 8, 4], ranks=[16, 16]),     # r = 16, q1 = 8: two 64-column blocks
]


@pytest.mark.parametrize("skew", [False, True])
@pytest.mark.parametrize("num_tables", [1, 2])
@pytest.mark.parametrize("shape", SHAPES)
def test_x_kernels_are_fp32_grade(ext, shape, num_tables, skew):
    """Forward, dense gradients, fused SGD and fused Adagrad through the CSR entry points against the fp64 oracle.
    skew=True concentrates the lookups on few middle-core indices: buckets longer than max_run tiles are split over
    several runs (partial dCore1 blocks meet in the scratch and are swept after the grid barrier); the others are
    applied straight from TMEM.  Both routes must give the same fp32-grade answer."""
    p, q, ranks = shape["p"], shape["q"], shape["ranks"]
    assert ext.csr_supported(num_tables, 64, int(np.prod(q)), p, q, [1] + ranks + [1], 100)
    rng = np.random.RandomState(3 + num_tables + (7 if skew else 0))
    E, D, B = int(np.prod(p)), int(np.prod(q)), 53
    cores = make_cores(rng, num_tables, p, q, ranks)
    idx, off = ragged_batch(rng, B, E, 40 if skew else 9, 6, num_tables)
    if skew:  # two middle-core indices take most of the batch
        L = O.make_L(p)
        hot = rng.rand(len(idx)) < 0.8
        i1 = (idx // L[1]) % p[1]
        idx = np.where(hot, idx - i1 * L[1] + (rng.randint(0, 2, size=len(idx)) * L[1]), idx).astype(np.int64)
    idx[:3] = [E - 1, 0, idx[3]]
    nnz = len(idx)
    R = [1] + ranks + [1]
    r0, t0 = O.compute_rowidx(off, num_tables)
    want = O.tt_forward(num_tables, B, D, p, q, ranks, O.make_L(p), nnz, idx, r0, t0, cores, dtype=np.float64)
    out = ext.tt_forward_csr(num_tables, B, D, p, q, R, t(idx), t(off), [t(c) for c in cores])
    assert rel_err(out.cpu().numpy(), want) < 2e-5
    ok, worst = elem_close(out.cpu().numpy(), want, rtol=1e-3, atol=1e-5 * float(np.abs(want).max()))
    assert ok, f"element-wise 1e-3 bound exceeded {worst:.2f}x"
    dout = rng.uniform(-1, 1, size=(num_tables, B, D)).astype(np.float32)
    g_want = O.tt_backward_dense(D, p, q, ranks, O.make_L(p), nnz, idx, r0, t0, dout, cores)
    i_t, o_t = t(idx), t(off)
    grads = ext.tt_backward_csr(ext.OPTIM_DENSE, D, 0.0, 0.0, p, q, R, i_t, o_t, t(dout), [t(c) for c in cores])
    for i in range(3):
        assert rel_err(grads[i].cpu().numpy(), g_want[i]) < 5e-5, f"dense gradient of core {i}"
    lr, eps = 0.05, 1e-3
    cs = [t(c) for c in cores]
    ext.tt_backward_csr(ext.OPTIM_SGD, D, lr, 0.0, p, q, R, i_t, o_t, t(dout), cs)
    w_want = O.sgd_step(cores, g_want, lr)
    for i in range(3):
        assert rel_err(cs[i].cpu().numpy(), w_want[i]) < 5e-5, f"fused SGD, core {i}"
    state0 = [rng.uniform(0.05, 0.3, size=c.shape).astype(np.float32) for c in cores]
    cs, st = [t(c) for c in cores], [t(s) for s in state0]
    ext.tt_backward_csr(ext.OPTIM_ADAGRAD, D, lr, eps, p, q, R, i_t, o_t, t(dout), cs, st)
    w_want, s_want = O.adagrad_step(cores, state0, g_want, lr, eps)
    for i in range(3):
        assert rel_err(st[i].cpu().numpy(), s_want[i]) < 1e-4, f"Adagrad state, core {i}"
        assert rel_err(cs[i].cpu().numpy(), w_want[i]) < 5e-4, f"fused Adagrad, core {i}"  # lr / (sqrt(s) + eps) amplifies
    # the fused modes hand back an all-zero gradient scratch and zero sync words: a second step on fresh weights
    # gives the same answer
    cs2 = [t(c) for c in cores]
    ext.tt_backward_csr(ext.OPTIM_SGD, D, lr, 0.0, p, q, R, i_t, o_t, t(dout), cs2)
    for i in range(3):
        assert rel_err(cs2[i].cpu().numpy(), O.sgd_step(cores, g_want, lr)[i]) < 5e-5
    flat, _ = ext.grad_scratch(cs2)
    assert int(flat.count_nonzero()) == 0


def test_csr_equals_coo_and_drops_uncovered_lookups(ext):
    """offsets[0] > 0 and offsets[-1] < nnz: lookups outside every bag are ignored (they have no row to pool into);
    empty bags at both ends; the CSR result equals the reference op sequence (preprocess -> tt_forward)."""
    p, q, ranks = [20, 22, 25], [4, 4, 4], [32, 32]
    R = [1] + ranks + [1]
    E, D, B = int(np.prod(p)), 64, 40
    rng = np.random.RandomState(5)
    cores = [t(c) for c in make_cores(rng, 1, p, q, ranks)]
    lens = rng.randint(0, 9, size=B)
    lens[[0, 1, B - 1]] = 0
    off = (np.concatenate([[0], np.cumsum(lens)]) + 7).astype(np.int64)
    nnz = int(off[-1]) + 5
    idx = rng.randint(0, E, size=nnz).astype(np.int64)
    out = ext.tt_forward_csr(1, B, D, p, q, R, t(idx), t(off), cores)
    covered = idx[7:int(off[-1])]
    off0 = off - 7
    e64, e32 = torch.empty(0, dtype=torch.int64, device=DEV), torch.empty(0, dtype=torch.int32, device=DEV)
    col, row, tbl, n, _ = ext.preprocess_indices_sync(t(covered), t(off0), 1, True, e64, e32)
    want = ext.tt_forward(1000, 1, B, D, p, q, R, t(O.make_L(p)), n, col, row, tbl, cores)
    assert rel_err(out.cpu().numpy(), want.cpu().numpy()) < 1e-6


@pytest.mark.parametrize("optimizer", ["SGD", "EXACT_ADAGRAD"])
def test_module_step_is_three_launches_and_matches_reference_op_sequence(ext, optimizer):
    """The cache-less module goes plan -> forward -> backward(+optimizer): 3 libttb launches per training step, no
    preprocess launch, no sweep launch; same weights as the reference's op sequence (csr_fast_path=False:
    preprocess_indices_sync -> tt_forward -> tt_*_backward) and as the oracle."""
    from fbtt_embedding_b200 import OptimType, TTEmbeddingBag

    p, q, ranks = [20, 22, 25], [4, 4, 4], [32, 32]
    E, D, B = int(np.prod(p)), 64, 128
    kw = dict(tt_p_shapes=p, tt_q_shapes=q, tt_ranks=ranks, optimizer=getattr(OptimType, optimizer), learning_rate=0.1,
              eps=1e-3, use_cache=False, sparse=True, weight_dist="uniform")
    a, b = TTEmbeddingBag(E, D, **kw), TTEmbeddingBag(E, D, **kw)
    b.csr_fast_path = False
    with torch.no_grad():
        for x, y in zip(a.tt_cores, b.tt_cores):
            y.copy_(x)
    rng = np.random.RandomState(9)
    for step in range(3):
        idx, off = ragged_batch(rng, B, E, 12, 4)
        g = torch.rand(B, D, device=DEV) * 0.1
        n0 = ext.launch_count()
        oa = a(t(idx), t(off))
        oa.backward(g)
        assert ext.launch_count() - n0 == 3, "plan + forward + backward(with optimizer)"
        ob = b(t(idx), t(off))
        ob.backward(g)
        assert rel_err(oa.detach().cpu().numpy(), ob.detach().cpu().numpy()) < 1e-5
        for x, y in zip(a.tt_cores, b.tt_cores):
            assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 2e-4
        for x, y in zip(a.optimizer_state, b.optimizer_state):
            if x.numel():
                assert rel_err(x.cpu().numpy(), y.cpu().numpy()) < 1e-4
    # inference: no plan is left behind
    ext._plan_cache.clear()
    with torch.no_grad():
        a(t(idx), t(off))
    assert len(ext._plan_cache) == 0


def test_module_step_captures_in_a_cuda_graph(ext):
    """Cooperative launches (plan, backward) inside a captured graph: replays reproduce the eager step."""
    from fbtt_embedding_b200 import OptimType, TTEmbeddingBag

    p, q, ranks = [20, 22, 25], [4, 4, 4], [32, 32]
    E, D, B = int(np.prod(p)), 64, 128
    kw = dict(tt_p_shapes=p, tt_q_shapes=q, tt_ranks=ranks, optimizer=OptimType.SGD, learning_rate=0.1, use_cache=False,
              sparse=True, weight_dist="uniform")
    a, b = TTEmbeddingBag(E, D, **kw), TTEmbeddingBag(E, D, **kw)
    with torch.no_grad():
        for x, y in zip(a.tt_cores, b.tt_cores):
            y.copy_(x)
    rng = np.random.RandomState(13)
    idx0, off0 = ragged_batch(rng, B, E, 12, 0)  # fixed bag length: the static buffers keep their size
    s_idx, s_off = t(idx0).clone(), t(off0).clone()
    g = torch.rand(B, D, device=DEV) * 0.1
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            a(s_idx, s_off).backward(g)
            b(s_idx, s_off).backward(g)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = a(s_idx, s_off)
        out.backward(g)
    for step in range(3):
        idx, _ = ragged_batch(rng, B, E, 12, 0)
        s_idx.copy_(t(idx))
        graph.replay()
        ob = b(s_idx, s_off)
        ob.backward(g)
        torch.cuda.synchronize()
        assert rel_err(out.detach().cpu().numpy(), ob.detach().cpu().numpy()) < 1e-5
        for x, y in zip(a.tt_cores, b.tt_cores):
            assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# bf16 CORES (BASELINE configs[2]: "bf16 cores / fp32 accum"): the reference is fp32-only, so the oracle is fed the same
# bf16-representable values as fp32
# ---------------------------------------------------------------------------------------------------------------------
def bf16_round(x):
    return torch.as_tensor(x).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[1], SHAPES[2], SHAPES[4], SHAPES[6]])
def test_bf16_cores_match_the_oracle_on_the_same_values(ext, shape):
    p, q, ranks = shape["p"], shape["q"], shape["ranks"]
    R = [1] + ranks + [1]
    rng = np.random.RandomState(17)
    E, D, B = int(np.prod(p)), int(np.prod(q)), 61
    cores = [bf16_round(c) for c in make_cores(rng, 1, p, q, ranks)]  # fp32 arrays holding bf16-representable values
    idx, off = ragged_batch(rng, B, E, 11, 5)
    nnz = len(idx)
    r0, t0 = O.compute_rowidx(off, 1)

    def dev_cores():
        return [t(c).to(torch.bfloat16) for c in cores]

    want = O.tt_forward(1, B, D, p, q, ranks, O.make_L(p), nnz, idx, r0, t0, cores, dtype=np.float64)
    out = ext.tt_forward_csr(1, B, D, p, q, R, t(idx), t(off), dev_cores())
    assert out.dtype == torch.float32
    assert rel_err(out.cpu().numpy(), want) < 2e-5, "bf16 x bf16 products are exact in the fp32 accumulator"
    dout = rng.uniform(-1, 1, size=(1, B, D)).astype(np.float32)
    g_want = O.tt_backward_dense(D, p, q, ranks, O.make_L(p), nnz, idx, r0, t0, dout, cores)
    grads = ext.tt_backward_csr(ext.OPTIM_DENSE, D, 0.0, 0.0, p, q, R, t(idx), t(off), t(dout), dev_cores())
    for i in range(3):
        assert grads[i].dtype == torch.float32
        assert rel_err(grads[i].cpu().numpy(), g_want[i]) < 5e-5, f"dense gradient of core {i}"
    # fused SGD / Adagrad: the update is computed in fp32 and the weight rounded back to bf16: within two bf16 ulps (2^-6 of the value: an ulp is 2^-7 of the binade) of the fp32 result
    lr, eps = 0.05, 1e-3
    cs = dev_cores()
    ext.tt_backward_csr(ext.OPTIM_SGD, D, lr, 0.0, p, q, R, t(idx), t(off), t(dout), cs)
    w_want = O.sgd_step(cores, g_want, lr)
    for i in range(3):
        assert cs[i].dtype == torch.bfloat16
        ok, worst = elem_close(cs[i].float().cpu().numpy(), w_want[i], rtol=2.0 ** -6, atol=2e-5)
        assert ok, f"fused SGD on bf16 core {i}: {worst:.2f}x the rounding bound"
    state0 = [rng.uniform(0.05, 0.3, size=c.shape).astype(np.float32) for c in cores]
    cs, st = dev_cores(), [t(s) for s in state0]
    ext.tt_backward_csr(ext.OPTIM_ADAGRAD, D, lr, eps, p, q, R, t(idx), t(off), t(dout), cs, st)
    w_want, s_want = O.adagrad_step(cores, state0, g_want, lr, eps)
    for i in range(3):
        assert st[i].dtype == torch.float32 and rel_err(st[i].cpu().numpy(), s_want[i]) < 1e-4
        ok, worst = elem_close(cs[i].float().cpu().numpy(), w_want[i], rtol=2.0 ** -6, atol=2e-5)
        assert ok, f"fused Adagrad on bf16 core {i}: {worst:.2f}x the rounding bound"


@pytest.mark.parametrize("async_cache", [False, True])
def test_bf16_module_with_cache_and_adagrad(ext, async_cache):
    """BASELINE configs[2] at reduced size: bf16 cores, EXACT_ADAGRAD, LFU cache populated, one steady-state step against
    the oracle (TT half on the bf16 values, cached half on the fp32 cache rows)."""
    from fbtt_embedding_b200 import OptimType, TTEmbeddingBag
    from tests.helpers import collision_free_keys

    p, q, ranks = [20, 22, 25], [4, 4, 4], [32, 32]
    E, D, B, H = int(np.prod(p)), 64, 96, 8192
    emb = TTEmbeddingBag(E, D, ranks, p, q, optimizer=OptimType.EXACT_ADAGRAD, learning_rate=0.05, eps=1e-3, sparse=True,
                         use_cache=True, cache_size=64, hashtbl_size=H, weight_dist="uniform", async_cache=async_cache,
                         core_dtype=torch.bfloat16)
    assert all(c.dtype == torch.bfloat16 for c in emb.tt_cores) and emb.cache_weight.dtype == torch.float32
    rng = np.random.RandomState(4)
    hot = collision_free_keys(H, 400, rng, O.murmur_hash_3_32_i64)
    hot = hot[hot < E]

    def batch():
        _, off = ragged_batch(rng, B, E, 7, 2)
        w = 1.0 / np.arange(1, len(hot) + 1)
        return rng.choice(hot, size=int(off[-1]), p=w / w.sum()).astype(np.int64), off

    for _ in range(3):
        idx, off = batch()
        with torch.no_grad():
            emb(t(idx), t(off))
    emb.cache_populate()
    # cache rows = the fp32 chain on the bf16 values
    W = O.tt_matrix_to_full(p, q, ranks, [c.detach().float().cpu().numpy()[0] for c in emb.tt_cores])
    cached_keys = {int(k): int(s) for k, s in zip(emb.hashtbl.cpu(), emb.cache_state.cpu()) if k != -1 and s >= 0}
    assert len(cached_keys) > 10
    cw = emb.cache_weight.detach().cpu().numpy()
    for k, slot in list(cached_keys.items())[:20]:
        np.testing.assert_allclose(cw[slot], W[k], rtol=1e-5, atol=1e-7)
    idx, off = batch()
    cores0 = [c.detach().float().cpu().numpy() for c in emb.tt_cores]
    out = emb(t(idx), t(off))
    want = np.zeros((B, D), np.float64)
    r0, _ = O.compute_rowidx(off, 1)
    for n, k in enumerate(idx):
        want[r0[n]] += cw[cached_keys[int(k)]] if int(k) in cached_keys else W[k]
    assert rel_err(out.detach().cpu().numpy(), want) < 2e-5
    g = torch.rand(B, D, device=DEV) * 0.1
    out.backward(g)
    tt = np.array([int(k) not in cached_keys for k in idx])
    assert 0 < tt.sum() < len(idx)
    zt = np.zeros(int(tt.sum()), np.int64)
    grads = O.tt_backward_dense(D, p, q, ranks, O.make_L(p), int(tt.sum()), idx[tt], r0[tt], zt, g.cpu().numpy()[None], cores0)
    w_want, s_want = O.adagrad_step(cores0, [np.zeros_like(c) for c in cores0], grads, 0.05, 1e-3)
    for i in range(3):
        assert rel_err(emb.optimizer_state[i].cpu().numpy(), s_want[i]) < 1e-4
        ok, worst = elem_close(emb.tt_cores[i].detach().float().cpu().numpy(), w_want[i], rtol=2.0 ** -6, atol=2e-5)
        assert ok, f"bf16 core {i} after the cached Adagrad step: {worst:.2f}x the rounding bound"


def test_module_reads_pinned_host_inputs_in_place(ext):
    """Zero-copy CSR inputs: pinned host index / offset tensors are read by the plan kernel over PCIe; the result and the
    fused update equal those of the same step fed from device tensors."""
    from fbtt_embedding_b200 import OptimType, TTEmbeddingBag

    p, q, ranks = [20, 22, 25], [4, 4, 4], [32, 32]
    E, D, B = int(np.prod(p)), 64, 96
    kw = dict(tt_p_shapes=p, tt_q_shapes=q, tt_ranks=ranks, optimizer=OptimType.SGD, learning_rate=0.1, use_cache=False,
              sparse=True, weight_dist="uniform")
    a, b = TTEmbeddingBag(E, D, **kw), TTEmbeddingBag(E, D, **kw)
    with torch.no_grad():
        for x, y in zip(a.tt_cores, b.tt_cores):
            y.copy_(x)
    rng = np.random.RandomState(31)
    idx, off = ragged_batch(rng, B, E, 9, 3)
    g = torch.rand(B, D, device=DEV) * 0.1
    oa = a(torch.from_numpy(idx).pin_memory(), torch.from_numpy(off).pin_memory())
    ob = b(t(idx), t(off))
    assert oa.is_cuda and rel_err(oa.detach().cpu().numpy(), ob.detach().cpu().numpy()) < 1e-6
    oa.backward(g)
    ob.backward(g)
    for x, y in zip(a.tt_cores, b.tt_cores):
        assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 1e-5
    # pageable host tensors are copied to the device first (no zero-copy without pinning)
    oc = a(torch.from_numpy(idx), torch.from_numpy(off))
    assert oc.is_cuda


def test_module_dense_mode_routes_core_gradients(ext):
    """sparse=False through TTCsrLookupFunction: the dense core gradients land in `.grad` of the right parameters
    (and equal the oracle's); nothing is updated in place."""
    from fbtt_embedding_b200 import OptimType, TTEmbeddingBag

    p, q, ranks = [20, 22, 25], [4, 4, 4], [32, 32]
    E, D, B = int(np.prod(p)), 64, 80
    emb = TTEmbeddingBag(E, D, ranks, p, q, optimizer=OptimType.SGD, learning_rate=0.1, sparse=False, use_cache=False,
                         weight_dist="uniform")
    rng = np.random.RandomState(41)
    idx, off = ragged_batch(rng, B, E, 8, 3)
    cores0 = [c.detach().cpu().numpy().copy() for c in emb.tt_cores]
    n0 = ext.launch_count()
    out = emb(t(idx), t(off))
    g = torch.rand(B, D, device=DEV) * 0.1
    out.backward(g)
    assert ext.launch_count() - n0 == 3
    r0, t0 = O.compute_rowidx(off, 1)
    want = O.tt_backward_dense(D, p, q, ranks, O.make_L(p), len(idx), idx, r0, t0, g.cpu().numpy()[None], cores0)
    for i, c in enumerate(emb.tt_cores):
        assert c.grad is not None and c.grad.shape == c.shape
        assert rel_err(c.grad.cpu().numpy(), want[i]) < 5e-5, f"core {i}"
        assert np.array_equal(c.detach().cpu().numpy(), cores0[i]), "dense mode must not touch the weights"
