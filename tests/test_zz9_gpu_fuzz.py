"""GPU fuzz of the TT forward / dense backward against the oracle (hypothesis, derandomised so a failure
reproduces): random T in {2,3,4}, odd p / q / ranks, several tables, ragged bags with empty ones -- the way the
reference's own property tests draw their cases (tt_embeddings_test.py:55-107 uses hypothesis too) -- plus random
members of the two tensor-core shape families so the bucketed kernels see shapes no hand-written case lists.
(File name sorts last: it runs after every targeted parity test.)"""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, assume, given, settings
from hypothesis import strategies as st

from oracle import tt_oracle as O
from tests.helpers import make_cores, ragged_batch, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FUZZ = settings(max_examples=25, deadline=None, derandomize=True,
                suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow,
                                       HealthCheck.filter_too_much])


@pytest.fixture(scope="module")
def ext():
    from fbtt_embedding_b200 import tt_embeddings as e

    yield e
    e.set_path(e.PATH_AUTO)


def t(x):
    return torch.as_tensor(np.ascontiguousarray(x), device=DEV)


def _generic_bwd_smem_bytes(q, ranks):
    """Shared memory of tt_bwd_generic_kernel: 4 warps x (v_0..v_{T-2} + two buffers of the widest state), as
    make_chain_dims / launch_bwd_generic compute it (csrc/ttb_api.cu, csrc/ttb_tt_generic.cu)."""
    R = [1] + list(ranks) + [1]
    m, vsum, vmax = 1, 0, 0
    for t_, qt in enumerate(q):
        m *= qt
        v = m * R[t_ + 1]
        vmax = max(vmax, v)
        if t_ < len(q) - 1:
            vsum += (v + 3) // 4 * 4
    vmax = (max(vmax, int(np.prod(q))) + 3) // 4 * 4
    return 4 * (vsum + 2 * vmax) * 4


@st.composite
def generic_shapes(draw):
    T = draw(st.integers(2, 4))
    p = [draw(st.integers(1, 9)) for _ in range(T)]
    q = [draw(st.sampled_from([1, 2, 3, 4, 5, 6, 8])) for _ in range(T)]
    if int(np.prod(q)) % 4:  # D % 4 == 0 is the reference's own requirement (tt_embeddings_cuda.cu:989)
        q[draw(st.integers(0, T - 1))] *= 4
    ranks = [draw(st.integers(1, 17)) for _ in range(T - 1)]
    assume(_generic_bwd_smem_bytes(q, ranks) <= 160 * 1024)  # the generic kernels keep the chain state in smem
    return dict(p=p, q=q, ranks=ranks)


@st.composite
def tcgen05_shapes(draw):  # T = 3, q0 = q1 = 4, r2 = 32, r1 in {4..32 step 4}, q2 in {4, 8}
    return dict(p=[draw(st.integers(1, 12)) for _ in range(3)], q=[4, 4, draw(st.sampled_from([4, 8]))],
                ranks=[4 * draw(st.integers(1, 8)), 32])


@st.composite
def warp_mma_shapes(draw):  # T = 3, q0 = 4, r1 = r2 in {16, 64}, any q1, q2 in {4, 8}
    r = draw(st.sampled_from([16, 64]))
    return dict(p=[draw(st.integers(1, 10)) for _ in range(3)], q=[4, draw(st.integers(1, 5)), draw(st.sampled_from([4, 8]))],
                ranks=[r, r])


def _check(ext, shape, num_tables, B, mean_len, seed, exact):
    p, q, ranks = shape["p"], shape["q"], shape["ranks"]
    T, E, D = len(p), int(np.prod(p)), int(np.prod(q))
    R = [1] + ranks + [1]
    rng = np.random.RandomState(seed)
    cores = make_cores(rng, num_tables, p, q, ranks)
    idx, off = ragged_batch(rng, B, E, mean_len, 2.0, num_tables)
    nnz = len(idx)
    L = t(O.make_L(p))
    e64, e32 = torch.empty(0, dtype=torch.int64, device=DEV), torch.empty(0, dtype=torch.int32, device=DEV)
    col, row, tbl, n, _ = ext.preprocess_indices_sync(t(idx), t(off), num_tables, True, e64, e32)
    assert n == nnz
    dc = [t(c) for c in cores]
    out = ext.tt_forward(1000, num_tables, B, D, p, q, R, L, nnz, col, row, tbl, dc)
    r0, t0 = O.compute_rowidx(off, num_tables)
    want = O.tt_forward(num_tables, B, D, p, q, ranks, O.make_L(p), nnz, idx, r0, t0, cores)
    assert out.shape == (num_tables, B, D)
    ftol, gtol = (2e-5, 5e-5) if exact else (1e-3, 1e-2)
    assert rel_err(out.cpu().numpy(), want) < ftol or nnz == 0, (shape, num_tables, B, nnz)
    if nnz == 0:
        assert int(out.count_nonzero()) == 0
    dout = rng.uniform(-1, 1, size=(num_tables, B, D)).astype(np.float32)
    grads = ext.tt_dense_backward(1000, D, p, q, R, L, nnz, col, row, tbl, t(dout), dc)
    g_want = O.tt_backward_dense(D, p, q, ranks, O.make_L(p), nnz, idx, r0, t0, dout, cores)
    for i in range(T):
        g = grads[i].cpu().numpy()
        assert g.shape == cores[i].shape
        if nnz == 0 or not np.abs(g_want[i]).max():
            assert not np.abs(g).max()
        else:
            assert rel_err(g, g_want[i]) < gtol, (shape, num_tables, B, nnz, i)


@FUZZ
@given(shape=generic_shapes(), num_tables=st.integers(1, 3), B=st.integers(1, 40), mean_len=st.sampled_from([0.0, 1.0, 4.0]),
       seed=st.integers(0, 10_000))
def test_fuzz_generic_path(ext, shape, num_tables, B, mean_len, seed):
    ext.set_path(ext.PATH_GENERIC)
    _check(ext, shape, num_tables, B, mean_len, seed, exact=True)


@FUZZ
@given(shape=tcgen05_shapes(), num_tables=st.integers(1, 2), B=st.integers(1, 96), mean_len=st.sampled_from([0.0, 2.0, 9.0]),
       seed=st.integers(0, 10_000))
def test_fuzz_tcgen05_family(ext, shape, num_tables, B, mean_len, seed):
    ext.set_path(ext.PATH_AUTO)
    _check(ext, shape, num_tables, B, mean_len, seed, exact=False)


@FUZZ
@given(shape=warp_mma_shapes(), num_tables=st.integers(1, 2), B=st.integers(1, 64), mean_len=st.sampled_from([0.0, 2.0, 7.0]),
       seed=st.integers(0, 10_000))
def test_fuzz_warp_mma_family(ext, shape, num_tables, B, mean_len, seed):
    ext.set_path(ext.PATH_AUTO)
    _check(ext, shape, num_tables, B, mean_len, seed, exact=False)
