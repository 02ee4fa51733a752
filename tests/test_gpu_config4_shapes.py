"""Fused heterogeneous batch at the REAL Criteo-Terabyte p-shapes of BASELINE configs[3] (26 tables, D = 128, q = [4,4,8],
ranks [64,64]; incl. the degenerate p = [1,1,3] / [1,2,2] tables and the 39.9 M-row ones), one training step against the
per-table numpy oracle: pooled rows of every table (element-wise 1e-3 + max-norm) and the fused-SGD cores of every
table.  bench.py runs the same check at the full batch before it times anything; this is the pytest-sized version."""
import numpy as np
import pytest
import torch

import bench_config4 as c4
from oracle import tt_oracle as O
from tests.helpers import elem_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_fused_criteo_tables_one_step_matches_the_oracle():
    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200 import tt_embeddings as ext
    from fbtt_embedding_b200.fused import FusedTTEmbeddingBag, pack_table_major

    ext.set_path(ext.PATH_AUTO)
    B, T = 256, len(c4.CARD)
    torch.manual_seed(3)
    fused = FusedTTEmbeddingBag(c4.CARD, c4.D, c4.RANKS, [c4.PSHAPE[E] for E in c4.CARD], c4.Q, optimizer=OptimType.SGD,
                                learning_rate=c4.LR, sparse=True, weight_dist="uniform")
    rng = np.random.RandomState(5)
    idx = [(rng.zipf(c4.ZIPF_A, size=B) % E).astype(np.int64) for E in c4.CARD]
    for t, E in enumerate(c4.CARD):
        idx[t][0] = E - 1  # the largest valid row of every table
    off1 = torch.arange(0, B + 1, dtype=torch.int64)
    pi, po = pack_table_major([torch.from_numpy(i) for i in idx], [off1] * T)
    cores0 = [[c.detach().cpu().numpy() for c in fused.table_cores(k)] for k in range(T)]
    out = fused(pi.to(DEV), po.to(DEV))  # [T, B, D]
    g = torch.rand(T, B, c4.D, device=DEV) * 0.1
    out.backward(g)
    torch.cuda.synchronize()
    got, gh = out.detach().cpu().numpy(), g.cpu().numpy()
    row, tbl0 = np.arange(B, dtype=np.int64), np.zeros(B, dtype=np.int64)
    for k, E in enumerate(c4.CARD):
        p = c4.PSHAPE[E]
        L = O.make_L(p)
        want = O.tt_forward(1, B, c4.D, p, c4.Q, c4.RANKS, L, B, idx[k], row, tbl0, cores0[k])[0]
        assert rel_err(got[k], want) < 2e-5, f"table {k} (p = {p})"
        ok, worst = elem_close(got[k], want, rtol=1e-3, atol=1e-5 * float(np.abs(want).max()))
        assert ok, f"table {k}: element-wise bound exceeded {worst:.2f}x"
        grads = O.tt_backward_dense(c4.D, p, c4.Q, c4.RANKS, L, B, idx[k], row, tbl0, gh[k][None], cores0[k])
        want_c = O.sgd_step(cores0[k], grads, c4.LR)
        for t_, (a, b) in enumerate(zip(fused.table_cores(k), want_c)):
            assert rel_err(a.detach().cpu().numpy(), b) < 1e-4, f"table {k} (p = {p}) core {t_} after fused SGD"
