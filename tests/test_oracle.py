"""CPU tests: pin the oracle (oracle/tt_oracle.py) to golden vectors produced by the
reference itself (tests/golden/make_golden.py) before anything trusts it."""
import json
import os

import numpy as np
import pytest

from oracle import tt_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def load(T):
    d = np.load(os.path.join(G, f"tt_golden_T{T}.npz"))
    p, q, ranks = d["p"].tolist(), d["q"].tolist(), d["ranks"].tolist()
    cores = [d[f"core{t}"] for t in range(T)]
    return d, p, q, ranks, cores


@pytest.mark.parametrize("T", [2, 3, 4])
def test_full_weight_matches_reference(T):
    d, p, q, ranks, cores = load(T)
    W = O.tt_matrix_to_full(p, q, ranks, [c[0] for c in cores])
    np.testing.assert_allclose(W[d["rows_sel"]], d["W_rows"], rtol=1.3e-6, atol=1e-5)
    assert abs(float(W.astype(np.float64).sum()) - float(d["W_sum"])) < 1e-3 * max(1.0, abs(float(d["W_sum"])))


@pytest.mark.parametrize("T", [2, 3, 4])
def test_chain_rows_equal_full_weight_rows(T):
    d, p, q, ranks, cores = load(T)
    L = O.make_L(p)
    idx = d["rows_sel"].astype(np.int64)
    rows = O.tt_rows(p, q, ranks, L, idx, np.zeros_like(idx), cores)
    np.testing.assert_allclose(rows, d["W_rows"], rtol=1.3e-6, atol=1e-5)


@pytest.mark.parametrize("T", [2, 3, 4])
def test_forward_backward_sgd_adagrad_match_reference_autograd(T):
    d, p, q, ranks, cores = load(T)
    L = O.make_L(p)
    idx, off = d["indices"], d["offsets"]
    B, D = d["out"].shape
    rowidx, tableidx = O.compute_rowidx(off, 1)
    out = O.tt_forward(1, B, D, p, q, ranks, L, len(idx), idx, rowidx, tableidx, cores)
    np.testing.assert_allclose(out[0], d["out"], rtol=1.3e-6, atol=1e-5)
    grads = O.tt_backward_dense(D, p, q, ranks, L, len(idx), idx, rowidx, tableidx, d["d_out"][None], cores)
    for t in range(T):
        np.testing.assert_allclose(grads[t], d[f"grad{t}"], rtol=2e-5, atol=1e-5)
    sgd = O.sgd_step(cores, grads, float(d["lr"]))
    ada_c, ada_s = O.adagrad_step(cores, [np.zeros_like(c) for c in cores], grads, float(d["lr"]), float(d["eps"]))
    for t in range(T):
        np.testing.assert_allclose(sgd[t], d[f"sgd{t}"], rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(ada_s[t], d[f"state{t}"], rtol=1e-4, atol=1e-6)
        # g/(|g|+eps) amplifies the fp32 rounding of a tiny g by ~lr/eps -> looser atol
        np.testing.assert_allclose(ada_c[t], d[f"adagrad{t}"], rtol=1e-4, atol=2e-4)


def test_hash_known_answers():
    kat = json.load(open(os.path.join(G, "hash_kat.json")))["kat"]
    for key, C, want in kat:
        assert int(O.murmur_hash_3_32_i64(np.int64(key), C)) == want, (key, C)
    keys = np.array([k for k, _, _ in kat], dtype=np.int64)
    Cs = sorted({c for _, c, _ in kat})
    for C in Cs:
        got = O.murmur_hash_3_32_i64(keys, C)
        want = {k: w for k, c, w in kat if c == C}
        assert [int(g) for g in got] == [want[int(k)] for k in keys]


def test_decompose_and_L():
    p = [200, 220, 250]
    L = O.make_L(p)
    assert L.tolist() == [55000, 250, 1]
    d = O.decompose(np.array([0, 10999999, 55000 * 3 + 250 * 7 + 9]), L)
    assert [x.tolist() for x in d] == [[0, 199, 3], [0, 219, 7], [0, 249, 9]]


def test_rowidx_with_empty_bags_and_tables():
    off = np.array([0, 2, 2, 5, 5, 6, 9])  # 2 tables x 3 bags
    r, t = O.compute_rowidx(off, 2)
    assert r.tolist() == [0, 0, 2, 2, 2, 1, 2, 2, 2]
    assert t.tolist() == [0, 0, 0, 0, 0, 1, 1, 1, 1]


def test_hashtable_insert_find_populate_partition():
    H, C = 64, 8
    hashtbl = np.full(H, -1, np.int64)
    freq = np.zeros(H, np.int64)
    state = np.full(H, -1, np.int32)
    rng = np.random.RandomState(0)
    keys = rng.zipf(1.3, 400) % 1000
    dropped = O.update_cache_state(keys, hashtbl, freq)
    # every key that was not dropped is findable, frequencies add up
    assert freq.sum() == len(keys) - len(dropped)
    for k in set(keys.tolist()) - set(dropped):
        s = O.hashtbl_find(k, hashtbl)
        assert s >= 0 and hashtbl[s] == k
    before = {int(k): int(f) for k, f in zip(hashtbl, freq) if k != -1}
    sorted_keys = O.cache_populate_state(C, hashtbl, freq, state)
    kept = {int(k) for k in hashtbl if k != -1}
    assert len(kept) == min(C, len(before))
    top = sorted(before.values(), reverse=True)[: len(kept)]
    assert sorted((before[k] for k in kept), reverse=True) == top
    for n in range(len(kept)):
        s = O.hashtbl_find(int(sorted_keys[n]), hashtbl)
        assert state[s] == n
    # lookup + CUB-style partition (cached tail reversed)
    col = np.array(list(kept)[:3] + [999999, 5], np.int64)
    offsets = np.array([0, 2, 5])
    c2, r2, t2, ntt, loc = O.preprocess_indices(col, offsets, 1, False, hashtbl, state)
    is_tt, l0 = O.cache_lookup(col, hashtbl, state)
    assert ntt == int(is_tt.sum())
    assert c2[:ntt].tolist() == col[is_tt].tolist()
    assert c2[ntt:].tolist() == col[~is_tt][::-1].tolist()
    assert loc[ntt:].tolist() == l0[~is_tt][::-1].tolist()
