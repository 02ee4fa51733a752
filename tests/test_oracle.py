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


# ---------------------------------------------------------------------------------------------
# the C restatement (oracle/hash_oracle.c) against the reference's known answers and the numpy oracle
# ---------------------------------------------------------------------------------------------
def _c_oracle():
    import ctypes
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, "oracle", "_build", "libhash_oracle.so")
    src = os.path.join(root, "oracle", "hash_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.ttb_oracle_hash.restype = ctypes.c_uint32
    lib.ttb_oracle_hash.argtypes = [ctypes.c_int64, ctypes.c_int32]
    lib.ttb_oracle_update.restype = ctypes.c_int64
    p64, p32 = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32)
    lib.ttb_oracle_update.argtypes = [p64, ctypes.c_int64, p64, p64, ctypes.c_int32]
    lib.ttb_oracle_populate_state.restype = None
    lib.ttb_oracle_populate_state.argtypes = [ctypes.c_int64, p64, p64, p32, ctypes.c_int32, p64]
    return lib, p64, p32


def test_c_oracle_hash_matches_reference_known_answers():
    lib, _, _ = _c_oracle()
    kat = json.load(open(os.path.join(G, "hash_kat.json")))["kat"]
    for key, C, want in kat:
        assert lib.ttb_oracle_hash(key, C) == want, (key, C)


def test_c_oracle_agrees_with_numpy_oracle_on_a_large_key_stream():
    lib, p64, p32 = _c_oracle()
    H, C = 1 << 14, 1 << 10
    rng = np.random.RandomState(5)
    keys = (rng.zipf(1.1, size=200_000) % 3_000_000).astype(np.int64)
    tc, fc = np.full(H, -1, np.int64), np.zeros(H, np.int64)
    sc = np.full(H, -1, np.int32)
    dropped_c = lib.ttb_oracle_update(keys.ctypes.data_as(p64), len(keys), tc.ctypes.data_as(p64),
                                      fc.ctypes.data_as(p64), H)
    tn, fn = np.full(H, -1, np.int64), np.zeros(H, np.int64)
    sn = np.full(H, -1, np.int32)
    sub = keys[:20_000]  # the Python loop is slow: compare on a prefix, then on the C result's invariants
    dropped_n = O.update_cache_state(sub, tn, fn)
    t2, f2 = np.full(H, -1, np.int64), np.zeros(H, np.int64)
    assert lib.ttb_oracle_update(sub.ctypes.data_as(p64), len(sub), t2.ctypes.data_as(p64), f2.ctypes.data_as(p64), H) == len(dropped_n)
    assert np.array_equal(t2, tn) and np.array_equal(f2, fn)
    assert fc.sum() == len(keys) - dropped_c
    sorted_c = np.empty(H, np.int64)
    lib.ttb_oracle_populate_state(C, t2.ctypes.data_as(p64), f2.ctypes.data_as(p64), sc.ctypes.data_as(p32), H,
                                  sorted_c.ctypes.data_as(p64))
    sorted_n = O.cache_populate_state(C, tn, fn, sn)
    assert np.array_equal(t2, tn) and np.array_equal(f2, fn) and np.array_equal(sc, sn)
    assert np.array_equal(sorted_c[:C], sorted_n[:C])
