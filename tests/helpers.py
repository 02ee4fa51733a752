"""Shared test helpers: synthetic inputs, the optional reference CUDA extension."""
import importlib.machinery
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "tt_embeddings.cpython-312-x86_64-linux-gnu.so")

S1 = dict(p=[200, 220, 250], q=[4, 4, 4], ranks=[32, 32], E=11_000_000, D=64)  # README / BASELINE shape

_ref_mod = None


def load_reference_extension():
    """The UNMODIFIED reference CUDA extension built for sm_100a by oracle/build_ref.sh, or None.
    Test infrastructure only (never imported by the product)."""
    global _ref_mod
    if _ref_mod is not None:
        return _ref_mod
    if not os.path.exists(REF_SO):
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)

    loader = importlib.machinery.ExtensionFileLoader("tt_embeddings", REF_SO)
    spec = importlib.util.spec_from_loader("tt_embeddings", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _ref_mod = mod
    return mod


def ragged_batch(rng, B, E, mean_len, std_len=0.0, num_tables=1):
    """Bag lengths ~ round(N(mean, std)) clipped at 0 (empty bags occur), indices uniform with
    replacement (duplicates occur) -- the generator of tt_embeddings_test.py:22-50."""
    lens = np.round(rng.normal(mean_len, std_len, B * num_tables)).astype(np.int64)
    lens = np.where(lens < 0, 0, lens)
    nnz = int(lens.sum())
    indices = rng.randint(0, E, size=nnz).astype(np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return indices, offsets


def make_cores(rng, num_tables, p, q, ranks, lo=-0.5, hi=0.5):
    R = [1] + list(ranks) + [1]
    return [rng.uniform(lo, hi, size=(num_tables, p[t], R[t] * q[t] * R[t + 1])).astype(np.float32)
            for t in range(len(p))]


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def collision_free_keys(H, n, rng, hash_fn):
    """Keys whose 3-slot probe windows (hashtbl_cuda_utils.cuh:102-133) are pairwise disjoint: the slot a key
    takes, and therefore the TT / cached partition of a batch, does not depend on which thread wins a CAS."""
    keys, used = [], set()
    cand = rng.permutation(50 * n)
    homes = hash_fn(cand.astype(np.int64), H)
    for k, h in zip(cand.tolist(), homes.tolist()):
        win = {h % H, (h + 1) % H, (h + 2) % H, (h - 1) % H, (h - 2) % H}
        if not (win & used):
            used |= {h % H, (h + 1) % H, (h + 2) % H}
            keys.append(k)
            if len(keys) == n:
                break
    return np.array(keys, dtype=np.int64)


def elem_close(a, b, rtol, atol):
    """Element-wise |a-b| <= rtol*|b| + atol (the north-star '<= rtol rel' read per element; atol covers
    entries that cancel to ~0).  Returns (ok, worst excess ratio) for assert messages."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    ratio = np.abs(a - b) / (rtol * np.abs(b) + atol)
    return bool((ratio <= 1.0).all()), float(ratio.max()) if ratio.size else 0.0
