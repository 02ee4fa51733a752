"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ttb.h declares, validates shapes and reports errors -- no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ttb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ttb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "fbtt_embedding_b200", "lib", "libttb.so"))
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ttb.h but not exported by libttb.so"


def test_shim_binds_every_symbol_and_eleven_ops():
    from fbtt_embedding_b200 import tt_embeddings as ext

    assert sorted(ext.EXPORTED_SYMBOLS) == declared_symbols()
    for op in ["tt_forward", "tt_dense_backward", "tt_sgd_backward", "tt_adagrad_backward", "update_cache_state",
               "cache_populate", "preprocess_indices_sync", "cache_forward", "cache_backward_sgd",
               "cache_backward_dense", "cache_backward_rowwise_adagrad_approx"]:  # tt_embeddings.cpp:131-161
        assert callable(getattr(ext, op))


def test_shape_validation_and_error_string():
    from fbtt_embedding_b200 import tt_embeddings as ext

    s = ext._shape(1, 8, 64, [200, 220, 250], [4, 4, 4], [1, 32, 32, 1])
    assert list(s.L)[:3] == [55000, 250, 1]
    assert ext._lib.ttb_tt_workspace_bytes(ctypes.byref(s), 0) >= 0
    bad = ext._shape(1, 8, 62, [200, 220, 250], [4, 4, 4], [1, 32, 32, 1])  # D % 4 != 0 and != prod(q)
    arr = (ctypes.c_void_p * 4)()
    rc = ext._lib.ttb_tt_forward(ctypes.byref(bad), 1, None, None, None, arr, None, None, 0, 0, None)
    assert rc != 0 and b"D=" in ext._lib.ttb_last_error()
    with pytest.raises(RuntimeError):
        ext._shape(1, 8, 64, [200], [64], [1, 1])  # T < 2
    with pytest.raises(RuntimeError):
        ext.set_path(99)
    ext.set_path(ext.PATH_AUTO)


def test_dropin_names_import():
    import importlib
    import sys

    sys.path.insert(0, os.path.join(ROOT, "fbtt_embedding_b200", "dropin"))
    try:
        for k in ("tt_embeddings", "tt_embeddings_ops"):
            sys.modules.pop(k, None)
        ops = importlib.import_module("tt_embeddings_ops")
        ext = importlib.import_module("tt_embeddings")
        assert hasattr(ops, "TTEmbeddingBag") and hasattr(ops, "TableBatchedTTEmbeddingBag") and hasattr(ops, "OptimType")
        assert hasattr(ext, "tt_forward") and hasattr(ext, "preprocess_indices_sync")
    finally:
        sys.path.pop(0)
        for k in ("tt_embeddings", "tt_embeddings_ops"):
            sys.modules.pop(k, None)


def test_host_glue_cpu():
    """tt_matrix_to_full / suggested_tt_shapes / OptimType are pure host code: check them here."""
    import numpy as np
    import torch

    from fbtt_embedding_b200 import OptimType, suggested_tt_shapes, tt_matrix_to_full
    from oracle import tt_oracle as O

    d = np.load(os.path.join(ROOT, "tests", "golden", "tt_golden_T3.npz"))
    p, q, ranks = d["p"].tolist(), d["q"].tolist(), d["ranks"].tolist()
    W = tt_matrix_to_full(p, q, ranks, [torch.tensor(d[f"core{i}"]) for i in range(3)], [1, 0, 2, 3])
    np.testing.assert_allclose(W.numpy()[d["rows_sel"]], d["W_rows"], rtol=1.3e-6, atol=1e-5)
    np.testing.assert_allclose(W.numpy(), O.tt_matrix_to_full(p, q, ranks, [d[f"core{i}"][0] for i in range(3)]),
                               rtol=1.3e-6, atol=1e-5)
    assert str(OptimType.EXACT_ADAGRAD) == "exact_adagrad" and len(OptimType) == 9
    s = suggested_tt_shapes(1000, 3)
    assert len(s) == 3 and int(np.prod(s)) >= 1000
    assert int(np.prod(suggested_tt_shapes(64, 3, allow_round_up=False))) == 64
