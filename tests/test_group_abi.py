"""CPU tests of the table-group boundary (include/ttb.h `ttb_group_*`): struct layout seen by the
Python shim == the one gcc compiles from the header, argument validation and error text -- no compute calls."""
import ctypes
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_LAYOUT_C = r"""
#include <stdio.h>
#include <stddef.h>
#include "ttb.h"
int main(void) {
  printf("%zu %zu", sizeof(ttb_shape_t), sizeof(ttb_group_item_t));
#define O(f) printf(" %zu", offsetof(ttb_group_item_t, f))
  O(shape); O(nnz); O(indices); O(offsets); O(rowidx); O(tableidx); O(cores); O(grads); O(opt_state);
  O(output); O(d_output); O(workspace); O(workspace_bytes); O(plan_ready); O(reserved);
  return 0;
}
"""


def test_group_item_layout_matches_header(tmp_path):
    from fbtt_embedding_b200 import tt_embeddings as ext

    src = tmp_path / "layout.c"
    src.write_text(_LAYOUT_C)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    G = ext._GroupItem
    want = [ctypes.sizeof(ext._Shape), ctypes.sizeof(G)] + [
        getattr(G, f).offset for f in ("shape", "nnz", "indices", "offsets", "rowidx", "tableidx", "cores", "grads",
                                       "opt_state", "output", "d_output", "workspace", "workspace_bytes",
                                       "plan_ready", "reserved")]
    assert got == want


def test_group_argument_validation_without_a_gpu():
    from fbtt_embedding_b200 import tt_embeddings as ext

    lib = ext._lib
    assert ext.group_get_streams() == 1
    with pytest.raises(RuntimeError):
        ext.group_set_streams(0)
    with pytest.raises(RuntimeError):
        ext.group_set_streams(17)
    # empty group and all-empty items are defined no-ops (no CUDA call is made)
    assert lib.ttb_group_forward(0, None, None) == 0
    items = (ext._GroupItem * 2)()
    assert lib.ttb_group_preprocess(2, items, None) == 0
    assert lib.ttb_group_forward(2, items, None) == 0
    assert lib.ttb_group_backward(2, items, ext.OPTIM_SGD, 0.1, 0.0, None) == 0
    assert lib.ttb_group_forward(-1, items, None) != 0
    assert lib.ttb_group_forward(2, None, None) != 0 and b"items" in lib.ttb_last_error()
    items[1].nnz = -3
    assert lib.ttb_group_forward(2, items, None) != 0 and b"item 1" in lib.ttb_last_error()
    # a bad shape in item 1 is reported with its item number and the per-table message
    bad = ext._shape(1, 8, 62, [200, 220, 250], [4, 4, 4], [1, 32, 32, 1])  # D != prod(q)
    ctypes.memmove(ctypes.byref(items[1].shape), ctypes.byref(bad), ctypes.sizeof(ext._Shape))
    items[1].nnz = 5
    assert lib.ttb_group_forward(2, items, None) != 0
    msg = lib.ttb_last_error()
    assert b"table group item 1" in msg and b"D=" in msg
    assert lib.ttb_group_backward(2, items, 7, 0.1, 0.0, None) != 0  # unknown optimizer, same prefix
    assert b"table group item 1" in lib.ttb_last_error()


class _FakeTable:
    """Just enough of TTEmbeddingBag for GroupedLookup's host-side layout code (no CUDA here)."""

    def __init__(self, shapes):
        self.num_tables, self.use_cache, self.embedding_dim = 1, False, 16
        self.sparse, self.optimizer = True, None
        self.tt_cores = [torch.zeros(s) for s in shapes]
        self.optimizer_state = []


def test_grouped_lookup_gradient_scratch_layout():
    from fbtt_embedding_b200.grouped import GroupedLookup

    g = GroupedLookup([_FakeTable([(1, 3, 5), (1, 4, 6), (1, 2, 7)]), _FakeTable([(1, 2, 2), (1, 9, 9)])])
    n = g._grad_numel()
    assert n == 16 + 24 + 16 + 4 + 84  # every core rounded up to a multiple of 4 floats (16-byte aligned views)
    flat = torch.zeros(n)
    views = g._grad_views(flat)
    assert [tuple(v.shape) for per in views for v in per] == [(1, 3, 5), (1, 4, 6), (1, 2, 7), (1, 2, 2), (1, 9, 9)]
    for k, v in enumerate(x for per in views for x in per):
        v.fill_(k + 1)  # views are disjoint windows of the flat buffer
    assert flat.count_nonzero() == 15 + 24 + 14 + 4 + 81
    offs = [(v.data_ptr() - flat.data_ptr()) // 4 for per in views for v in per]
    assert offs == [0, 16, 40, 56, 60] and all(o % 4 == 0 for o in offs)
