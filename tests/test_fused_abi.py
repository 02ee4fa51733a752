"""CPU tests of the fused-batch boundary (include/ttb.h: ttb_het_table_t, ttb_row_map_t, ttb_tt_*_het): struct
layouts seen by the Python shim == the ones gcc compiles from the header, argument validation and error text of
the entry points -- everything that returns before a CUDA call.  No compute."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_LAYOUT_C = r"""
#include <stdio.h>
#include <stddef.h>
#include "ttb.h"
int main(void) {
  printf("%zu %zu", sizeof(ttb_het_table_t), sizeof(ttb_row_map_t));
#define H(f) printf(" %zu", offsetof(ttb_het_table_t, f))
#define M(f) printf(" %zu", offsetof(ttb_row_map_t, f))
  H(rows); H(L); H(p); H(off);
  M(world); M(rows_per_rank); M(tables_total); M(reserved); M(peer_offset); M(table_gid);
  printf(" %d", TTB_ABI_VERSION);
  return 0;
}
"""


def test_het_and_row_map_layouts_match_header(tmp_path):
    from fbtt_embedding_b200 import tt_embeddings as ext

    src = tmp_path / "layout.c"
    src.write_text(_LAYOUT_C)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    H, M = ext._HetTable, ext._RowMap
    want = [ctypes.sizeof(H), ctypes.sizeof(M)] + [getattr(H, f).offset for f in ("rows", "L", "p", "off")] + [
        getattr(M, f).offset for f in ("world", "rows_per_rank", "tables_total", "reserved", "peer_offset", "table_gid")]
    assert got[:-1] == want
    assert got[-1] == ext._lib.ttb_abi_version()


def _args(ext, B=8, n=2):
    lay = ext.HetLayout([[3, 4, 5], [2, 2, 2]][:n])
    shape = lay.cat_shape(B, 64, [4, 4, 4], [1, 32, 32, 1])
    ptr = ctypes.cast(lay.host, ctypes.c_void_p).value  # any non-NULL address: validation never dereferences it
    return lay, shape, ptr


def test_het_entry_points_validate_before_touching_the_gpu():
    from fbtt_embedding_b200 import tt_embeddings as ext

    lib = ext._lib
    lay, shape, tabs = _args(ext)
    cores = (ctypes.c_void_p * 4)()
    fwd = lambda sh, n, t, m, nnz: lib.ttb_tt_forward_het(ctypes.byref(sh), n, t, m, nnz, None, None, None, cores, None,  # noqa: E731
                                                         None, 0, 0, None)
    bwd = lambda sh, n, t, m, optim, nnz: lib.ttb_tt_backward_het(ctypes.byref(sh), n, t, m, optim, 0.1, 0.0, nnz, None,  # noqa: E731
                                                                 None, None, None, cores, cores, None, None, 0, 0, None)
    # nnz == 0 is a defined no-op (tt_embeddings_cuda.cu:983-985, 448-450)
    assert fwd(shape, lay.n_tables, tabs, None, 0) == 0
    assert bwd(shape, lay.n_tables, tabs, None, ext.OPTIM_SGD, 0) == 0
    # the concatenated shape is ONE table
    two = ext._Shape()
    ctypes.memmove(ctypes.byref(two), ctypes.byref(shape), ctypes.sizeof(ext._Shape))
    two.num_tables = 2
    assert fwd(two, lay.n_tables, tabs, None, 0) != 0 and b"num_tables must be 1" in lib.ttb_last_error()
    assert fwd(shape, 0, tabs, None, 0) != 0 and b"n_tables" in lib.ttb_last_error()
    assert fwd(shape, lay.n_tables, None, None, 0) != 0 and b"descriptors" in lib.ttb_last_error()
    assert fwd(shape, 99, tabs, None, 0) != 0 and b"slices" in lib.ttb_last_error()  # more tables than slices
    assert bwd(shape, lay.n_tables, tabs, None, 7, 5) != 0 and b"optimizer" in lib.ttb_last_error()
    assert fwd(shape, lay.n_tables, tabs, None, -1) != 0
    assert fwd(shape, lay.n_tables, tabs, None, 5) != 0 and b"NULL" in lib.ttb_last_error()  # no index arrays


def test_row_map_is_validated_against_the_batch():
    from fbtt_embedding_b200 import tt_embeddings as ext

    lib = ext._lib
    lay, shape, tabs = _args(ext, B=8)
    cores = (ctypes.c_void_p * 4)()
    off = (ctypes.c_int64 * 2)(0, 4096)
    gid = (ctypes.c_int32 * 2)(0, 3)

    def fwd(world, bw, tt, po=off, tg=gid):
        m = ext._RowMap(world, bw, tt, 0, ctypes.cast(po, ctypes.c_void_p).value if po is not None else None,
                        ctypes.cast(tg, ctypes.c_void_p).value if tg is not None else None)
        return lib.ttb_tt_forward_het(ctypes.byref(shape), lay.n_tables, tabs, ctypes.byref(m), 0, None, None, None, cores,
                                      None, None, 0, 0, None)

    assert fwd(2, 4, 5) == 0                       # 2 ranks x 4 rows == B
    assert fwd(2, 3, 5) != 0 and b"rows_per_rank" in lib.ttb_last_error()
    assert fwd(0, 8, 5) != 0
    assert fwd(2, 4, 1) != 0 and b"tables_total" in lib.ttb_last_error()   # fewer tables in all than locally
    assert fwd(2, 4, 5, po=None) != 0 and b"peer_offset" in lib.ttb_last_error()
    assert fwd(2, 4, 5, tg=None) != 0
    # the Python wrapper refuses a map without its buffer and vice versa
    with pytest.raises(RuntimeError):
        ext.tt_forward_het(lay, 8, 64, [4, 4, 4], [1, 32, 32, 1], 0, None, None, None, [], row_map=None,
                           out=ctypes)  # type: ignore[arg-type]
