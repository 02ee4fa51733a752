"""Drop-in replacement for the reference's compiled ``tt_embeddings`` extension.

The reference binds eleven C++/CUDA functions with pybind11
(tt_embeddings.cpp:131-161) and ``tt_embeddings_ops.py:14`` does
``import tt_embeddings``.  This module exposes the same eleven names with the same
positional signatures and return values, but every one of them is a thin adapter
(torch tensors -> raw device pointers) over the C-ABI CUDA library ``libttb.so``
(``include/ttb.h``, sources in ``fbtt_embedding_b200/csrc``).  Output and scratch
tensors are allocated here with torch's caching allocator and all kernels are enqueued
on ``torch.cuda.current_stream()``, so stream semantics match the reference
(tt_embeddings_cuda.cu:54-55).

There is NO CPU fallback and no PyTorch fallback: if ``libttb.so`` is missing the import
fails loudly.  Errors reported by the library surface as ``RuntimeError`` (the reference
raises the same type from ``TORCH_CHECK``).
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import List, Optional, Sequence, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("TTB_LIB", os.path.join(_HERE, "lib", "libttb.so"))
TTB_MAX_CORES = 4

OPTIM_SGD, OPTIM_ADAGRAD, OPTIM_DENSE = 0, 1, 2
PATH_AUTO, PATH_GENERIC, PATH_FAST = 0, 1, 2


class _Shape(ctypes.Structure):  # mirrors ttb_shape_t (include/ttb.h)
    _fields_ = [
        ("T", ctypes.c_int32),
        ("num_tables", ctypes.c_int32),
        ("B", ctypes.c_int32),
        ("D", ctypes.c_int32),
        ("p", ctypes.c_int32 * TTB_MAX_CORES),
        ("q", ctypes.c_int32 * TTB_MAX_CORES),
        ("R", ctypes.c_int32 * (TTB_MAX_CORES + 1)),
        ("L", ctypes.c_int64 * TTB_MAX_CORES),
    ]


class _GroupItem(ctypes.Structure):  # mirrors ttb_group_item_t (include/ttb.h)
    _fields_ = [
        ("shape", _Shape),
        ("nnz", ctypes.c_int64),
        ("indices", ctypes.c_void_p),
        ("offsets", ctypes.c_void_p),
        ("rowidx", ctypes.c_void_p),
        ("tableidx", ctypes.c_void_p),
        ("cores", ctypes.c_void_p * TTB_MAX_CORES),
        ("grads", ctypes.c_void_p * TTB_MAX_CORES),
        ("opt_state", ctypes.c_void_p * TTB_MAX_CORES),
        ("output", ctypes.c_void_p),
        ("d_output", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p),
        ("workspace_bytes", ctypes.c_size_t),
        ("plan_ready", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class _HetTable(ctypes.Structure):  # mirrors ttb_het_table_t (include/ttb.h)
    _fields_ = [
        ("rows", ctypes.c_int64),
        ("L", ctypes.c_int64 * TTB_MAX_CORES),
        ("p", ctypes.c_int32 * TTB_MAX_CORES),
        ("off", ctypes.c_int32 * TTB_MAX_CORES),
    ]


class _RowMap(ctypes.Structure):  # mirrors ttb_row_map_t (include/ttb.h)
    _fields_ = [
        ("world", ctypes.c_int32),
        ("rows_per_rank", ctypes.c_int32),
        ("tables_total", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("peer_offset", ctypes.c_void_p),
        ("table_gid", ctypes.c_void_p),
    ]


class _Batch(ctypes.Structure):  # mirrors ttb_batch_t (include/ttb.h)
    _fields_ = [
        ("nnz", ctypes.c_int64),
        ("indices", ctypes.c_void_p),
        ("rowidx", ctypes.c_void_p),
        ("tableidx", ctypes.c_void_p),
        ("offsets", ctypes.c_void_p),
        ("num_bags_total", ctypes.c_int64),
        ("cache_locations", ctypes.c_void_p),
        ("n_het_tables", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("het_tables", ctypes.c_void_p),
        ("row_map", ctypes.POINTER(_RowMap)),
    ]


BATCH_ZERO_OUTPUT = 1  # TTB_BATCH_ZERO_OUTPUT
BATCH_BF16_CORES = 2   # TTB_BATCH_BF16_CORES


def _load() -> ctypes.CDLL:
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"libttb.so not found at {_LIB_PATH}: build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or fbtt_embedding_b200/csrc/build.sh "
            "(there is no CPU / PyTorch fallback for the TT-EmbeddingBag hot path)"
        )
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i64, i32, f32, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_size_t
    sp = ctypes.POINTER(_Shape)
    pp = ctypes.POINTER(ctypes.c_void_p)
    sig = {
        "ttb_abi_version": (ctypes.c_int, []),
        "ttb_last_error": (ctypes.c_char_p, []),
        "ttb_set_path": (ctypes.c_int, [ctypes.c_int]),
        "ttb_get_path": (ctypes.c_int, []),
        "ttb_launch_count": (i64, []),
        "ttb_timing_enable": (ctypes.c_int, [ctypes.c_int]),
        "ttb_trace_set": (ctypes.c_int, [vp, vp]),
        "ttb_timing_collect": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), ctypes.c_int]),
        "ttb_tt_workspace_bytes": (sz, [sp, i64]),
        "ttb_tt_workspace_header_bytes": (sz, [sp, i64]),
        "ttb_tt_forward": (ctypes.c_int, [sp, i64, vp, vp, vp, pp, vp, vp, sz, ctypes.c_int, vp]),
        "ttb_tt_backward": (ctypes.c_int, [sp, ctypes.c_int, f32, f32, i64, vp, vp, vp, vp, pp, pp, pp, vp, sz,
                                           ctypes.c_int, vp]),
        "ttb_tt_forward_masked": (ctypes.c_int, [sp, i64, vp, vp, vp, vp, pp, vp, vp, sz, ctypes.c_int, vp]),
        "ttb_tt_backward_masked": (ctypes.c_int, [sp, ctypes.c_int, f32, f32, i64, vp, vp, vp, vp, vp, pp, pp, pp, vp, sz,
                                                  ctypes.c_int, vp]),
        "ttb_cache_frontend": (ctypes.c_int, [i64, vp, i64, i32, vp, i64, vp, vp, vp, vp, vp, vp, vp]),
        "ttb_optimizer_step": (ctypes.c_int, [sp, ctypes.c_int, f32, f32, pp, pp, pp, vp]),
        "ttb_group_set_streams": (ctypes.c_int, [ctypes.c_int]),
        "ttb_group_get_streams": (ctypes.c_int, []),
        "ttb_group_preprocess": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_GroupItem), vp]),
        "ttb_group_forward": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_GroupItem), vp]),
        "ttb_group_backward": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_GroupItem), ctypes.c_int, f32, f32, vp]),
        "ttb_het_describe": (ctypes.c_int, [i32, i32, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(_HetTable),
                                            ctypes.POINTER(ctypes.c_int32)]),
        "ttb_het_digits": (ctypes.c_int, [i32, i32, ctypes.POINTER(_HetTable), i64, i64, ctypes.POINTER(ctypes.c_int32),
                                          ctypes.POINTER(ctypes.c_int32)]),
        "ttb_row_map_offset": (ctypes.c_int, [ctypes.POINTER(_RowMap), i32, i32, i64, i64, ctypes.POINTER(ctypes.c_int64)]),
        "ttb_tt_forward_het": (ctypes.c_int, [sp, i32, vp, ctypes.POINTER(_RowMap), i64, vp, vp, vp, pp, vp, vp, sz,
                                              ctypes.c_int, vp]),
        "ttb_tt_backward_het": (ctypes.c_int, [sp, i32, vp, ctypes.POINTER(_RowMap), ctypes.c_int, f32, f32, i64, vp, vp,
                                               vp, vp, pp, pp, pp, vp, sz, ctypes.c_int, vp]),
        "ttb_tt_forward_batch": (ctypes.c_int, [sp, ctypes.POINTER(_Batch), pp, vp, vp, sz, ctypes.c_int, vp]),
        "ttb_tt_backward_batch": (ctypes.c_int, [sp, ctypes.POINTER(_Batch), ctypes.c_int, f32, f32, vp, pp, pp, pp, vp, sz,
                                                 ctypes.c_int, vp]),
        "ttb_update_cache_state": (ctypes.c_int, [i64, vp, i64, vp, vp, vp]),
        "ttb_cache_populate_temp_bytes": (sz, [i64]),
        "ttb_cache_populate": (ctypes.c_int, [sp, pp, i64, vp, vp, vp, i64, vp, vp, vp, vp, sz, vp]),
        "ttb_preprocess_rowidx": (ctypes.c_int, [i64, i64, i32, vp, vp, vp, vp]),
        "ttb_preprocess_tile_count": (i64, [i64]),
        "ttb_preprocess_cached": (ctypes.c_int, [i64, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]),
        "ttb_cache_forward": (ctypes.c_int, [i32, i64, i32, vp, vp, vp, vp, vp]),
        "ttb_cache_backward_sgd": (ctypes.c_int, [i64, i32, vp, vp, vp, f32, vp, vp]),
        "ttb_cache_backward_dense": (ctypes.c_int, [i64, i32, vp, vp, vp, vp, vp]),
        "ttb_cache_backward_rowwise_adagrad_approx": (ctypes.c_int, [i64, i32, vp, vp, vp, f32, f32, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == ABI mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.ttb_abi_version() != 8:
        raise ImportError("libttb.so ABI version mismatch")
    return lib


_lib = _load()
EXPORTED_SYMBOLS = [
    "ttb_abi_version", "ttb_last_error", "ttb_set_path", "ttb_get_path", "ttb_launch_count",
    "ttb_timing_enable", "ttb_timing_collect", "ttb_trace_set",
    "ttb_tt_workspace_bytes", "ttb_tt_workspace_header_bytes", "ttb_tt_forward", "ttb_tt_backward", "ttb_optimizer_step",
    "ttb_group_set_streams", "ttb_group_get_streams", "ttb_group_preprocess", "ttb_group_forward", "ttb_group_backward",
    "ttb_tt_forward_masked", "ttb_tt_backward_masked", "ttb_cache_frontend",
    "ttb_het_describe", "ttb_het_digits", "ttb_row_map_offset", "ttb_tt_forward_het", "ttb_tt_backward_het",
    "ttb_tt_forward_batch", "ttb_tt_backward_batch",
    "ttb_update_cache_state",
    "ttb_cache_populate_temp_bytes", "ttb_cache_populate", "ttb_preprocess_rowidx",
    "ttb_preprocess_tile_count", "ttb_preprocess_cached", "ttb_cache_forward", "ttb_cache_backward_sgd",
    "ttb_cache_backward_dense", "ttb_cache_backward_rowwise_adagrad_approx",
]


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("libttb: " + _lib.ttb_last_error().decode("utf-8", "replace"))


def set_path(path: int) -> None:
    """Select the compute path: PATH_AUTO (default), PATH_GENERIC (fp32 FFMA, exact), PATH_FAST."""
    _check(_lib.ttb_set_path(int(path)))


def get_path() -> int:
    return int(_lib.ttb_get_path())


# TTB_PATH=auto|generic|fast selects the compute path of a process from outside (the reference's own test-suite runs
# through the `import tt_embeddings` seam with TTB_PATH=generic: its tolerances are fp32 FFMA tolerances).
_env_path = os.environ.get("TTB_PATH", "").strip().lower()
if _env_path:
    if _env_path not in ("auto", "generic", "fast"):
        raise ImportError(f"TTB_PATH={_env_path!r}: expected auto, generic or fast")
    set_path({"auto": PATH_AUTO, "generic": PATH_GENERIC, "fast": PATH_FAST}[_env_path])


def launch_count() -> int:
    """Kernels launched by libttb since load (bench.py reports the delta as gpu_launches)."""
    return int(_lib.ttb_launch_count())


def group_set_streams(k: int) -> None:
    """Lanes a table group (``ttb_group_*``) spreads its items over: 1 = the caller's stream only (default)."""
    _check(_lib.ttb_group_set_streams(int(k)))


def group_get_streams() -> int:
    return int(_lib.ttb_group_get_streams())


KERNEL_KINDS = ["fwd", "bwd", "sweep", "plan", "cache"]  # TTB_KIND_* of include/ttb.h


def kernel_timing_begin() -> None:
    """Start bracketing every libttb kernel launch with CUDA events (bench.py roofline pass)."""
    _check(_lib.ttb_timing_enable(1))


def kernel_timing_end() -> dict:
    """Stop, synchronise, and return {kind: {"total_ms", "count", "mean_ms"}} per kernel class."""
    _check(_lib.ttb_timing_enable(0))
    n = len(KERNEL_KINDS)
    ms = (ctypes.c_double * n)()
    cnt = (ctypes.c_int64 * n)()
    _check(_lib.ttb_timing_collect(ms, cnt, n))
    return {k: {"total_ms": ms[i], "count": int(cnt[i]), "mean_ms": (ms[i] / cnt[i]) if cnt[i] else None}
            for i, k in enumerate(KERNEL_KINDS)}


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------
_shape_cache: dict = {}


def _shape(num_tables: int, B: int, D: int, p: Sequence[int], q: Sequence[int], ranks: Sequence[int]):
    fast_key = (num_tables, B, D, tuple(p), tuple(q), tuple(ranks))  # hot path: no per-element conversion
    s = _shape_cache.get(fast_key)
    if s is not None:
        return s
    key = (int(num_tables), int(B), int(D), tuple(int(x) for x in p), tuple(int(x) for x in q),
           tuple(int(x) for x in ranks))
    s = _shape_cache.get(key)
    if s is not None:
        _shape_cache[fast_key] = s
        return s
    T = len(key[3])
    if not (2 <= T <= TTB_MAX_CORES) or len(key[4]) != T or len(key[5]) != T + 1:
        raise RuntimeError(f"libttb: bad TT shape: p={key[3]} q={key[4]} ranks={key[5]} (need 2..4 cores, len(ranks)==T+1)")
    s = _Shape()
    s.T, s.num_tables, s.B, s.D = T, key[0], key[1], key[2]
    Lv = 1
    for t in range(T - 1, -1, -1):  # tt_embeddings_ops.py:506-512
        s.L[t] = Lv
        Lv *= key[3][t]
    for t in range(T):
        s.p[t], s.q[t] = key[3][t], key[4][t]
    for t in range(T + 1):
        s.R[t] = key[5][t]
    if len(_shape_cache) > 4096:
        _shape_cache.clear()
    _shape_cache[key] = s
    _shape_cache[fast_key] = s
    return s


_wsb_cache: dict = {}


def _workspace_bytes(shape, nnz: int) -> int:
    key = (id(shape), nnz, _lib.ttb_get_path())
    hit = _wsb_cache.get(key)
    if hit is None or hit[0] is not shape:  # the entry pins its shape object, so an id() cannot be recycled under it
        hit = (shape, int(_lib.ttb_tt_workspace_bytes(ctypes.byref(shape), nnz)))
        if len(_wsb_cache) > 4096:
            _wsb_cache.clear()
        _wsb_cache[key] = hit
    return hit[1]


_core_cache: dict = {}


def _core_ptrs(tt_cores: Sequence[torch.Tensor], what: str = "tt_cores"):
    """(ctypes pointer array, first core) for a list of core tensors.  Validation (CUDA, fp32, contiguous,
    16-byte aligned) runs once per distinct set of storages; afterwards a call costs T data_ptr() reads."""
    key = tuple([c.data_ptr() for c in tt_cores])
    arr = _core_cache.get(key)
    if arr is None:
        arr = _ptr_array(_cores_inplace(tt_cores, what))
        if len(_core_cache) > 1024:
            _core_cache.clear()
        _core_cache[key] = arr
    return arr


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * TTB_MAX_CORES)()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)  # torch.cuda.current_device() minus the lazy-init check


def _current_device() -> int:
    return _raw_device() if _raw_device is not None else torch.cuda.current_device()


def _stream() -> int:
    # torch.cuda.current_stream() builds a Stream object (~15 us); the raw accessor is ~0.3 us
    if _raw_stream is not None:
        return _raw_stream(_current_device())
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise RuntimeError(f"libttb: {what} must be a CUDA float32 tensor")
    return t if t.is_contiguous() else t.contiguous()


def _i64c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.int64 or not t.is_cuda:
        raise RuntimeError(f"libttb: {what} must be a CUDA int64 tensor")
    return t if t.is_contiguous() else t.contiguous()


def _i64c_or_pinned(t: torch.Tensor, what: str) -> torch.Tensor:
    """CSR inputs may also live in PINNED host memory: under unified addressing the plan kernel reads them over PCIe
    directly (zero-copy) -- one DMA node less than a host->device copy in front of the step, same bytes on the bus."""
    if t.dtype != torch.int64 or not (t.is_cuda or t.is_pinned()):
        raise RuntimeError(f"libttb: {what} must be an int64 tensor on the GPU or in pinned host memory")
    return t if t.is_contiguous() else t.contiguous()


def _cores_inplace(tt_cores: Sequence[torch.Tensor], what: str = "tt_cores") -> List[torch.Tensor]:
    out = []
    for c in tt_cores:
        c = c.data if isinstance(c, torch.nn.Parameter) else c
        ok_dtype = c.dtype == torch.float32 or (what == "tt_cores" and c.dtype == torch.bfloat16)
        if not ok_dtype or not c.is_cuda or not c.is_contiguous() or c.data_ptr() % 16:
            raise RuntimeError(f"libttb: {what} must be contiguous, 16-byte aligned CUDA float32 tensors"
                               + (" (or bfloat16 cores)" if what == "tt_cores" else ""))
        out.append(c)
    if out and any(c.dtype != out[0].dtype for c in out):
        raise RuntimeError(f"libttb: {what} must share one dtype")
    return out


def _core_flags(tt_cores: Sequence[torch.Tensor]) -> int:
    return BATCH_BF16_CORES if tt_cores[0].dtype == torch.bfloat16 else 0


class _DeviceGuard:
    __slots__ = ("prev", "idx")

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = _current_device()
        if self.idx is not None and cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *a):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


_tls = threading.local()


def set_keep_plans(flag: bool) -> bool:
    """Whether forwards on this thread park their bucketing plan for a backward (default True).  A custom autograd
    Function cannot see the caller's grad mode from inside ``forward`` (it always runs with grad disabled and
    ``needs_input_grad`` mirrors ``requires_grad`` even under ``torch.no_grad()``), so the MODULES set this around
    their lookup: no backward expected -> the plan goes straight back to the pool.  Returns the previous value."""
    prev = getattr(_tls, "keep_plans", True)
    _tls.keep_plans = bool(flag)
    return prev


def _keep(keep_plan: bool) -> bool:
    return bool(keep_plan) and getattr(_tls, "keep_plans", True)


_plan_cache: dict = {}    # key -> (plan buffer, indices, rowidx, tableidx, recyclable, header bytes)
_plan_free: dict = {}     # (device index, stream, header bytes) -> header-clean plan buffers whose step is over
_grad_cache: dict = {}    # (device, numels) -> (flat zero buffer, views)
_pinned: dict = {}
_PLAN_POOL = os.environ.get("TTB_PLAN_POOL", "1") != "0"
_PLAN_CACHE_BYTES = int(os.environ.get("TTB_PLAN_CACHE_BYTES", str(256 << 20)))


def _plan_key(shape, nnz: int, indices, rowidx, tableidx, stream: int, mask=None):
    return (indices.data_ptr(), indices._version, rowidx.data_ptr(), rowidx._version, tableidx.data_ptr(),
            tableidx._version, nnz, id(shape), (mask.data_ptr(), mask._version) if mask is not None else None, stream)


def _plan_retire(entry, stream: int) -> None:
    """The step that owned this plan is over (its backward ran, or the entry aged out): park the buffer for the
    next forward on the same device and stream.  Its header is zero again (include/ttb.h: the plan kernels
    leave it zero), so reuse needs neither an allocation nor a memset -- but only for a plan with the SAME
    header size: a larger header would reach into what was this plan's (dirty) body, hence the pool key.
    Buffers born inside a CUDA-graph capture belong to that graph's pool and are never handed to eager work."""
    if not (_PLAN_POOL and entry[4]):
        return
    free = _plan_free.setdefault((entry[0].device.index, stream, entry[5]), [])
    if len(free) < 8:
        free.append(entry[0])


def _plan_for(shape, nnz: int, indices: torch.Tensor, rowidx: torch.Tensor, tableidx: torch.Tensor, nbytes: int,
              build: bool, stream: int, mask: Optional[torch.Tensor] = None):
    """Bucketing-plan buffer of the tensor-core path for this exact batch -> (buffer, plan_ready, key).
    The forward builds the plan (plan_ready = 0) and parks it here; the backward of the same step finds it
    (plan_ready = 1), skips the plan kernels and retires the entry (_plan_done).  A hit requires the same
    index / row / table tensors (storage address AND in-place version counter), shape, nnz and stream; the
    entry keeps the tensors alive so the address cannot be recycled."""
    if nbytes == 0:
        return None, 0, None
    key = _plan_key(shape, nnz, indices, rowidx, tableidx, stream, mask)
    hit = _plan_cache.get(key)
    if hit is not None and not build:
        return hit[0], 1, key
    if hit is not None and hit[0].numel() >= nbytes:
        return hit[0], 0, key
    capturing = torch.cuda.is_current_stream_capturing()
    hb = int(_lib.ttb_tt_workspace_header_bytes(ctypes.byref(shape), nnz))
    plan = None
    if not capturing:
        free = _plan_free.get((indices.device.index, stream, hb))
        if free:
            for n in range(len(free) - 1, -1, -1):
                if free[n].numel() >= nbytes:
                    plan = free.pop(n)
                    break
        if plan is None and free:
            free.pop(0)  # nothing fits: let the oldest (too small) buffer go so the pool turns over
    if plan is None:
        # capacity rounded up to a power of two: batches whose nnz drifts step to step share one buffer
        plan = torch.empty(1 << max(12, (nbytes - 1).bit_length()), dtype=torch.uint8, device=indices.device)
        plan[:hb].zero_()  # header contract of include/ttb.h: zero on entry, the kernels leave it zero
    # bounded by entries AND bytes: a forward whose backward never comes (a caller that does not say keep_plan=False)
    # must not pin GPU memory without limit
    while _plan_cache and (len(_plan_cache) >= 64 or
                           sum(e[0].numel() for e in _plan_cache.values()) + plan.numel() > _PLAN_CACHE_BYTES):
        old_key = next(iter(_plan_cache))
        _plan_retire(_plan_cache.pop(old_key), old_key[-1])
    _plan_cache[key] = (plan, indices, rowidx, tableidx, not capturing, hb, mask)
    return plan, 0, key


def _plan_done(key, ok: bool) -> None:
    """After the backward (ok) or a failed call (not ok: the header may be dirty, drop the buffer)."""
    if key is None:
        return
    entry = _plan_cache.pop(key, None)
    if entry is not None and ok:
        _plan_retire(entry, key[-1])


def _grad_scratch(cores: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """Zero-filled, core-shaped gradient scratch.  ttb_tt_backward re-zeroes what it touched
    before returning (fused modes), so the same buffers are reused without a memset."""
    key = (cores[0].device.index, _stream(), tuple(c.numel() for c in cores))
    hit = _grad_cache.get(key)
    if hit is None:
        offs, total = [], 0
        for c in cores:
            offs.append(total)
            total += (c.numel() + 3) // 4 * 4
        flat = torch.zeros(total, dtype=torch.float32, device=cores[0].device)  # fp32 also for bf16 cores
        views = [flat[o:o + c.numel()].view(c.shape) for o, c in zip(offs, cores)]
        if len(_grad_cache) > 64:
            _grad_cache.clear()
        hit = (flat, views)
        _grad_cache[key] = hit
    return hit[1]


def grad_scratch(cores: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """(flat buffer, core-shaped views into it) of the zero-on-exit gradient scratch for these cores on the current
    stream.  The flat tensor is what a data-parallel step all-reduces; ``optimizer_step`` re-zeroes it."""
    _grad_scratch(cores)
    return _grad_cache[(cores[0].device.index, _stream(), tuple(c.numel() for c in cores))]


def _drop_grad_scratch() -> None:
    _grad_cache.clear()


# ------------------------------------------------------------------------------------------
# the eleven ops (tt_embeddings.cpp:131-161)
# ------------------------------------------------------------------------------------------
def _mask(cache_locations: Optional[torch.Tensor], nnz: int) -> Optional[torch.Tensor]:
    if cache_locations is None:
        return None
    m = cache_locations
    if m.dtype != torch.int32 or not m.is_cuda or not m.is_contiguous() or m.numel() < nnz:
        raise RuntimeError("libttb: cache_locations must be a contiguous CUDA int32 tensor with one entry per lookup")
    return m


def tt_forward(batch_count: int, num_tables: int, B: int, D: int, tt_p_shapes, tt_q_shapes, tt_ranks,
               L: torch.Tensor, nnz: int, indices: torch.Tensor, rowidx: torch.Tensor,
               tableidx: torch.Tensor, tt_cores: Sequence[torch.Tensor],
               cache_locations: Optional[torch.Tensor] = None, keep_plan: bool = True) -> torch.Tensor:
    """tt_embeddings_forward_cuda (tt_embeddings_cuda.cu:964-1075).  ``batch_count`` is the
    reference's chunking hint; the fused kernels have no chunks and ignore it.  ``L`` is
    implied by ``tt_p_shapes`` (tt_embeddings_ops.py:506-512) and is not read back.
    ``cache_locations`` (beyond the reference's signature, see ``cache_frontend``): only lookups with
    ``cache_locations[n] == -1`` are computed.  ``keep_plan=False``: no backward will follow this forward (inference,
    ``torch.no_grad()``), so the bucketing plan goes straight back to the pool instead of waiting in the plan cache."""
    core_arr = _core_ptrs(tt_cores)
    with _DeviceGuard(rowidx):
        out = torch.zeros((int(num_tables), int(B), int(D)), dtype=torch.float32, device=tt_cores[0].device)
        nnz = int(nnz)
        if nnz == 0:
            return out
        if int(batch_count) <= 0:
            raise RuntimeError("libttb: batch_count must be > 0")  # tt_embeddings_cuda.cu:987
        shape = _shape(num_tables, B, D, tt_p_shapes, tt_q_shapes, tt_ranks)
        indices, rowidx, tableidx = _i64c(indices, "indices"), _i64c(rowidx, "rowidx"), _i64c(tableidx, "tableidx")
        mask = _mask(cache_locations, nnz)
        wsb = _workspace_bytes(shape, nnz)
        stream = _stream()
        ws, _, key = _plan_for(shape, nnz, indices, rowidx, tableidx, wsb, True, stream, mask)
        b = _Batch()
        b.nnz, b.indices, b.rowidx, b.tableidx = nnz, indices.data_ptr(), rowidx.data_ptr(), tableidx.data_ptr()
        b.cache_locations = mask.data_ptr() if mask is not None else None
        b.flags = _core_flags(tt_cores)
        try:
            _check(_lib.ttb_tt_forward_batch(ctypes.byref(shape), ctypes.byref(b), core_arr, out.data_ptr(),
                                             ws.data_ptr() if ws is not None else None, wsb, 0, stream))
        except RuntimeError:
            _plan_done(key, False)
            raise
        if not _keep(keep_plan):
            _plan_done(key, True)
        return out


def _tt_backward(optim: int, D: int, lr: float, eps: float, p, q, ranks, nnz: int, indices, rowidx, tableidx,
                 d_output: torch.Tensor, cores: List[torch.Tensor], grads: List[torch.Tensor],
                 state: Optional[List[torch.Tensor]], cache_locations: Optional[torch.Tensor] = None) -> None:
    nnz = int(nnz)
    if nnz == 0:
        return
    d_output = _f32c(d_output, "d_output")
    num_tables = cores[0].shape[0]
    if d_output.dim() != 3 or d_output.shape[0] != num_tables or d_output.shape[2] != int(D):
        raise RuntimeError(f"libttb: d_output must be [num_tables, B, D], got {tuple(d_output.shape)}")
    shape = _shape(num_tables, d_output.shape[1], D, p, q, ranks)
    indices, rowidx, tableidx = _i64c(indices, "indices"), _i64c(rowidx, "rowidx"), _i64c(tableidx, "tableidx")
    mask = _mask(cache_locations, nnz)
    wsb = _workspace_bytes(shape, nnz)
    stream = _stream()
    ws, ready, key = _plan_for(shape, nnz, indices, rowidx, tableidx, wsb, False, stream, mask)
    b = _Batch()
    b.nnz, b.indices, b.rowidx, b.tableidx = nnz, indices.data_ptr(), rowidx.data_ptr(), tableidx.data_ptr()
    b.cache_locations = mask.data_ptr() if mask is not None else None
    b.flags = _core_flags(cores)
    try:
        _check(_lib.ttb_tt_backward_batch(ctypes.byref(shape), ctypes.byref(b), optim, float(lr), float(eps),
                                          d_output.data_ptr(), _core_ptrs(cores), _core_ptrs(grads, "gradient buffers"),
                                          _core_ptrs(state, "optimizer_state") if state is not None else None,
                                          ws.data_ptr() if ws is not None else None, wsb, ready, stream))
    except RuntimeError:
        _drop_grad_scratch()  # scratch may be dirty
        _plan_done(key, False)
        raise
    _plan_done(key, True)  # the step is over: the plan buffer goes back to the pool


def tt_dense_backward(batch_count: int, D: int, tt_p_shapes, tt_q_shapes, tt_ranks, L, nnz: int,
                      indices, rowidx, tableidx, d_output, tt_cores, cache_locations=None) -> List[torch.Tensor]:
    """tt_embeddings_backward_dense_cuda (tt_embeddings_cuda.cu:654-684): returns one dense,
    core-shaped gradient per core."""
    cores = _cores_inplace(tt_cores)
    with _DeviceGuard(d_output):
        grads = [torch.zeros_like(c, dtype=torch.float32) for c in cores]  # tt_embeddings_cuda.cu:444
        _tt_backward(OPTIM_DENSE, D, 0.0, 0.0, tt_p_shapes, tt_q_shapes, tt_ranks, nnz, indices, rowidx,
                     tableidx, d_output, cores, grads, None, cache_locations)
        return grads


def tt_dense_backward_into(D: int, tt_p_shapes, tt_q_shapes, tt_ranks, nnz: int, indices, rowidx, tableidx,
                           d_output, tt_cores, grads: Sequence[torch.Tensor]) -> None:
    """tt_dense_backward accumulating into caller-owned, core-shaped ``grads`` (zero on entry) instead of
    fresh ``zeros_like`` tensors -- the data-parallel step keeps one flat buffer and all-reduces it in place."""
    cores = _cores_inplace(tt_cores)
    with _DeviceGuard(d_output):
        _tt_backward(OPTIM_DENSE, D, 0.0, 0.0, tt_p_shapes, tt_q_shapes, tt_ranks, nnz, indices, rowidx,
                     tableidx, d_output, cores, list(grads), None)


def optimizer_step(optim: int, learning_rate: float, eps: float, num_tables: int, B: int, D: int, tt_p_shapes,
                   tt_q_shapes, tt_ranks, tt_cores, grads: Sequence[torch.Tensor],
                   optimizer_state: Optional[Sequence[torch.Tensor]]) -> None:
    """ttb_optimizer_step: the optimizer half of the fused backward (tt_embeddings_cuda.cu:392, 412-414) applied
    from dense core-shaped ``grads``, which are re-zeroed.  No counterpart among the reference's 11 ops: it is
    the epilogue of the data-parallel step (dense backward -> all-reduce -> this), SURVEY 8f-3."""
    cores = list(tt_cores)
    if cores[0].dtype != torch.float32:
        raise RuntimeError("libttb: optimizer_step takes fp32 cores (the data-parallel replica step keeps fp32 masters)")
    shape = _shape(num_tables, B, D, tt_p_shapes, tt_q_shapes, tt_ranks)
    state = list(optimizer_state) if optimizer_state is not None else None
    if state is not None:
        for c, s_ in zip(cores, state):
            if s_.shape != c.shape:
                raise RuntimeError("libttb: optimizer_state must have the shape of its core")
    for c, g in zip(cores, grads):
        if g.shape != c.shape:
            raise RuntimeError("libttb: gradient buffers must have the shape of their core")
    with _DeviceGuard(cores[0]):
        _check(_lib.ttb_optimizer_step(ctypes.byref(shape), int(optim), float(learning_rate), float(eps),
                                       _core_ptrs(cores), _core_ptrs(list(grads), "gradient buffers"),
                                       _core_ptrs(state, "optimizer_state") if state is not None else None,
                                       _stream()))


def tt_sgd_backward(batch_count: int, D: int, learning_rate: float, tt_p_shapes, tt_q_shapes, tt_ranks, L,
                    nnz: int, indices, rowidx, tableidx, d_output, tt_cores, cache_locations=None) -> None:
    """tt_embeddings_backward_sgd_cuda (tt_embeddings_cuda.cu:686-717): fused w -= lr * g."""
    cores = list(tt_cores)
    _core_ptrs(cores)  # validates on first sight
    with _DeviceGuard(d_output):
        _tt_backward(OPTIM_SGD, D, learning_rate, 0.0, tt_p_shapes, tt_q_shapes, tt_ranks, nnz, indices,
                     rowidx, tableidx, d_output, cores, _grad_scratch(cores), None, cache_locations)


def tt_adagrad_backward(batch_count: int, D: int, learning_rate: float, eps: float, tt_p_shapes, tt_q_shapes,
                        tt_ranks, L, nnz: int, indices, rowidx, tableidx, d_output, optimizer_state,
                        tt_cores, cache_locations=None) -> None:
    """tt_embeddings_backward_adagrad_cuda (tt_embeddings_cuda.cu:719-752): fused
    state += g*g; w -= lr * g / (sqrt(state) + eps)."""
    cores = list(tt_cores)
    state = list(optimizer_state)
    _core_ptrs(cores)
    _core_ptrs(state, "optimizer_state")
    for c, s in zip(cores, state):
        if s.shape != c.shape:
            raise RuntimeError("libttb: optimizer_state must have the shape of its core")
    with _DeviceGuard(d_output):
        _tt_backward(OPTIM_ADAGRAD, D, learning_rate, eps, tt_p_shapes, tt_q_shapes, tt_ranks, nnz, indices,
                     rowidx, tableidx, d_output, cores, _grad_scratch(cores), state, cache_locations)


# ------------------------------------------------------------------------------------------
# CSR batches (ttb_tt_forward_batch / ttb_tt_backward_batch): the CSR -> COO step happens inside the plan kernel
# ------------------------------------------------------------------------------------------
def csr_supported(num_tables: int, B: int, D: int, tt_p_shapes, tt_q_shapes, tt_ranks, nnz: int) -> bool:
    """True when the bucketed kernels take this shape on the current path: a lookup can then go straight from
    (indices, offsets) to pooled rows -- plan, forward, backward(+optimizer) = 3 launches per training step."""
    return int(nnz) > 0 and _workspace_bytes(_shape(num_tables, B, D, tt_p_shapes, tt_q_shapes, tt_ranks), int(nnz)) > 0


def _csr_batch(nnz: int, indices: torch.Tensor, offsets: torch.Tensor) -> _Batch:
    b = _Batch()
    b.nnz = nnz
    b.indices = indices.data_ptr()
    b.offsets = offsets.data_ptr()
    b.num_bags_total = offsets.numel() - 1
    return b


def _csr_plan_key(shape, nnz: int, indices: torch.Tensor, offsets: torch.Tensor, stream: int):
    return ("csr", indices.data_ptr(), indices._version, offsets.data_ptr(), offsets._version, nnz, id(shape), stream)


def _plan_for_key(key, shape, nnz: int, keepalive: tuple, nbytes: int, build: bool, stream: int, device):
    """_plan_for for an arbitrary key: (buffer, plan_ready).  `keepalive` pins the tensors the key's addresses name."""
    hit = _plan_cache.get(key)
    if hit is not None and not build:
        return hit[0], 1
    if hit is not None and hit[0].numel() >= nbytes:
        return hit[0], 0
    capturing = torch.cuda.is_current_stream_capturing()
    hb = int(_lib.ttb_tt_workspace_header_bytes(ctypes.byref(shape), nnz))
    plan = None
    if not capturing:
        free = _plan_free.get((device.index, stream, hb))
        if free:
            for n in range(len(free) - 1, -1, -1):
                if free[n].numel() >= nbytes:
                    plan = free.pop(n)
                    break
        if plan is None and free:
            free.pop(0)
    if plan is None:
        plan = torch.empty(1 << max(12, (nbytes - 1).bit_length()), dtype=torch.uint8, device=device)
        plan[:hb].zero_()
    while _plan_cache and (len(_plan_cache) >= 64 or
                           sum(e[0].numel() for e in _plan_cache.values()) + plan.numel() > _PLAN_CACHE_BYTES):
        old_key = next(iter(_plan_cache))
        _plan_retire(_plan_cache.pop(old_key), old_key[-1])
    _plan_cache[key] = (plan, keepalive, None, None, not capturing, hb, None)
    return plan, 0


def tt_forward_csr(num_tables: int, B: int, D: int, tt_p_shapes, tt_q_shapes, tt_ranks, indices: torch.Tensor,
                   offsets: torch.Tensor, tt_cores: Sequence[torch.Tensor], keep_plan: bool = True) -> torch.Tensor:
    """tt_forward from the CSR pair (``indices`` int64 [nnz], ``offsets`` int64 [num_tables * B + 1]) -- no
    preprocess launch: the plan kernel of the bucketed path derives each lookup's bag itself
    (compute_rowidx_kernel, tt_embeddings_cuda.cu:1338-1354, folded in).  Only for shapes / paths where
    ``csr_supported`` is True.  Returns ``[num_tables, B, D]``."""
    core_arr = _core_ptrs(tt_cores)
    with _DeviceGuard(tt_cores[0]):
        nnz = indices.numel()
        if nnz == 0:
            return torch.zeros((int(num_tables), int(B), int(D)), dtype=torch.float32, device=tt_cores[0].device)
        # uninitialised: the plan kernel zero-fills it on its way (TTB_BATCH_ZERO_OUTPUT), no memset launch
        out = torch.empty((int(num_tables), int(B), int(D)), dtype=torch.float32, device=tt_cores[0].device)
        shape = _shape(num_tables, B, D, tt_p_shapes, tt_q_shapes, tt_ranks)
        indices, offsets = _i64c_or_pinned(indices, "indices"), _i64c_or_pinned(offsets, "offsets")
        wsb = _workspace_bytes(shape, nnz)
        if wsb == 0:
            raise RuntimeError("libttb: tt_forward_csr needs a shape / path the bucketed kernels cover (csr_supported)")
        stream = _stream()
        key = _csr_plan_key(shape, nnz, indices, offsets, stream)
        ws, _ = _plan_for_key(key, shape, nnz, (indices, offsets), wsb, True, stream, tt_cores[0].device)
        b = _csr_batch(nnz, indices, offsets)
        b.flags = BATCH_ZERO_OUTPUT | _core_flags(tt_cores)
        try:
            _check(_lib.ttb_tt_forward_batch(ctypes.byref(shape), ctypes.byref(b), core_arr, out.data_ptr(), ws.data_ptr(),
                                             wsb, 0, stream))
        except RuntimeError:
            _plan_done(key, False)
            raise
        if not _keep(keep_plan):
            _plan_done(key, True)
        return out


def tt_backward_csr(optim: int, D: int, learning_rate: float, eps: float, tt_p_shapes, tt_q_shapes, tt_ranks,
                    indices: torch.Tensor, offsets: torch.Tensor, d_output: torch.Tensor, tt_cores,
                    optimizer_state: Optional[Sequence[torch.Tensor]] = None) -> Optional[List[torch.Tensor]]:
    """The three backward ops on a CSR batch.  ``OPTIM_DENSE`` returns the core-shaped gradients; ``OPTIM_SGD`` /
    ``OPTIM_ADAGRAD`` update ``tt_cores`` (and ``optimizer_state``) in place and return None -- on the tcgen05 path
    the optimizer runs inside the backward kernel (slices of core 1 straight from the accumulator)."""
    cores = _cores_inplace(list(tt_cores))
    with _DeviceGuard(d_output):
        dense = int(optim) == OPTIM_DENSE
        grads = [torch.zeros_like(c, dtype=torch.float32) for c in cores] if dense else _grad_scratch(cores)
        nnz = indices.numel()
        if nnz == 0:
            return grads if dense else None
        d_output = _f32c(d_output, "d_output")
        num_tables = cores[0].shape[0]
        if d_output.dim() != 3 or d_output.shape[0] != num_tables or d_output.shape[2] != int(D):
            raise RuntimeError(f"libttb: d_output must be [num_tables, B, D], got {tuple(d_output.shape)}")
        state = None
        if int(optim) == OPTIM_ADAGRAD:
            state = list(optimizer_state) if optimizer_state is not None else []
            if len(state) != len(cores) or any(s_.shape != c.shape for c, s_ in zip(cores, state)):
                raise RuntimeError("libttb: optimizer_state must have the shape of its core")
        shape = _shape(num_tables, d_output.shape[1], D, tt_p_shapes, tt_q_shapes, tt_ranks)
        indices, offsets = _i64c_or_pinned(indices, "indices"), _i64c_or_pinned(offsets, "offsets")
        wsb = _workspace_bytes(shape, nnz)
        if wsb == 0:
            raise RuntimeError("libttb: tt_backward_csr needs a shape / path the bucketed kernels cover (csr_supported)")
        stream = _stream()
        key = _csr_plan_key(shape, nnz, indices, offsets, stream)
        ws, ready = _plan_for_key(key, shape, nnz, (indices, offsets), wsb, False, stream, d_output.device)
        b = _csr_batch(nnz, indices, offsets)
        b.flags = _core_flags(cores)
        try:
            _check(_lib.ttb_tt_backward_batch(ctypes.byref(shape), ctypes.byref(b), int(optim), float(learning_rate),
                                              float(eps), d_output.data_ptr(), _core_ptrs(cores),
                                              _core_ptrs(grads, "gradient buffers"),
                                              _core_ptrs(state, "optimizer_state") if state is not None else None,
                                              ws.data_ptr(), wsb, ready, stream))
        except RuntimeError:
            _drop_grad_scratch()
            _plan_done(key, False)
            raise
        _plan_done(key, True)
        return grads if dense else None


# ------------------------------------------------------------------------------------------
# fused heterogeneous table batch (include/ttb.h: ttb_het_describe / ttb_tt_forward_het / ttb_tt_backward_het)
# ------------------------------------------------------------------------------------------
class HetLayout:
    """Where each table of a fused heterogeneous batch lives in the concatenated cores.

    ``p_shapes[k]`` is table k's p-shape; all tables share q-shapes and ranks.  ``P[t]`` is the slice count of
    concatenated core t, ``off[k][t]`` table k's first slice there, ``rows[k] = prod(p_shapes[k])``.  The
    descriptors the kernels read (``ttb_het_table_t``) are filled by the library on the host (``host``) and
    uploaded once per device (``device_table``)."""

    def __init__(self, p_shapes: Sequence[Sequence[int]]) -> None:
        self.p_shapes = [[int(v) for v in p] for p in p_shapes]
        self.n_tables = len(self.p_shapes)
        if self.n_tables == 0:
            raise RuntimeError("libttb: a heterogeneous batch needs at least one table")
        self.T = len(self.p_shapes[0])
        if any(len(p) != self.T for p in self.p_shapes):
            raise RuntimeError("libttb: every table of a heterogeneous batch needs the same number of TT cores")
        flat = (ctypes.c_int32 * (self.n_tables * self.T))(*[v for p in self.p_shapes for v in p])
        self.host = (_HetTable * self.n_tables)()
        P = (ctypes.c_int32 * TTB_MAX_CORES)()
        _check(_lib.ttb_het_describe(self.T, self.n_tables, flat, self.host, P))
        self.P = [int(P[t]) for t in range(self.T)]
        self.off = [[int(h.off[t]) for t in range(self.T)] for h in self.host]
        self.rows = [int(h.rows) for h in self.host]
        self._dev: dict = {}
        self._shapes: dict = {}

    def digits(self, table: int, index: int) -> Optional[List[int]]:
        """Concatenated slice numbers of ``index`` of ``table`` as the kernels compute them (host evaluation of the
        same function), or None when the lookup is out of range."""
        dg = (ctypes.c_int32 * TTB_MAX_CORES)()
        ok = ctypes.c_int32(0)
        _check(_lib.ttb_het_digits(self.T, self.n_tables, self.host, int(table), int(index), dg, ctypes.byref(ok)))
        return [int(dg[t]) for t in range(self.T)] if ok.value else None

    def device_table(self, device: torch.device) -> torch.Tensor:
        t = self._dev.get(device)
        if t is None:
            raw = bytes(memoryview(self.host).cast("B"))
            t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
            self._dev[device] = t
        return t

    def cat_shape(self, B: int, D: int, q: Sequence[int], ranks: Sequence[int]):
        """The concatenated shape (one table, P_t slices per core) as a ttb_shape_t OWNED by this layout: its
        identity keys the plan / workspace caches, so a fused batch never shares a plan with an ordinary table of
        the same concatenated shape."""
        key = (int(B), int(D), tuple(int(v) for v in q), tuple(int(v) for v in ranks))
        s = self._shapes.get(key)
        if s is None:
            src = _shape(1, key[0], key[1], self.P, key[2], key[3])
            s = _Shape()
            ctypes.memmove(ctypes.byref(s), ctypes.byref(src), ctypes.sizeof(_Shape))
            if len(self._shapes) > 64:
                self._shapes.clear()
            self._shapes[key] = s
        return s


class RowMap:
    """ttb_row_map_t: where the pooled rows of a table-parallel rank go (fused exchange, include/ttb.h).

    ``peer_offset[w]`` = (rank w's batch-slice buffer - the local one) in floats, ``table_gid[k]`` = global number
    of local table k; the batch of ``world * rows_per_rank`` bags is split into ``world`` slices.  Holds the two
    device arrays the kernels read and a host copy for ``offset()`` (host evaluation of the same function)."""

    def __init__(self, world: int, rows_per_rank: int, tables_total: int, peer_offset: Sequence[int],
                 table_gid: Sequence[int], device) -> None:
        if len(peer_offset) != int(world):
            raise RuntimeError("libttb: row map needs one peer offset per rank")
        if any(not 0 <= int(g) < int(tables_total) for g in table_gid):
            raise RuntimeError("libttb: row map: table_gid out of range")
        self.world, self.rows_per_rank, self.tables_total = int(world), int(rows_per_rank), int(tables_total)
        self._host_off = (ctypes.c_int64 * len(peer_offset))(*[int(v) for v in peer_offset])
        self._host_gid = (ctypes.c_int32 * max(len(table_gid), 1))(*[int(v) for v in table_gid])
        self.peer_offset = torch.tensor([int(v) for v in peer_offset], dtype=torch.int64).to(device)
        self.table_gid = torch.tensor([int(v) for v in table_gid], dtype=torch.int32).to(device)
        self.n_tables = len(table_gid)
        self.c = _RowMap(self.world, self.rows_per_rank, self.tables_total, 0, self.peer_offset.data_ptr(),
                         self.table_gid.data_ptr())

    def offset(self, D: int, table: int, row: int) -> int:
        """Element offset of (table, row)'s pooled row relative to the local buffer, as the kernels compute it."""
        host = _RowMap(self.world, self.rows_per_rank, self.tables_total, 0,
                       ctypes.cast(self._host_off, ctypes.c_void_p).value,
                       ctypes.cast(self._host_gid, ctypes.c_void_p).value)
        out = ctypes.c_int64(0)
        _check(_lib.ttb_row_map_offset(ctypes.byref(host), self.world * self.rows_per_rank, int(D), int(table), int(row),
                                       ctypes.byref(out)))
        return int(out.value)


def tt_forward_het(layout: HetLayout, B: int, D: int, tt_q_shapes, tt_ranks, nnz: int, indices: torch.Tensor,
                   rowidx: torch.Tensor, tableidx: torch.Tensor, tt_cores: Sequence[torch.Tensor],
                   row_map: Optional[RowMap] = None, out: Optional[torch.Tensor] = None,
                   keep_plan: bool = True) -> torch.Tensor:
    """ttb_tt_forward_het: ``tt_forward`` for ``layout.n_tables`` differently-sized tables whose cores are
    concatenated along the slice dimension (``tt_cores[t]`` is ``[1, layout.P[t], S_t]``) -- one plan + one
    forward launch for all of them.  Returns ``[n_tables, B, D]``.

    With ``row_map`` (fused exchange) the pooled rows are ADDED into the ranks' batch-slice buffers instead:
    ``out`` is then the caller's local buffer ``[rows_per_rank, tables_total, D]`` (the peers' buffers lie at
    ``row_map.peer_offset``), zero-filled and rank-synchronised by the caller; it is returned as is."""
    core_arr = _core_ptrs(tt_cores)
    if (row_map is None) != (out is None):
        raise RuntimeError("libttb: row_map and out go together (the fused exchange writes into caller-owned buffers)")
    if row_map is not None:
        if (row_map.n_tables != layout.n_tables or row_map.world * row_map.rows_per_rank != int(B)
                or tuple(out.shape) != (row_map.rows_per_rank, row_map.tables_total, int(D))
                or out.dtype != torch.float32 or not out.is_contiguous() or out.data_ptr() % 16):
            raise RuntimeError("libttb: row map does not match the layout / batch / output buffer")
    for t, c in enumerate(tt_cores):
        if c.shape[0] != 1 or c.shape[1] != layout.P[t]:
            raise RuntimeError(f"libttb: concatenated core {t} must be [1, {layout.P[t]}, S], got {tuple(c.shape)}")
    with _DeviceGuard(rowidx):
        dev = tt_cores[0].device
        if out is None:
            out = torch.zeros((layout.n_tables, int(B), int(D)), dtype=torch.float32, device=dev)
        nnz = int(nnz)
        if nnz == 0:
            return out
        shape = layout.cat_shape(B, D, tt_q_shapes, tt_ranks)
        indices, rowidx, tableidx = _i64c(indices, "indices"), _i64c(rowidx, "rowidx"), _i64c(tableidx, "tableidx")
        wsb = _workspace_bytes(shape, nnz)
        stream = _stream()
        # a plan's records carry output-row offsets: one built under a row map is only good for that map
        salt = row_map.peer_offset if row_map is not None else None
        ws, _, key = _plan_for(shape, nnz, indices, rowidx, tableidx, wsb, True, stream, salt)
        try:
            _check(_lib.ttb_tt_forward_het(ctypes.byref(shape), layout.n_tables, layout.device_table(dev).data_ptr(),
                                           ctypes.byref(row_map.c) if row_map is not None else None,
                                           nnz, indices.data_ptr(), rowidx.data_ptr(), tableidx.data_ptr(), core_arr,
                                           out.data_ptr(), ws.data_ptr() if ws is not None else None, wsb, 0, stream))
        except RuntimeError:
            _plan_done(key, False)
            raise
        if not _keep(keep_plan):
            _plan_done(key, True)
        return out


def tt_backward_het(layout: HetLayout, optim: int, D: int, learning_rate: float, eps: float, tt_q_shapes, tt_ranks,
                    nnz: int, indices, rowidx, tableidx, d_output: torch.Tensor, tt_cores,
                    optimizer_state: Optional[Sequence[torch.Tensor]] = None,
                    row_map: Optional[RowMap] = None) -> Optional[List[torch.Tensor]]:
    """ttb_tt_backward_het.  ``OPTIM_DENSE`` returns the core-shaped gradients of the concatenated cores;
    ``OPTIM_SGD`` / ``OPTIM_ADAGRAD`` apply the fused update in place (tt_embeddings_cuda.cu:686-752 semantics)
    and return None.  With ``row_map`` (fused exchange) ``d_output`` is the LOCAL gradient buffer
    ``[rows_per_rank, tables_total, D]``; the rows of the other batch slices are read from the peers' buffers."""
    cores = _cores_inplace(list(tt_cores))
    nnz = int(nnz)
    with _DeviceGuard(d_output):
        dense = int(optim) == OPTIM_DENSE
        grads = [torch.zeros_like(c) for c in cores] if dense else _grad_scratch(cores)
        if nnz == 0:
            return grads if dense else None
        d_output = _f32c(d_output, "d_output")
        if row_map is not None:
            if (tuple(d_output.shape) != (row_map.rows_per_rank, row_map.tables_total, int(D))
                    or row_map.n_tables != layout.n_tables or d_output.data_ptr() % 16):
                raise RuntimeError("libttb: row map does not match the layout / gradient buffer")
            B = row_map.world * row_map.rows_per_rank
        elif d_output.dim() != 3 or d_output.shape[0] != layout.n_tables or d_output.shape[2] != int(D):
            raise RuntimeError(f"libttb: d_output must be [{layout.n_tables}, B, {int(D)}], got {tuple(d_output.shape)}")
        else:
            B = d_output.shape[1]
        state = None
        if int(optim) == OPTIM_ADAGRAD:
            state = list(optimizer_state) if optimizer_state is not None else []
            if len(state) != len(cores) or any(s_.shape != c.shape for c, s_ in zip(cores, state)):
                raise RuntimeError("libttb: optimizer_state must have the shape of its core")
        shape = layout.cat_shape(B, D, tt_q_shapes, tt_ranks)
        indices, rowidx, tableidx = _i64c(indices, "indices"), _i64c(rowidx, "rowidx"), _i64c(tableidx, "tableidx")
        wsb = _workspace_bytes(shape, nnz)
        stream = _stream()
        salt = row_map.peer_offset if row_map is not None else None
        ws, ready, key = _plan_for(shape, nnz, indices, rowidx, tableidx, wsb, False, stream, salt)
        try:
            _check(_lib.ttb_tt_backward_het(ctypes.byref(shape), layout.n_tables,
                                            layout.device_table(d_output.device).data_ptr(),
                                            ctypes.byref(row_map.c) if row_map is not None else None, int(optim),
                                            float(learning_rate), float(eps), nnz, indices.data_ptr(),
                                            rowidx.data_ptr(), tableidx.data_ptr(), d_output.data_ptr(),
                                            _core_ptrs(cores), _core_ptrs(grads, "gradient buffers"),
                                            _core_ptrs(state, "optimizer_state") if state is not None else None,
                                            ws.data_ptr() if ws is not None else None, wsb, ready, stream))
        except RuntimeError:
            _drop_grad_scratch()
            _plan_done(key, False)
            raise
        _plan_done(key, True)
        return grads if dense else None


def update_cache_state(indices: torch.Tensor, hashtbl: torch.Tensor, cache_freq: torch.Tensor) -> None:
    """update_cache_state_cuda (tt_embeddings_cuda.cu:1091-1113)."""
    nnz = indices.numel()
    if nnz == 0:
        return
    if hashtbl.numel() == 0 or hashtbl.numel() != cache_freq.numel():
        raise RuntimeError("libttb: hashtbl must be non-empty and match cache_freq")  # :1099-1100
    with _DeviceGuard(indices):
        indices = _i64c(indices, "indices")
        _check(_lib.ttb_update_cache_state(nnz, indices.data_ptr(), hashtbl.numel(), hashtbl.data_ptr(),
                                           cache_freq.data_ptr(), _stream()))


def cache_frontend(colidx: torch.Tensor, offsets: torch.Tensor, num_tables: int, hashtbl: torch.Tensor,
                   cache_freq: torch.Tensor, cache_state: torch.Tensor
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """ttb_cache_frontend: ``update_cache_state`` + ``preprocess_indices_sync`` of a populated cache
    in one launch, without the device->host round trip of the latter (tt_embeddings_cuda.cu:1481-1488).
    Returns ``(colidx, rowidx, tableidx, cache_locations)`` in BATCH order: ``cache_locations[n]`` is the cache row
    of lookup n or -1; pass it to ``tt_forward`` / ``tt_*_backward`` as ``cache_locations=`` (they take the -1
    entries) and to ``cache_forward`` / ``cache_backward_*`` with the full ``nnz`` (they take the rest)."""
    with _DeviceGuard(colidx):
        colidx, offsets = _i64c(colidx, "colidx"), _i64c(offsets, "offsets")
        nnz = colidx.numel()
        rowidx = torch.empty_like(colidx)
        tableidx = torch.empty_like(colidx)
        loc = torch.empty(nnz, dtype=torch.int32, device=colidx.device)
        if nnz == 0:
            return colidx, rowidx, tableidx, loc
        if hashtbl.numel() == 0 or hashtbl.numel() != cache_freq.numel() or hashtbl.numel() != cache_state.numel():
            raise RuntimeError("libttb: hashtbl, cache_freq and cache_state must be non-empty and equally long")
        num_bags = offsets.numel() - 1
        _check(_lib.ttb_cache_frontend(nnz, colidx.data_ptr(), num_bags, num_bags // int(num_tables),
                                       offsets.data_ptr(), hashtbl.numel(), hashtbl.data_ptr(),
                                       cache_freq.data_ptr(), cache_state.data_ptr(), rowidx.data_ptr(),
                                       tableidx.data_ptr(), loc.data_ptr(), _stream()))
        return colidx, rowidx, tableidx, loc


def cache_populate(num_embeddings: int, tt_p_shapes, tt_q_shapes, tt_ranks, tt_cores, L, hashtbl, cache_freq,
                   cache_state, cache_weight) -> None:
    """cache_populate_cuda (tt_embeddings_cuda.cu:1260-1336)."""
    cores = _cores_inplace(list(tt_cores))
    if cores[0].dtype == torch.bfloat16:  # rows are materialised by the exact fp32 chain from the bf16 values
        cores = [c.float() for c in cores]
    cw = cache_weight.data if isinstance(cache_weight, torch.nn.Parameter) else cache_weight
    H, C, D = hashtbl.numel(), cw.shape[0], cw.shape[1]
    if H == 0 or H != cache_freq.numel() or H < C:
        raise RuntimeError("libttb: cache_populate: bad hashtbl / cache_freq / cache_weight sizes")  # :1271-1274
    with _DeviceGuard(cw):
        shape = _shape(1, max(C, 1), D, tt_p_shapes, tt_q_shapes, tt_ranks)
        sorted_keys = torch.empty_like(hashtbl)
        sorted_freq = torch.empty_like(cache_freq)
        tb = _lib.ttb_cache_populate_temp_bytes(H)
        temp = torch.empty(tb, dtype=torch.uint8, device=cw.device)
        _check(_lib.ttb_cache_populate(ctypes.byref(shape), _ptr_array(cores), H, hashtbl.data_ptr(),
                                       cache_freq.data_ptr(), cache_state.data_ptr(), C, cw.data_ptr(),
                                       sorted_keys.data_ptr(), sorted_freq.data_ptr(), temp.data_ptr(), tb,
                                       _stream()))


def preprocess_indices_sync(colidx: torch.Tensor, offsets: torch.Tensor, num_tables: int, warmup: bool,
                            hashtbl: torch.Tensor, cache_state: torch.Tensor
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int, Optional[torch.Tensor]]:
    """preprocess_indices_sync_cuda (tt_embeddings_cuda.cu:1377-1496)."""
    with _DeviceGuard(colidx):
        colidx, offsets = _i64c(colidx, "colidx"), _i64c(offsets, "offsets")
        rowidx = torch.empty_like(colidx)
        tableidx = torch.empty_like(colidx)
        nnz = colidx.numel()
        if nnz == 0:
            return colidx, rowidx, tableidx, 0, None
        num_bags = offsets.numel() - 1
        B = num_bags // int(num_tables)
        _check(_lib.ttb_preprocess_rowidx(nnz, num_bags, B, offsets.data_ptr(), rowidx.data_ptr(),
                                          tableidx.data_ptr(), _stream()))
        if warmup or int(num_tables) != 1:
            return colidx, rowidx, tableidx, nnz, None
        out_col = torch.empty_like(colidx)
        out_row = torch.empty_like(rowidx)
        out_loc = torch.empty(nnz, dtype=torch.int32, device=colidx.device)
        scratch = torch.empty(_lib.ttb_preprocess_tile_count(nnz), dtype=torch.int32, device=colidx.device)
        pin_key = ("num_tt", colidx.device.index, threading.get_ident())  # one per device and host thread
        host = _pinned.get(pin_key)
        if host is None:
            host = torch.zeros(1, dtype=torch.int32).pin_memory()
            _pinned[pin_key] = host
        _check(_lib.ttb_preprocess_cached(nnz, colidx.data_ptr(), rowidx.data_ptr(), hashtbl.numel(),
                                          hashtbl.data_ptr(), cache_state.data_ptr(), out_col.data_ptr(),
                                          out_row.data_ptr(), out_loc.data_ptr(), scratch.data_ptr(),
                                          host.data_ptr(), _stream()))
        return out_col, out_row, tableidx, int(host[0]), out_loc


def cache_forward(B: int, nnz: int, cache_locations: torch.Tensor, rowidx: torch.Tensor,
                  cache_weight: torch.Tensor, output: torch.Tensor) -> None:
    """cache_forward_cuda (tt_embeddings_cuda.cu:1540-1572); ``output`` updated in place."""
    D = cache_weight.shape[1]
    with _DeviceGuard(rowidx):
        if not output.is_contiguous():
            raise RuntimeError("libttb: output must be contiguous")
        loc = cache_locations if cache_locations.is_contiguous() else cache_locations.contiguous()
        row = _i64c(rowidx, "rowidx")
        _check(_lib.ttb_cache_forward(int(B), int(nnz), D, loc.data_ptr(), row.data_ptr(),
                                      cache_weight.data_ptr(), output.data_ptr(), _stream()))


def cache_backward_sgd(nnz: int, grad_output: torch.Tensor, cache_locations, rowidx, learning_rate: float,
                       cache_weight: torch.Tensor) -> None:
    """cache_backward_sgd_cuda (tt_embeddings_cuda.cu:1623-1657)."""
    if int(nnz) == 0:
        return
    cw = cache_weight.data if isinstance(cache_weight, torch.nn.Parameter) else cache_weight
    with _DeviceGuard(cw):
        go = _f32c(grad_output, "grad_output")
        loc = cache_locations if cache_locations.is_contiguous() else cache_locations.contiguous()
        _check(_lib.ttb_cache_backward_sgd(int(nnz), cw.shape[1], go.data_ptr(), loc.data_ptr(),
                                           _i64c(rowidx, "rowidx").data_ptr(), float(learning_rate),
                                           cw.data_ptr(), _stream()))


def cache_backward_dense(nnz: int, grad_output: torch.Tensor, cache_locations, rowidx, learning_rate: float,
                         cache_weight: torch.Tensor) -> torch.Tensor:
    """cache_backward_dense_cuda (tt_embeddings_cuda.cu:1699-1733): returns zeros_like + scatter."""
    cw = cache_weight.data if isinstance(cache_weight, torch.nn.Parameter) else cache_weight
    with _DeviceGuard(cw):
        grad = torch.zeros_like(cw)
        if int(nnz) == 0:
            return grad
        go = _f32c(grad_output, "grad_output")
        loc = cache_locations if cache_locations.is_contiguous() else cache_locations.contiguous()
        _check(_lib.ttb_cache_backward_dense(int(nnz), cw.shape[1], go.data_ptr(), loc.data_ptr(),
                                             _i64c(rowidx, "rowidx").data_ptr(), grad.data_ptr(), _stream()))
        return grad


def cache_backward_rowwise_adagrad_approx(nnz: int, grad_output: torch.Tensor, cache_locations, rowidx,
                                          learning_rate: float, eps: float, cache_optimizer_state: torch.Tensor,
                                          cache_weight: torch.Tensor) -> None:
    """cache_backward_rowwise_adagrad_approx_cuda (tt_embeddings_cuda.cu:1797-1835)."""
    if int(nnz) == 0:
        return
    cw = cache_weight.data if isinstance(cache_weight, torch.nn.Parameter) else cache_weight
    if not cache_optimizer_state.is_cuda:
        raise RuntimeError("libttb: cache_optimizer_state must live on the GPU")
    with _DeviceGuard(cw):
        go = _f32c(grad_output, "grad_output")
        loc = cache_locations if cache_locations.is_contiguous() else cache_locations.contiguous()
        _check(_lib.ttb_cache_backward_rowwise_adagrad_approx(
            int(nnz), cw.shape[1], go.data_ptr(), loc.data_ptr(), _i64c(rowidx, "rowidx").data_ptr(),
            float(learning_rate), float(eps), cache_optimizer_state.data_ptr(), cw.data_ptr(), _stream()))
