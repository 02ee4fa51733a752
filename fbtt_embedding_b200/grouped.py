"""Heterogeneous TT tables looked up in ONE call per phase.

The reference can batch tables only when they have identical shapes
(``TableBatchedTTEmbeddingBag``, tt_embeddings_ops.py:424, README.md:136-141); a DLRM with 26
differently-sized tables (BASELINE config 4) therefore runs 26 modules, i.e. 26 x (cache update +
preprocess + autograd node + forward op + backward op) of Python per step, and was host-launch bound
here too (12.6 ms per step on one GPU for ~0.1 ms of kernels per table).

``TTEmbeddingBagGroup`` keeps one ``TTEmbeddingBag`` per table for parameters, initialisation and
``state_dict`` (keys ``tables.<i>.tt_cores.<t>`` ...), but its forward / backward go through the
``ttb_group_*`` entry points of libttb: one ctypes call for CSR->COO of every table, one for every
forward, one for every fused backward, under ONE autograd node.  The kernels are the same ones the
per-table modules launch, so results are bit-compatible with ``torch.stack([tbl(i, o) ...])`` up to the
usual atomic-order noise.  Host logic only; no cache support (like the reference's table-batched module).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import nn

from . import tt_embeddings as ext
from .tt_embeddings_ops import _SGD_FAMILY, OptimType, TTEmbeddingBag


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


class GroupedLookup:
    """Drives a list of cache-less, single-table ``TTEmbeddingBag`` modules through the group ABI.
    All tables must share the embedding dim, optimizer settings and device; shapes may differ."""

    def __init__(self, tables: Sequence[TTEmbeddingBag]) -> None:
        self.tables = list(tables)
        assert len(self.tables) > 0, "a table group needs at least one table"
        t0 = self.tables[0]
        for t in self.tables:
            assert t.num_tables == 1 and not t.use_cache, "table groups take cache-less single-table modules"
            assert t.embedding_dim == t0.embedding_dim, "all tables of a group share embedding_dim"
            assert (t.sparse, t.optimizer) == (t0.sparse, t0.optimizer), "all tables of a group share the optimizer"
        self.D = t0.embedding_dim
        self._template = None          # (_GroupItem * n) with the static fields, rebuilt when storages move
        self._template_key: Optional[Tuple] = None
        self._shapes: List = []        # the shim's cached ttb_shape_t objects behind the template
        self._grad_flat: Optional[torch.Tensor] = None  # zero-on-exit gradient scratch of the fused backward
        self._plan_free: Dict[Tuple, List[torch.Tensor]] = {}

    # ---- static part of the item array ------------------------------------------------------
    def _storage_key(self, B: int) -> Tuple:
        key: List[int] = [B]
        for t in self.tables:
            key.extend(c.data_ptr() for c in t.tt_cores)
            key.extend(s.data_ptr() for s in t.optimizer_state)
        return tuple(key)

    def _grad_views(self, flat: torch.Tensor) -> List[List[torch.Tensor]]:
        views, off = [], 0
        for t in self.tables:
            per = []
            for c in t.tt_cores:
                per.append(flat[off:off + c.numel()].view(c.shape))
                off += _align(c.numel(), 4)
            views.append(per)
        return views

    def _grad_numel(self) -> int:
        return sum(_align(c.numel(), 4) for t in self.tables for c in t.tt_cores)

    def _items_template(self, B: int):
        key = self._storage_key(B)
        if self._template is not None and key == self._template_key:
            return self._template
        n = len(self.tables)
        arr = (ext._GroupItem * n)()
        dev = self.tables[0].tt_cores[0].device
        if self._grad_flat is None or self._grad_flat.device != dev or self._grad_flat.numel() != self._grad_numel():
            self._grad_flat = torch.zeros(self._grad_numel(), dtype=torch.float32, device=dev)
        gviews = self._grad_views(self._grad_flat)
        self._shapes = []
        for i, t in enumerate(self.tables):
            cores = ext._cores_inplace(list(t.tt_cores))  # CUDA, fp32, contiguous, 16-byte aligned
            it = arr[i]
            shp = ext._shape(1, B, self.D, t.tt_p_shapes, t.tt_q_shapes, t.tt_ranks)
            ctypes.memmove(ctypes.byref(it.shape), ctypes.byref(shp), ctypes.sizeof(ext._Shape))
            self._shapes.append(shp)
            adagrad = t.sparse and t.optimizer not in _SGD_FAMILY
            for k, c in enumerate(cores):
                assert c.device == dev, "all tables of a group live on one device"
                it.cores[k] = c.data_ptr()
                it.grads[k] = gviews[i][k].data_ptr()
                if adagrad:
                    st = t.optimizer_state[k]
                    if st.shape != c.shape or st.dtype != torch.float32 or not st.is_contiguous():
                        raise RuntimeError("libttb: optimizer_state must be a contiguous fp32 tensor shaped like its core")
                    it.opt_state[k] = st.data_ptr()
        self._template, self._template_key = arr, key
        return arr

    # ---- plan buffers ---------------------------------------------------------------------------
    def _plan_layout(self, nnzs: Sequence[int]) -> Tuple[List[int], List[int], int]:
        """Per-item offset / size of the bucketing-plan workspaces inside one flat buffer (0 bytes for items on
        the generic path or without lookups)."""
        offs, sizes, total = [], [], 0
        for shp, nnz in zip(self._shapes, nnzs):
            b = ext._workspace_bytes(shp, nnz) if nnz else 0
            if b:  # power-of-two capacity: the layout (and the pool key) survives a drifting nnz
                b = 1 << max(12, (b - 1).bit_length())
            offs.append(total)
            sizes.append(b)
            total += b
        return offs, sizes, total

    def _plan_take(self, key: Tuple, total: int, dev) -> Tuple[Optional[torch.Tensor], bool]:
        if total == 0:
            return None, False
        capturing = torch.cuda.is_current_stream_capturing()
        free = self._plan_free.get(key)
        if free and not capturing:
            return free.pop(), True
        # header contract of include/ttb.h: the plan headers are zero on entry and the kernels leave them zero,
        # so one memset at birth serves every later step that draws this buffer from the pool
        return torch.zeros(total + 256, dtype=torch.uint8, device=dev), not capturing

    def _plan_give(self, key: Tuple, buf: Optional[torch.Tensor], poolable: bool) -> None:
        if buf is None or not poolable:
            return
        free = self._plan_free.setdefault(key, [])
        if len(self._plan_free) > 8:  # shapes drifted: start over rather than hoard
            self._plan_free = {key: free}
        if len(free) < 4:
            free.append(buf)

    # ---- the two phases -----------------------------------------------------------------------
    def lookup(self, indices: Sequence[torch.Tensor], offsets: Sequence[torch.Tensor]) -> torch.Tensor:
        cores = [c for t in self.tables for c in t.tt_cores]
        if not (torch.is_grad_enabled() and any(c.requires_grad for c in cores)):
            out, state = self._forward(indices, offsets)  # inference: no backward will return the plan buffer
            self._plan_give(state[5], state[4], state[6])
            return out
        return _GroupLookupFunction.apply(self, tuple(indices), tuple(offsets), *cores)

    def _forward(self, indices: Sequence[torch.Tensor], offsets: Sequence[torch.Tensor]):
        n = len(self.tables)
        if len(indices) != n or len(offsets) != n:
            raise RuntimeError(f"libttb: table group of {n} tables got {len(indices)} index / {len(offsets)} offset tensors")
        B = offsets[0].numel() - 1
        if B <= 0:
            raise RuntimeError("libttb: offsets must hold B + 1 >= 2 entries")
        dev = self.tables[0].tt_cores[0].device
        idx_l, off_l, nnzs = [], [], []
        for i in range(n):
            ix, of = indices[i], offsets[i]
            if ix.dtype != torch.int64:
                ix = ix.long()
            if of.dtype != torch.int64:
                of = of.long()
            ix, of = ext._i64c(ix, "indices"), ext._i64c(of, "offsets")
            if of.numel() != B + 1:
                raise RuntimeError("libttb: every table of a group takes the same number of bags")
            if ix.device != dev or of.device != dev:
                raise RuntimeError("libttb: group inputs must live on the tables' device")
            idx_l.append(ix)
            off_l.append(of)
            nnzs.append(ix.numel())
        with ext._DeviceGuard(self.tables[0].tt_cores[0]):
            template = self._items_template(B)
            items = (ext._GroupItem * n)()
            ctypes.memmove(items, template, ctypes.sizeof(items))
            total = sum(nnzs)
            out = torch.zeros((n, B, self.D), dtype=torch.float32, device=dev)
            coo = torch.empty(2 * max(total, 1), dtype=torch.int64, device=dev)
            stream = ext._stream()
            offs, sizes, ws_total = self._plan_layout(nnzs)
            plan_key = (tuple(sizes), B, ext.get_path(), stream)
            plan, poolable = self._plan_take(plan_key, ws_total, dev)
            plan_base = (plan.data_ptr() + 255) // 256 * 256 if plan is not None else 0
            coo_ptr, out_ptr, pos = coo.data_ptr(), out.data_ptr(), 0
            for i in range(n):
                it = items[i]
                it.nnz = nnzs[i]
                it.indices = idx_l[i].data_ptr()
                it.offsets = off_l[i].data_ptr()
                it.rowidx = coo_ptr + 8 * pos
                it.tableidx = coo_ptr + 8 * (total + pos)
                it.output = out_ptr + 4 * i * B * self.D
                if sizes[i]:
                    it.workspace = plan_base + offs[i]
                    it.workspace_bytes = sizes[i]
                it.plan_ready = 0
                pos += nnzs[i]
            try:
                ext._check(ext._lib.ttb_group_preprocess(n, items, stream))
                ext._check(ext._lib.ttb_group_forward(n, items, stream))
            except RuntimeError:
                poolable = False  # a plan header may be dirty: never reuse this buffer
                raise
        # everything the backward needs, and everything that must stay alive until then
        state = (items, idx_l, off_l, coo, plan, plan_key, poolable, B)
        return out, state

    def _backward(self, state, d_output: torch.Tensor) -> Optional[List[torch.Tensor]]:
        items, idx_l, off_l, coo, plan, plan_key, poolable, B = state
        n = len(self.tables)
        t0 = self.tables[0]
        d_output = ext._f32c(d_output, "d_output")
        if tuple(d_output.shape) != (n, B, self.D):
            raise RuntimeError(f"libttb: d_output must be [{n}, {B}, {self.D}], got {tuple(d_output.shape)}")
        dense = not t0.sparse
        with ext._DeviceGuard(d_output):
            stream = ext._stream()
            if stream != plan_key[3]:
                raise RuntimeError("libttb: a table group's backward must run on the stream of its forward")
            grad_views = None
            if dense:
                flat = torch.zeros(self._grad_numel(), dtype=torch.float32, device=d_output.device)
                grad_views = self._grad_views(flat)
            base = d_output.data_ptr()
            for i in range(n):
                it = items[i]
                it.d_output = base + 4 * i * B * self.D
                it.plan_ready = 1 if it.workspace_bytes else 0
                if dense:
                    for k, g in enumerate(grad_views[i]):
                        it.grads[k] = g.data_ptr()
            if dense:
                optim, lr, eps = ext.OPTIM_DENSE, 0.0, 0.0
            elif t0.optimizer in _SGD_FAMILY:
                optim, lr, eps = ext.OPTIM_SGD, float(t0.learning_rate), 0.0
            else:  # every other optimizer runs the Adagrad kernels (tt_embeddings_ops.py:248)
                optim, lr, eps = ext.OPTIM_ADAGRAD, float(t0.learning_rate), float(t0.eps)
            try:
                ext._check(ext._lib.ttb_group_backward(n, items, optim, lr, eps, stream))
            except RuntimeError:
                self._grad_flat = None  # scratch may be dirty
                self._template = None
                raise
            self._plan_give(plan_key, plan, poolable)
        if dense:
            return [g for per in grad_views for g in per]
        return None


class _GroupLookupFunction(torch.autograd.Function):
    """One autograd node for the whole group; the TT cores ride along as inputs so that the node is part of
    the graph (and, with ``sparse=False``, so that their dense gradients have somewhere to go)."""

    @staticmethod
    def forward(ctx, group: GroupedLookup, indices, offsets, *cores):
        out, state = group._forward(indices, offsets)
        ctx.group, ctx.state, ctx.n_cores = group, state, len(cores)
        return out

    @staticmethod
    def backward(ctx, d_output):
        state, ctx.state = ctx.state, None
        if state is None:
            raise RuntimeError("libttb: a table group's backward ran twice (retain_graph is not supported)")
        grads = ctx.group._backward(state, d_output)
        if grads is None:
            grads = [None] * ctx.n_cores
        return (None, None, None, *grads)


class TTEmbeddingBagGroup(nn.Module):
    """``len(specs)`` differently-shaped TT ``EmbeddingBag(mode="sum")`` tables behind one forward.

    specs[i] = dict(num_embeddings, embedding_dim, tt_ranks, tt_p_shapes=None, tt_q_shapes=None); every table
    shares ``embedding_dim`` and the optimizer settings given here (``TTEmbeddingBag`` keyword names).
    ``forward(indices, offsets)`` takes one (indices, offsets) pair per table -- all with the same number of
    bags B -- and returns the pooled rows ``[len(specs), B, D]`` (== ``torch.stack`` of the per-table results).
    """

    def __init__(self, specs: Sequence[dict], optimizer: OptimType = OptimType.SGD, learning_rate: float = 0.1,
                 eps: float = 1.0e-10, sparse: bool = True, weight_dist: str = "approx-normal",
                 enforce_embedding_dim: bool = False) -> None:
        super().__init__()
        self.tables = nn.ModuleList(
            TTEmbeddingBag(**spec, optimizer=optimizer, learning_rate=learning_rate, eps=eps, sparse=sparse,
                           use_cache=False, weight_dist=weight_dist, enforce_embedding_dim=enforce_embedding_dim)
            for spec in specs)
        self.embedding_dim = self.tables[0].embedding_dim
        self._group = GroupedLookup(list(self.tables))

    def forward(self, indices: Sequence[torch.Tensor], offsets: Sequence[torch.Tensor]) -> torch.Tensor:
        return self._group.lookup(indices, offsets)

    def set_learning_rate(self, lr: float) -> None:
        for t in self.tables:
            t.set_learning_rate(lr)

    def get_params(self) -> List[torch.Tensor]:
        return [p for t in self.tables for p in t.get_params()]
