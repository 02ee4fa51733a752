"""Differently-sized TT tables behind ONE plan / forward / backward / sweep launch (second step).

The reference batches tables only when their shapes are identical (``TableBatchedTTEmbeddingBag``,
tt_embeddings_ops.py:424: one ``[num_tables, p_t, S_t]`` tensor per core).  The tables of a DLRM share the
q-shapes and the ranks (same ``embedding_dim``, same rank setting) and differ only in their p-shapes -- in how
MANY slices each core has.  ``FusedTTEmbeddingBag`` therefore stores core t of all tables as one tensor
``[1, sum_k p_t(k), S_t]`` (table k owns the slice range ``[off_t(k), off_t(k) + p_t(k))``) and runs the whole
batch through ``ttb_tt_forward_het`` / ``ttb_tt_backward_het``: to the kernels this is one table; only the
index decomposition looks at the table number.  A training step of 26 tables is then 5 launches (CSR->COO,
plan, forward, backward, sweep) instead of 26 x 5, and lookups of different tables share the 32-lookup tensor
core tiles' grid instead of running as 26 small grids one after the other.

Where ``TTEmbeddingBagGroup`` (grouped.py) removes the per-table HOST cost and keeps per-table kernels (any mix
of shapes), this module removes the per-table launches too, for tables that share q-shapes and ranks.

Inputs follow ``TableBatchedTTEmbeddingBag``: ``indices`` table-major, ``offsets`` CSR over ``n_tables * B``
bags; a sequence of per-table ``(indices, offsets)`` pairs is accepted too and packed on the device.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import torch
from torch import nn

from . import tt_embeddings as ext
from .tt_embeddings_ops import _SGD_FAMILY, BufferList, OptimType, init_tt_cores, suggested_tt_shapes


def pack_table_major(indices: Sequence[torch.Tensor], offsets: Sequence[torch.Tensor]):
    """Per-table (indices, offsets[B+1]) pairs -> the table-major (indices, offsets[n*B+1]) pair of
    ``TableBatchedTTEmbeddingBag.forward`` (tt_embeddings_ops.py:821-830).  Device-side, no host sync: table k's
    offsets are shifted by the entries before it, which is the sum of the earlier tables' index counts -- a
    host-known size."""
    n = len(indices)
    if n == 0 or len(offsets) != n:
        raise RuntimeError(f"libttb: need one offsets tensor per index tensor, got {n} / {len(offsets)}")
    B = offsets[0].numel() - 1
    parts, base = [], 0
    for k in range(n):
        if offsets[k].numel() != B + 1:
            raise RuntimeError("libttb: every table of a fused batch takes the same number of bags")
        o = offsets[k].long()
        parts.append((o[:-1] if k < n - 1 else o) + base)
        base += indices[k].numel()
    return torch.cat([i.long() for i in indices]), torch.cat(parts)


class FusedTTLookupFunction(torch.autograd.Function):
    """Autograd node of the fused batch; dispatch on (sparse, optimizer) as ``TTLookupFunction``
    (tt_embeddings_ops.py:130-356)."""

    @staticmethod
    def forward(ctx, layout, B, D, tt_q_shapes, tt_ranks, indices, rowidx, tableidx, optimizer, learning_rate, eps,
                sparse, optimizer_state, *tt_cores):
        ctx.cfg = (layout, D, tt_q_shapes, tt_ranks, optimizer, learning_rate, eps, sparse)
        ctx.tt_cores = tt_cores
        ctx.optimizer_state = optimizer_state
        ctx.save_for_backward(indices, rowidx, tableidx)
        return ext.tt_forward_het(layout, B, D, tt_q_shapes, tt_ranks, indices.numel(), indices, rowidx, tableidx,
                                  list(tt_cores), keep_plan=any(ctx.needs_input_grad))

    @staticmethod
    def backward(ctx, d_output):
        layout, D, q, ranks, optimizer, lr, eps, sparse = ctx.cfg
        indices, rowidx, tableidx = ctx.saved_tensors
        cores = list(ctx.tt_cores)
        n_fixed = 13  # positional inputs before *tt_cores
        grads: List[Optional[torch.Tensor]] = [None] * (n_fixed + len(cores))
        nnz = indices.numel()
        if sparse:
            if optimizer in _SGD_FAMILY:
                ext.tt_backward_het(layout, ext.OPTIM_SGD, D, lr, 0.0, q, ranks, nnz, indices, rowidx, tableidx,
                                    d_output, cores)
            else:  # every other optimizer runs the Adagrad kernels (tt_embeddings_ops.py:248)
                ext.tt_backward_het(layout, ext.OPTIM_ADAGRAD, D, lr, eps, q, ranks, nnz, indices, rowidx, tableidx,
                                    d_output, cores, ctx.optimizer_state)
            return tuple(grads)
        grads[n_fixed:] = ext.tt_backward_het(layout, ext.OPTIM_DENSE, D, 0.0, 0.0, q, ranks, nnz, indices, rowidx,
                                              tableidx, d_output, cores)
        return tuple(grads)


class FusedTTEmbeddingBag(nn.Module):
    """``len(num_embeddings)`` TT ``EmbeddingBag(mode="sum")`` tables of different sizes, same ``embedding_dim``,
    q-shapes and ranks, looked up in one pass.

    ``tt_p_shapes[k]`` is table k's p-shape (default: ``suggested_tt_shapes(num_embeddings[k])``).  Parameters:
    ``tt_cores[t]`` = ``[1, P_t, S_t]`` with every table's slices back to back; ``table_cores(k)`` returns table
    k's ``[1, p_t(k), S_t]`` views (the tensors a ``TTEmbeddingBag`` of that table holds), ``load_table`` copies
    such tensors in.  ``forward`` returns ``[n_tables, B, D]``."""

    def __init__(self, num_embeddings: Sequence[int], embedding_dim: int, tt_ranks: List[int],
                 tt_p_shapes: Optional[Sequence[Optional[Sequence[int]]]] = None,
                 tt_q_shapes: Optional[List[int]] = None, optimizer: OptimType = OptimType.SGD,
                 learning_rate: float = 0.1, eps: float = 1.0e-10, sparse: bool = True,
                 weight_dist: str = "approx-normal", enforce_embedding_dim: bool = False,
                 device: Optional[torch.device] = None) -> None:
        """``device`` defaults to the current CUDA device (like the reference's modules, tt_embeddings_ops.py:520);
        the kernels exist for CUDA only -- there is no CPU path behind this module."""
        super().__init__()
        assert device is not None or torch.cuda.is_available()
        n = len(num_embeddings)
        assert n > 0 and embedding_dim > 0 and all(int(e) > 0 for e in num_embeddings)
        T = len(tt_ranks) + 1
        if tt_q_shapes is None:
            tt_q_shapes = suggested_tt_shapes(embedding_dim, T, allow_round_up=not enforce_embedding_dim)
        p_shapes: List[List[int]] = []
        for k in range(n):
            pk = tt_p_shapes[k] if tt_p_shapes is not None and tt_p_shapes[k] is not None else None
            if pk is None:
                pk = suggested_tt_shapes(int(num_embeddings[k]), T)
            pk = [int(v) for v in pk]
            assert len(pk) == T and all(v > 0 for v in pk)
            assert int(np.prod(pk, dtype=np.int64)) >= int(num_embeddings[k])
            p_shapes.append(pk)
        self.tt_q_shapes: List[int] = [int(v) for v in tt_q_shapes]
        assert len(self.tt_q_shapes) == T and 2 <= T <= 4
        assert int(np.prod(self.tt_q_shapes, dtype=np.int64)) == embedding_dim
        self.num_tables = n
        self.tt_ndim = T
        self.num_embeddings = [int(e) for e in num_embeddings]
        self.embedding_dim = int(embedding_dim)
        self.tt_ranks = [1] + [int(r) for r in tt_ranks] + [1]
        self.tt_p_shapes = p_shapes
        self.sparse, self.optimizer, self.learning_rate, self.eps = sparse, optimizer, learning_rate, eps
        self.layout = ext.HetLayout(p_shapes)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.tt_cores = nn.ParameterList()
        self.optimizer_state = BufferList("optimizer_state")
        for t in range(T):
            slice_elems = self.tt_ranks[t] * self.tt_q_shapes[t] * self.tt_ranks[t + 1]
            core = torch.empty((1, self.layout.P[t], slice_elems), device=dev, dtype=torch.float32)
            self.tt_cores.append(nn.Parameter(core))
            state_shape = core.shape if optimizer not in _SGD_FAMILY else (0,)
            self.optimizer_state.append(torch.zeros(state_shape, device=dev, dtype=torch.float32))
        for k in range(n):  # each table with its own num_embeddings (the scale of every scheme depends on it)
            init_tt_cores(self.table_cores(k), self.num_embeddings[k], self.embedding_dim, self.tt_ranks, p_shapes[k],
                          self.tt_q_shapes, weight_dist, 1)

    # ---- per-table views ----------------------------------------------------------------------
    def table_cores(self, k: int) -> List[torch.Tensor]:
        off, p = self.layout.off[k], self.tt_p_shapes[k]
        return [c.data[:, off[t]:off[t] + p[t]] for t, c in enumerate(self.tt_cores)]

    def table_optimizer_state(self, k: int) -> List[torch.Tensor]:
        off, p = self.layout.off[k], self.tt_p_shapes[k]
        return [s[:, off[t]:off[t] + p[t]] if s.dim() == 3 else s for t, s in enumerate(self.optimizer_state)]

    def load_table(self, k: int, tt_cores: Sequence[torch.Tensor]) -> None:
        """Copy the cores of a single-table module (``TTEmbeddingBag.tt_cores``) into table k's slice ranges."""
        with torch.no_grad():
            for dst, src in zip(self.table_cores(k), tt_cores):
                dst.copy_(src.reshape(dst.shape))

    def table_state_dict(self, k: int) -> dict:
        """Table k as the ``state_dict`` of a cache-less single-table ``TTEmbeddingBag`` (the reference's key names,
        SURVEY 5: ``tt_cores.<t>``, ``optimizer_state.optimizer_state<t>``, ``L``, empty ``hashtbl`` /
        ``cache_state``) -- what a 26-module DLRM checkpoint holds per table.  Tensors are copies."""
        dev = self.tt_cores[0].device
        sd = {f"tt_cores.{t}": c.clone() for t, c in enumerate(self.table_cores(k))}
        for t, s_ in enumerate(self.table_optimizer_state(k)):
            sd[f"optimizer_state.optimizer_state{t}"] = s_.clone()
        sd["L"] = torch.tensor([int(v) for v in self.layout.host[k].L][:self.tt_ndim], dtype=torch.int64)
        sd["hashtbl"] = torch.empty(0, dtype=torch.int64, device=dev)
        sd["cache_state"] = torch.empty(0, dtype=torch.int32, device=dev)
        return sd

    def load_table_state_dict(self, k: int, state_dict: dict) -> None:
        """Inverse of ``table_state_dict``: take table k's cores (and Adagrad state, when both sides have it) from a
        single-table module's ``state_dict``; the table's p-shape must be the one this module was built with."""
        cores = [state_dict[f"tt_cores.{t}"] for t in range(self.tt_ndim)]
        for t, (dst, src) in enumerate(zip(self.table_cores(k), cores)):
            if src.numel() != dst.numel():
                raise RuntimeError(f"libttb: table {k} core {t}: checkpoint has {tuple(src.shape)}, module {tuple(dst.shape)}")
        self.load_table(k, cores)
        with torch.no_grad():
            for t, dst in enumerate(self.table_optimizer_state(k)):
                src = state_dict.get(f"optimizer_state.optimizer_state{t}")
                if src is not None and dst.dim() == 3 and src.numel() == dst.numel():
                    dst.copy_(src.reshape(dst.shape))

    # ---- lookup -------------------------------------------------------------------------------
    def forward(self, indices: Union[torch.Tensor, Sequence[torch.Tensor]],
                offsets: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
        if not isinstance(indices, torch.Tensor):
            indices, offsets = pack_table_major(indices, offsets)
        indices, offsets = indices.long(), offsets.long()
        bags = offsets.numel() - 1
        if bags <= 0 or bags % self.num_tables:
            raise RuntimeError(f"libttb: offsets must cover {self.num_tables} x B bags, got {bags}")
        # CSR -> COO exactly as for identical tables (tt_embeddings_cuda.cu:1377-1434, multi-table branch)
        indices, rowidx, tableidx, _, _ = ext.preprocess_indices_sync(indices, offsets, self.num_tables, True,
                                                                      None, None)
        training = torch.is_grad_enabled() and any(c.requires_grad for c in self.tt_cores)
        prev = ext.set_keep_plans(training)
        try:
            return self._apply(bags, indices, rowidx, tableidx)
        finally:
            ext.set_keep_plans(prev)

    def _apply(self, bags, indices, rowidx, tableidx):
        return FusedTTLookupFunction.apply(self.layout, bags // self.num_tables, self.embedding_dim, self.tt_q_shapes,
                                           self.tt_ranks, indices, rowidx, tableidx, self.optimizer,
                                           self.learning_rate, self.eps, self.sparse, list(self.optimizer_state),
                                           *self.tt_cores)

    def set_learning_rate(self, lr: float) -> None:
        self.learning_rate = lr

    def get_params(self) -> List[torch.Tensor]:
        return list(self.tt_cores)
