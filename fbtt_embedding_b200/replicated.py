"""Data-parallel replicas of ONE TT-EmbeddingBag table.

The reference is single-GPU.  TT cores are tiny (3.8 MB at the README shape), so a single big table scales
over GPUs by REPLICATING the cores and sharding the BAGS: every rank looks up its slice of the batch, and the
fused backward is split at its natural seam::

    ttb_tt_backward(TTB_OPTIM_DENSE)   per-rank gradient of the cores, into one flat fp32 buffer
    all_reduce(flat, SUM)              the only collective: sum(numel(core_t)) floats over NCCL / NVLink
    ttb_optimizer_step                 SGD / Adagrad from the summed gradient, identical on every rank

which is arithmetically the single-GPU fused step on the concatenated batch (sum over bags commutes with the
shard sum; Adagrad squares the SUMMED gradient, exactly as one GPU would).  Replicas therefore stay equal up
to the fp32 order of the all-reduce, which NCCL makes identical on all ranks.
Host logic only; one process per GPU, ``torch.distributed`` (NCCL on the box, gloo on CPU in the tests).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist
from torch import nn


def shard_bags(indices: torch.Tensor, offsets: torch.Tensor, rank: int, world_size: int
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Rank ``rank``'s contiguous slice of a CSR batch (``offsets`` includes the last offset, as everywhere in
    the reference: tt_embeddings_ops.py:821-874).  Bags are split as evenly as possible, the first
    ``B % world_size`` ranks get one more.  Returns (local indices, local offsets starting at 0)."""
    B = offsets.numel() - 1
    base, extra = divmod(B, world_size)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    lo, hi = int(offsets[b0]), int(offsets[b1])
    return indices[lo:hi], offsets[b0:b1 + 1] - lo


def allreduce_and_step(flat: torch.Tensor, apply_update: Callable[[], None], group=None,
                       average: bool = False) -> None:
    """The replica-synchronising half of the step: sum the flat core gradient over the ranks of ``group``, then
    run ``apply_update`` (which reads the summed gradient).  ``average`` divides by the world size first (loss
    averaged over the global batch instead of summed)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(dist.get_world_size(group))
    apply_update()


class _ReplicatedLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner: "ReplicatedTTEmbeddingBag", indices: torch.Tensor, offsets: torch.Tensor,
                *tt_cores: torch.Tensor) -> torch.Tensor:
        from . import tt_embeddings

        tbl = owner.table
        empty_tbl = torch.empty(0, dtype=torch.int64, device=indices.device)
        empty_state = torch.empty(0, dtype=torch.int32, device=indices.device)
        col, rowidx, tableidx, nnz, _ = tt_embeddings.preprocess_indices_sync(indices, offsets, 1, True, empty_tbl,
                                                                              empty_state)
        B = offsets.numel() - 1
        ctx.owner, ctx.nnz, ctx.B = owner, nnz, B
        ctx.save_for_backward(col, rowidx, tableidx)
        cores = [c.data for c in tbl.tt_cores]
        out = tt_embeddings.tt_forward(1000, 1, B, tbl.embedding_dim, tbl.tt_p_shapes, tbl.tt_q_shapes,
                                       tbl.tt_ranks, tbl.L, nnz, col, rowidx, tableidx, cores)
        return out[0]

    @staticmethod
    def backward(ctx, d_output: torch.Tensor):
        from . import tt_embeddings
        from .tt_embeddings_ops import _SGD_FAMILY

        owner = ctx.owner
        tbl = owner.table
        col, rowidx, tableidx = ctx.saved_tensors
        cores = [c.data for c in tbl.tt_cores]
        flat, views = tt_embeddings.grad_scratch(cores)
        d_out = d_output.contiguous().view(1, ctx.B, tbl.embedding_dim)
        sgd = tbl.optimizer in _SGD_FAMILY
        state = None if sgd else list(tbl.optimizer_state)

        def apply_update() -> None:
            tt_embeddings.optimizer_step(tt_embeddings.OPTIM_SGD if sgd else tt_embeddings.OPTIM_ADAGRAD,
                                         tbl.learning_rate, tbl.eps, 1, ctx.B, tbl.embedding_dim, tbl.tt_p_shapes,
                                         tbl.tt_q_shapes, tbl.tt_ranks, cores, views, state)

        try:
            tt_embeddings.tt_dense_backward_into(tbl.embedding_dim, tbl.tt_p_shapes, tbl.tt_q_shapes, tbl.tt_ranks,
                                                 ctx.nnz, col, rowidx, tableidx, d_out, cores, views)
            allreduce_and_step(flat, apply_update, owner.group, owner.average)
        except BaseException:
            tt_embeddings._drop_grad_scratch()  # the shared zero-on-exit scratch may hold a partial gradient
            raise
        return (None, None, None) + (None,) * len(cores)  # fused: the cores are already updated


class ReplicatedTTEmbeddingBag(nn.Module):
    """One TT table replicated on every rank of ``group``; ``forward(indices, offsets)`` takes THIS rank's bags
    (see ``shard_bags``) and returns their pooled rows ``[B_local, D]``; ``backward`` performs the synchronised
    fused update described in the module docstring.  Constructor arguments are ``TTEmbeddingBag``'s; the LFU
    cache is not available here (a cached row would need its own cross-replica update)."""

    def __init__(self, *args, group=None, average: bool = False, **kwargs) -> None:
        super().__init__()
        from .tt_embeddings_ops import TTEmbeddingBag

        if kwargs.get("use_cache", False):
            raise ValueError("ReplicatedTTEmbeddingBag: use_cache is not supported")
        kwargs["use_cache"] = False
        kwargs["sparse"] = True
        self.group, self.average = group, average
        self.table = TTEmbeddingBag(*args, **kwargs)
        self.sync_replicas()

    def sync_replicas(self) -> None:
        """Broadcast rank 0's cores and optimizer state (replicas are initialised from different random streams)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            for t in list(self.table.tt_cores) + list(self.table.optimizer_state):
                if t.numel():
                    dist.broadcast(t.data, src=src, group=self.group)

    def forward(self, indices: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        return _ReplicatedLookup.apply(self, indices.long(), offsets.long(), *self.table.tt_cores)
