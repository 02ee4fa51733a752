"""Table-parallel sharding of many TT-EmbeddingBag tables over the GPUs of one box.

The reference is single-GPU (no NCCL/MPI call site anywhere).  Tables are independent, so the
multi-table workload (BASELINE config 4: 26 Criteo-Terabyte tables, D=128, B=4096) shards by
TABLE: every rank owns a subset of the tables, looks up the WHOLE batch for them, and ONE
``all_to_all_single`` re-shards the pooled rows batch-wise::

    rank r:  pooled[T_local(r), B, D]  --a2a-->  out[B/W, T_total, D]     (forward)
             d_out[B/W, T_total, D]    --a2a-->  d_pooled[T_local(r), B, D]  (backward, the mirror)

One process per GPU, ``torch.distributed`` (NCCL over NVLink on the box; gloo on CPU for tests).
The exchange is the only collective on the data path; the fused TT backward then runs locally.
This file is host logic only (placement, packing, the autograd-aware exchange).

``exchange="peer"`` (with ``fused=True``) folds the exchange INTO the kernels (include/ttb.h, ``ttb_row_map_t``):
every rank owns one symmetric buffer ``[X0 | X1 | dX]`` (``[3, B/W, T_total, D]``, peer-mapped over NVLink); the
forward kernel adds each pooled row straight into ``X`` of the rank that owns the row's batch slice, the backward
kernel reads ``dX`` rows from there -- no all-to-all launch, no pack / unpack copies, the transfer overlaps the math
tile by tile.  What remains between the ranks is two stream-ordered barriers per step (rows landed -> read;
gradients written -> gather); ``X`` is double-buffered so that zeroing the next step's buffer needs no barrier of
its own (see ``_PeerLookup``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import nn


def assign_tables(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time placement: tables sorted by cost (lookups x flop per lookup),
    each goes to the currently lightest rank.  Cost of a TT table is per LOOKUP, not per row, so
    the few huge-cardinality tables do not dominate.  Deterministic."""
    order = sorted(range(len(costs)), key=lambda t: (-float(costs[t]), t))
    load = [0.0] * world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for t in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        owned[r].append(t)
        load[r] += float(costs[t])
    for o in owned:
        o.sort()
    return owned


def tt_lookup_cost(q: Sequence[int], ranks: Sequence[int], lookups: float, p: Optional[Sequence[int]] = None) -> float:
    """3F * lookups with F the forward flop per lookup (benchmark convention) -- times, when the p-shape
    is given, the TILE factor of the bucketed kernels: they work in 32-lookup tiles per middle-core index, so a
    table whose lookups spread over many buckets pays for partially filled tiles (measured on config 4: the rank
    holding four 39M-row tables, p1 ~ 350, ran 1.5x longer than a rank holding four tiny ones).  Expected tiles =
    lookups / 32 + min(p1, lookups) / 2 (half a tile of padding per occupied bucket)."""
    R = [1] + list(ranks) + [1]
    f, m = 0, 1
    for t in range(1, len(q)):
        m *= q[t - 1]
        f += 2 * m * R[t] * q[t] * R[t + 1]
    cost = 3.0 * f * float(lookups)
    if p is not None and len(p) == 3 and lookups > 0:
        cost *= 1.0 + 16.0 * min(float(p[1]), float(lookups)) / float(lookups)
    return cost


_order_cache: dict = {}


def _table_orders(owned: List[List[int]], device) -> tuple:
    """(src_order, inv) as device tensors, built once per (placement, device): src_order lists the global table
    numbers in source-major order (rank 0's tables, rank 1's, ...), inv is its inverse permutation.  Cached so
    that a training step does no host -> device copy of index tensors (and can be captured in a CUDA graph)."""
    key = (tuple(tuple(o) for o in owned), str(device))
    hit = _order_cache.get(key)
    if hit is None:
        src_order = [t for o in owned for t in o]
        inv = torch.empty(len(src_order), dtype=torch.long)
        inv[torch.tensor(src_order, dtype=torch.long)] = torch.arange(len(src_order))
        hit = (torch.tensor(src_order, dtype=torch.long).to(device), inv.to(device))
        if len(_order_cache) > 64:
            _order_cache.clear()
        _order_cache[key] = hit
    return hit


class _ExchangePooled(torch.autograd.Function):
    """pooled [T_local, B, D] on every rank -> [B/W, T_total, D]; backward is the mirror exchange.
    One pack copy + the collective + one unpack copy per direction."""

    @staticmethod
    def forward(ctx, pooled: torch.Tensor, owned: List[List[int]], group) -> torch.Tensor:
        W = dist.get_world_size(group)
        r = dist.get_rank(group)
        t_local, B, D = pooled.shape
        assert t_local == len(owned[r]) and B % W == 0
        bw = B // W
        ctx.owned, ctx.group, ctx.shape = owned, group, (t_local, B, D)
        # destination-major packing: chunk w = pooled[:, w*bw:(w+1)*bw, :]
        send = pooled.view(t_local, W, bw, D).permute(1, 0, 2, 3).contiguous()
        t_total = sum(len(o) for o in owned)
        recv = pooled.new_empty(t_total * bw * D)
        in_splits = [t_local * bw * D] * W
        out_splits = [len(owned[s]) * bw * D for s in range(W)]
        dist.all_to_all_single(recv, send.view(-1), out_splits, in_splits, group=group)
        # source-major [sum_s T_local(s), bw, D] -> [bw, T_total (global order), D] in one gather
        _, inv = _table_orders(owned, pooled.device)
        return torch.index_select(recv.view(t_total, bw, D).permute(1, 0, 2), 1, inv)

    @staticmethod
    def backward(ctx, d_out: torch.Tensor):
        owned, group = ctx.owned, ctx.group
        W = dist.get_world_size(group)
        t_local, B, D = ctx.shape
        bw = B // W
        t_total = sum(len(o) for o in owned)
        # [bw, T_total, D] -> [T_total (grouped by owner rank), bw, D] in one gather
        order, _ = _table_orders(owned, d_out.device)
        send = torch.index_select(d_out.permute(1, 0, 2), 0, order)
        in_splits = [len(owned[s]) * bw * D for s in range(W)]
        out_splits = [t_local * bw * D] * W
        recv = d_out.new_empty(W * t_local * bw * D)
        dist.all_to_all_single(recv, send.view(-1), out_splits, in_splits, group=group)
        d_pooled = recv.view(W, t_local, bw, D).permute(1, 0, 2, 3).reshape(t_local, B, D)
        assert t_total >= t_local
        return d_pooled.contiguous(), None, None


def exchange_pooled(pooled: torch.Tensor, owned: List[List[int]], group=None) -> torch.Tensor:
    return _ExchangePooled.apply(pooled, owned, group)


class LocalPeers:
    """All "ranks" of a fused exchange inside ONE process on one device: rank w's ``[X | dX]`` buffer is an ordinary
    tensor, the peer offsets are the distances between those allocations (one unified address space -- exactly what
    peer-mapped NVLink memory looks like to a kernel).  Lets a single GPU validate the mapped kernels; the caller
    drives the phases of all ranks itself, so ``barrier`` has nothing to wait for."""

    def __init__(self, world: int, rows_per_rank: int, tables_total: int, D: int, device) -> None:
        self.world = world
        self.bufs = [torch.zeros(3, rows_per_rank, tables_total, D, dtype=torch.float32, device=device)
                     for _ in range(world)]

    def view(self, rank: int) -> "PeerView":
        base = self.bufs[rank].data_ptr()
        offs = []
        for b in self.bufs:
            assert (b.data_ptr() - base) % 4 == 0
            offs.append((b.data_ptr() - base) // 4)
        return PeerView(self.bufs[rank], offs, lambda: None)


class PeerView:
    """One rank's window on the exchange buffers: ``x`` (the current one of the two ``X`` regions) and ``dx`` are
    the LOCAL ``[B/W, T_total, D]`` regions, ``peer_offset[w]`` the distance (in floats) to rank w's buffer -- the
    same for every region, all ranks share one layout -- and ``barrier()`` a stream-ordered rank barrier."""

    def __init__(self, buf: torch.Tensor, peer_offset: Sequence[int], barrier) -> None:
        self.buf, self.xs, self.dx = buf, (buf[0], buf[1]), buf[2]
        self.cur = 0
        self.peer_offset = [int(v) for v in peer_offset]
        self.barrier = barrier

    @property
    def x(self) -> torch.Tensor:
        return self.xs[self.cur]

    def flip(self) -> None:
        """Start a step: the other X region becomes current (it was zeroed during the previous step) and the one
        just retired is zeroed for the step after this."""
        self.cur ^= 1
        self.xs[self.cur ^ 1].zero_()


def symmetric_peers(rows_per_rank: int, tables_total: int, D: int, device, group=None) -> PeerView:
    """The real thing: ``torch.distributed._symmetric_memory`` allocates the ``[X | dX]`` buffer of every rank and
    maps the peers' buffers into this process (CUDA VMM handles over NVLink / NVSwitch).  Collective."""
    import torch.distributed._symmetric_memory as symm_mem

    grp = group if group is not None else dist.group.WORLD
    buf = symm_mem.empty((3, rows_per_rank, tables_total, D), dtype=torch.float32, device=device)
    buf.zero_()
    hdl = symm_mem.rendezvous(buf, grp)
    base = int(hdl.buffer_ptrs[hdl.rank])
    offs = []
    for ptr in hdl.buffer_ptrs:
        if (int(ptr) - base) % 16:
            raise RuntimeError("symmetric buffers are not mutually 16-byte aligned")
        offs.append((int(ptr) - base) // 4)
    view = PeerView(buf, offs, lambda: hdl.barrier(channel=0))
    view._hdl = hdl  # keeps the mapping alive
    view.barrier()   # every rank's buffers are zero before the first scatter
    return view


class _PeerLookup(torch.autograd.Function):
    """Forward + backward of one rank with the exchange folded into the TT kernels.

    Ordering without a "buffers are zero" barrier: X is double-buffered.  Step k scatters into X[k % 2] and, before
    its barrier, zeroes X[(k+1) % 2] -- whose rows (step k-1) this rank has already copied out, and which no peer adds
    into before it has passed THIS step's barrier, i.e. after the zeroing (stream order).  dX is written after the
    forward's barrier of the same step, which no rank passes before its previous backward has finished reading."""

    @staticmethod
    def forward(ctx, mod, peers: PeerView, indices, offsets, *cores):
        ctx.mod, ctx.peers = mod, peers
        # ONE dX region and a two-deep X ring: the backward of a step is only safe while that step is the LATEST
        # forward (a newer forward has recycled the buffers, and two backwards in a row would overwrite dX while a
        # peer may still gather from it).  Each forward takes a ticket; the backward checks it.
        peers.step = getattr(peers, "step", 0) + 1
        ctx.step = peers.step
        peers.flip()
        ctx.state = mod._phase_forward(peers, indices, offsets)
        peers.barrier()                      # every rank's rows have landed in my X
        return peers.x.clone()               # X is recycled two steps from now

    @staticmethod
    def backward(ctx, d_out):
        mod, peers, state = ctx.mod, ctx.peers, ctx.state
        ctx.state = None
        if ctx.step != getattr(peers, "step", 0) or getattr(peers, "backward_of", 0) == ctx.step:
            raise RuntimeError("exchange=\"peer\": this backward belongs to a forward whose exchange buffers have been "
                               "recycled (one training step in flight: forward, backward, forward, ...); use "
                               "exchange=\"nccl\" for several micro-batches in flight")
        peers.backward_of = ctx.step
        peers.dx.copy_(d_out)
        peers.barrier()                      # every rank's dX is in place before anyone gathers from it
        grads = mod._phase_backward(peers, state)
        return (None, None, None, None, *(grads if grads is not None else [None] * len(mod.fused.tt_cores)))


class TableShardedTTEmbeddingBag(nn.Module):
    """``len(specs)`` TT tables sharded table-parallel over the ranks of ``group``.

    specs[t] = dict(num_embeddings, embedding_dim, tt_ranks, tt_p_shapes, tt_q_shapes); all tables share
    ``embedding_dim``.  forward(indices, offsets) takes, per LOCAL table (in ``self.local_tables`` order), the
    indices / offsets of the GLOBAL batch and returns this rank's batch slice ``[B/W, T_total, D]``.
    """

    def __init__(self, specs: Sequence[dict], lookups_per_table: Optional[Sequence[float]] = None, group=None,
                 grouped: bool = False, fused: bool = False, exchange: str = "nccl",
                 world_size: Optional[int] = None, rank: Optional[int] = None, **tt_kwargs) -> None:
        """``grouped=True`` runs this rank's tables through the table-group entry points (one host call per phase
        for all of them, ``fbtt_embedding_b200/grouped.py``) instead of one module call per table; parameters,
        ``state_dict`` keys and results are the same.  ``fused=True`` (tables must share q-shapes and ranks, as
        a DLRM's do) stores this rank's tables as ONE ``FusedTTEmbeddingBag`` -- concatenated cores, one plan /
        forward / backward / sweep launch for all of them (``fbtt_embedding_b200/fused.py``); ``state_dict`` keys
        are then ``fused.tt_cores.<t>`` and ``fused.table_cores(i)`` gives local table i's cores.
        ``exchange="peer"`` (needs ``fused=True``): the all-to-all is folded into the TT kernels over peer-mapped
        symmetric memory (module docstring); ``"nccl"`` (default) is one ``all_to_all_single`` each way."""
        super().__init__()
        from .tt_embeddings_ops import TTEmbeddingBag

        self.group = group
        # (world_size, rank) default to the process group's; given explicitly, one process can hold several
        # "ranks" (LocalPeers: single-GPU validation of the fused exchange)
        W = int(world_size) if world_size is not None else dist.get_world_size(group)
        r = int(rank) if rank is not None else dist.get_rank(group)
        lookups = list(lookups_per_table) if lookups_per_table is not None else [1.0] * len(specs)
        costs = [tt_lookup_cost(s["tt_q_shapes"], s["tt_ranks"], n, s.get("tt_p_shapes")) for s, n in zip(specs, lookups)]
        self.owned = assign_tables(costs, W)
        self.local_tables = self.owned[r]
        self.embedding_dim = int(specs[0]["embedding_dim"])
        assert all(int(s["embedding_dim"]) == self.embedding_dim for s in specs)
        tt_kwargs.setdefault("use_cache", False)
        self.fused = None
        self._grouped = None
        assert exchange in ("nccl", "peer") and (exchange == "nccl" or fused), 'exchange="peer" needs fused=True'
        self.exchange = exchange
        self._peers = None      # PeerView of the current batch size (allocated collectively at first use)
        self._row_map = None
        self.tables_total = len(specs)
        if fused and len(self.local_tables):
            from .fused import FusedTTEmbeddingBag

            mine = [specs[t] for t in self.local_tables]
            assert all(list(s["tt_q_shapes"]) == list(mine[0]["tt_q_shapes"]) and
                       list(s["tt_ranks"]) == list(mine[0]["tt_ranks"]) for s in mine), \
                "fused=True needs tables that share tt_q_shapes and tt_ranks"
            tt_kwargs.pop("use_cache")
            self.tables = nn.ModuleList()
            self.fused = FusedTTEmbeddingBag([s["num_embeddings"] for s in mine], self.embedding_dim,
                                             list(mine[0]["tt_ranks"]), [s.get("tt_p_shapes") for s in mine],
                                             list(mine[0]["tt_q_shapes"]), **tt_kwargs)
            return
        self.tables = nn.ModuleList(TTEmbeddingBag(**specs[t], **tt_kwargs) for t in self.local_tables)
        if grouped and len(self.tables):
            from .grouped import GroupedLookup

            self._grouped = GroupedLookup(list(self.tables))

    # ---- fused exchange (exchange="peer") -------------------------------------------------------------
    def _peer_setup(self, peers: PeerView, B: int) -> None:
        """Bind the exchange buffers of a batch of B bags (B/W per rank) and build the row map over them."""
        from . import tt_embeddings as ext

        W = len(peers.peer_offset)
        assert B % W == 0 and tuple(peers.x.shape) == (B // W, self.tables_total, self.embedding_dim)
        self._peers = peers
        self._row_map = ext.RowMap(W, B // W, self.tables_total, peers.peer_offset, self.local_tables, peers.x.device)

    def _phase_forward(self, peers: PeerView, indices, offsets):
        """CSR->COO + the forward kernels of this rank; pooled rows are ADDED into the peers' X (zeroed before)."""
        from . import tt_embeddings as ext
        from .fused import pack_table_major

        f = self.fused
        if not isinstance(indices, torch.Tensor):
            indices, offsets = pack_table_major(indices, offsets)
        indices, offsets = indices.long(), offsets.long()
        B = (offsets.numel() - 1) // f.num_tables
        indices, rowidx, tableidx, _, _ = ext.preprocess_indices_sync(indices, offsets, f.num_tables, True, None, None)
        ext.tt_forward_het(f.layout, B, f.embedding_dim, f.tt_q_shapes, f.tt_ranks, indices.numel(), indices, rowidx,
                           tableidx, list(f.tt_cores), row_map=self._row_map, out=peers.x)
        return indices, rowidx, tableidx

    def _phase_backward(self, peers: PeerView, state):
        """The fused TT backward of this rank; d_output rows are READ from the peers' dX."""
        from . import tt_embeddings as ext
        from .tt_embeddings_ops import _SGD_FAMILY

        f = self.fused
        indices, rowidx, tableidx = state
        common = (f.embedding_dim, f.learning_rate)
        if not f.sparse:
            return ext.tt_backward_het(f.layout, ext.OPTIM_DENSE, common[0], 0.0, 0.0, f.tt_q_shapes, f.tt_ranks,
                                       indices.numel(), indices, rowidx, tableidx, peers.dx, list(f.tt_cores),
                                       row_map=self._row_map)
        sgd = f.optimizer in _SGD_FAMILY
        ext.tt_backward_het(f.layout, ext.OPTIM_SGD if sgd else ext.OPTIM_ADAGRAD, common[0], common[1],
                            0.0 if sgd else f.eps, f.tt_q_shapes, f.tt_ranks, indices.numel(), indices, rowidx, tableidx,
                            peers.dx, list(f.tt_cores), None if sgd else list(f.optimizer_state),
                            row_map=self._row_map)
        return None

    def forward(self, indices, offsets) -> torch.Tensor:
        """``indices`` / ``offsets``: one tensor per LOCAL table, or -- with ``fused=True`` -- the table-major pair of
        ``TableBatchedTTEmbeddingBag`` (one index tensor, offsets over ``T_local * B`` bags; what a keyed-jagged input
        pipeline delivers), which skips the per-step packing of ``fused.pack_table_major``."""
        packed = isinstance(indices, torch.Tensor)
        assert (packed and self.fused is not None) or len(indices) == len(self.local_tables) == len(offsets)
        if self.exchange == "peer":
            if self.fused is None:
                raise RuntimeError("this rank owns no table; use fewer ranks than tables")
            B = (offsets.numel() - 1) // len(self.local_tables) if packed else offsets[0].numel() - 1
            W = dist.get_world_size(self.group)
            if self._peers is None or self._peers.x.shape[0] * W != B:
                self._peer_setup(symmetric_peers(B // W, self.tables_total, self.embedding_dim,
                                                 self.fused.tt_cores[0].device, self.group), B)
            return _PeerLookup.apply(self, self._peers, indices if packed else tuple(indices),
                                     offsets if packed else tuple(offsets), *self.fused.tt_cores)
        if self.fused is not None:
            pooled = self.fused(indices, offsets)
        elif self._grouped is not None:
            pooled = self._grouped.lookup(indices, offsets)
        elif len(self.tables):
            pooled = torch.stack([tbl(i, o) for tbl, i, o in zip(self.tables, indices, offsets)])
        else:  # a rank may own no table when there are fewer tables than ranks
            raise RuntimeError("this rank owns no table; use fewer ranks than tables")
        return exchange_pooled(pooled, self.owned, self.group)
