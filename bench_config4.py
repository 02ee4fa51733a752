"""BASELINE configs[3] for bench.py: DLRM-style 26 tables (Criteo-Terabyte cardinalities, synthetic zipf indices),
D=128, ranks [64,64], q=[4,4,8], B=4096, one-hot (L=1), table-sharded over the ranks of one box with ONE exchange of
pooled rows forward and its mirror backward (fbtt_embedding_b200/sharded.py).

  strong scaling  the global batch is fixed (26 x 4096 lookups per step); N ranks own ~26/N tables each
  exchange        "nccl": one all_to_all_single each way;  "peer": folded into the TT kernels over symmetric memory
                  (pooled rows red.add-ed straight into the owner rank's batch slice over NVLink).  Both are timed,
                  the better one is the headline, the other is reported beside it.
  parity          before timing, ONE training step is checked against the per-table numpy oracle: pooled rows of
                  every table as delivered by the exchange (all ranks' batch slices gathered) and the fused-SGD
                  cores of this rank's tables.
  e2e             per-step H2D of the rank's pinned (table-major) indices + offsets, D2H of its [B/N, 26, D] slice.
"""
import os
import sys
import time

import numpy as np

CARD = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10, 2208, 11938, 155, 4,
        976, 14, 39979771, 25641295, 39664984, 585935, 12972, 108, 36]
# p-shapes: the reference's suggested_tt_shapes(E, 3) (pinned in tests/golden/suggested_shapes.json)
PSHAPE = {39884406: [304, 350, 375], 38532951: [320, 325, 375], 39979771: [334, 342, 350], 39664984: [250, 397, 400],
          25641295: [285, 300, 300], 2953546: [130, 142, 160], 585935: [75, 80, 100], 403346: [50, 82, 100],
          39043: [25, 40, 40], 20263: [25, 28, 29], 17289: [24, 25, 30], 12972: [20, 25, 26], 11938: [20, 24, 25],
          7420: [20, 20, 20], 7120: [20, 20, 20], 2208: [10, 13, 17], 1543: [10, 10, 16], 976: [10, 10, 10],
          155: [5, 5, 8], 108: [5, 5, 8], 63: [3, 3, 7], 36: [3, 3, 4], 14: [2, 2, 5], 10: [1, 2, 5], 4: [1, 2, 2],
          3: [1, 1, 3]}
D, RANKS, Q, B, POOL = 128, [64, 64], [4, 4, 8], 4096, 1
LR, ZIPF_A, N_BATCHES = 0.1, 1.2, 4
NNZ_TABLE = B * POOL
NNZ_STEP = NNZ_TABLE * len(CARD)
F_FWD = 2 * (Q[0] * RANKS[0] * Q[1] * RANKS[1] + Q[0] * Q[1] * RANKS[1] * Q[2])  # 147456 flop / lookup

WORKLOAD = ("BASELINE configs[3]: 26 Criteo-Terabyte tables D=128 q=[4,4,8] ranks=[64,64] B=4096 one-hot, zipf(1.2) "
            "indices, sparse fused SGD fp32 cores, table-sharded, one exchange of pooled rows each way")


def specs():
    return [dict(num_embeddings=E, embedding_dim=D, tt_ranks=RANKS, tt_p_shapes=PSHAPE[E], tt_q_shapes=Q) for E in CARD]


def make_batches(seed=1):
    """The same synthetic stream on every rank (replicated inputs): N_BATCHES x 26 index arrays."""
    rng = np.random.RandomState(seed)
    return [[(rng.zipf(ZIPF_A, size=NNZ_TABLE) % E).astype(np.int64) for E in CARD] for _ in range(N_BATCHES)]


def oracle_check(model, fused, local_tables, batch, out_slices, g_full, cores_before, rank):
    """One step against oracle/tt_oracle.py (numpy): pooled rows of this rank's tables for the WHOLE batch (as the
    exchange delivered them to all ranks) and the cores after the fused SGD step.  Returns (fwd_err, bwd_err)."""
    from oracle import tt_oracle as O

    fwd_err, bwd_err = 0.0, 0.0
    rowidx = np.arange(NNZ_TABLE, dtype=np.int64) // POOL
    tbl0 = np.zeros(NNZ_TABLE, dtype=np.int64)
    for k, t in enumerate(local_tables):
        p = PSHAPE[CARD[t]]
        L = O.make_L(p)
        cores0 = [c[k] for c in cores_before]
        want = O.tt_forward(1, B, D, p, Q, RANKS, L, NNZ_TABLE, batch[t], rowidx, tbl0, cores0)[0]  # [B, D]
        got = out_slices[:, t, :]  # [B, D] of table t, every rank's batch slice
        scale = max(float(np.abs(want).max()), 1e-30)
        fwd_err = max(fwd_err, float(np.abs(got - want).max()) / scale)
        if k == 0:  # the backward oracle keeps per-lookup intermediates in fp64: one table per rank is enough
            grads = O.tt_backward_dense(D, p, Q, RANKS, L, NNZ_TABLE, batch[t], rowidx, tbl0, g_full[:, t, :][None], cores0)
            want_c = O.sgd_step(cores0, grads, LR)
            now = [c.detach().cpu().numpy() for c in fused.table_cores(k)]
            for a, b in zip(now, want_c):
                bwd_err = max(bwd_err, float(np.abs(a - b).max()) / max(float(np.abs(b).max()), 1e-30))
    return fwd_err, bwd_err


def run(args, rank, local_rank, world, dev, steps, warmup, flush_buf, exchanges=("nccl", "peer"), check=True,
        graphs=True):
    """Times the table-sharded config-4 training step on `world` ranks.  Returns a dict (rank 0 prints it)."""
    import torch
    import torch.distributed as dist

    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200 import tt_embeddings as ext
    from fbtt_embedding_b200.fused import FusedTTEmbeddingBag, pack_table_major
    from fbtt_embedding_b200.sharded import TableShardedTTEmbeddingBag

    torch.manual_seed(0)  # identical table initialisation order on every rank is not needed: tables are owned
    host_batches = make_batches()
    bw = B // world
    res = {"workload": WORKLOAD, "n_gpus": world, "nnz_per_step": NNZ_STEP, "exchange": {}, "errors": {}}

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, n_steps, n_warm, sync_each=False):
        for i in range(n_warm):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        for i in range(n_steps):
            flush_buf.fill_(i & 0xFF)  # > L2: every step starts cold
            evs[i][0].record()
            fn(n_warm + i)
            evs[i][1].record()
            if sync_each:
                torch.cuda.current_stream().synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / n_steps)

    best = None
    exchanges = [e for e in exchanges if not (e == "peer" and world == 1)]
    # every model is built BEFORE the first graph capture: initialisation draws from the CUDA generator, which a
    # failed capture can leave unusable
    built = {}
    for exch in exchanges:
        if world == 1:
            built[exch] = (None, FusedTTEmbeddingBag(CARD, D, RANKS, [PSHAPE[E] for E in CARD], Q, optimizer=OptimType.SGD,
                                                     learning_rate=LR, sparse=True, weight_dist="uniform"),
                           list(range(len(CARD))))
        else:
            m = TableShardedTTEmbeddingBag(specs(), [NNZ_TABLE] * len(CARD), fused=True, exchange=exch,
                                           optimizer=OptimType.SGD, learning_rate=LR, sparse=True, weight_dist="uniform")
            built[exch] = (m, m.fused, m.local_tables)
    for exch in exchanges:
        try:
            model, fused, local_tables = built.pop(exch)
            off1 = torch.arange(0, NNZ_TABLE + 1, POOL, dtype=torch.int64)
            packed = [pack_table_major([torch.from_numpy(b[t]) for t in local_tables], [off1] * len(local_tables))
                      for b in host_batches]
            h_idx = [p[0].pin_memory() for p in packed]
            h_off = packed[0][1].pin_memory()
            d_idx = [h.to(dev) for h in h_idx]
            d_off = h_off.to(dev)
            g_full = (torch.rand(B, len(CARD), D, generator=torch.Generator().manual_seed(7)) * 0.1)
            g = (g_full[rank * bw:(rank + 1) * bw] if world > 1 else g_full.permute(1, 0, 2)).contiguous().to(dev)

            def lookup(idx, off):
                return fused(idx, off) if world == 1 else model(idx, off)

            # ---- parity of one training step against the oracle (before any timing) ----
            if check:
                cores_before = [[c.detach().cpu().numpy() for c in fused.table_cores(k)] for k in range(len(local_tables))]
                cores_before = [[cores_before[k][t] for k in range(len(local_tables))] for t in range(3)]
                out = lookup(d_idx[0], d_off)
                out.backward(g)
                torch.cuda.synchronize()
                if world > 1:
                    gathered = [torch.empty_like(out) for _ in range(world)]
                    dist.all_gather(gathered, out.detach().contiguous())
                    full = torch.cat(gathered, 0).cpu().numpy()  # [B, 26, D]
                else:
                    full = out.detach().permute(1, 0, 2).cpu().numpy()
                fe, be = oracle_check(model, fused, local_tables, host_batches[0], full, g_full.numpy(), cores_before, rank)
                fe, be = max_over_ranks(fe), max_over_ranks(be)
                res["exchange"].setdefault(exch, {})["parity"] = {"fwd_max_rel": fe, "bwd_core_max_rel": be}
                if not (fe < 1e-3 and be < 1e-2):
                    raise RuntimeError(f"config-4 parity failed ({exch}): forward {fe:.3g} (bound 1e-3), cores {be:.3g} (1e-2)")
                # nothing of the checked step stays alive into the captures below
                del out, full, cores_before
                if world > 1:
                    del gathered
                import gc

                gc.collect()
                torch.cuda.synchronize()

            def step_eager(i):
                lookup(d_idx[i % N_BATCHES], d_off).backward(g)

            ext_launch0 = ext.launch_count()
            eager_ms = timed(step_eager, steps, warmup)
            launches = (ext.launch_count() - ext_launch0) / float(steps + warmup)
            entry = res["exchange"].setdefault(exch, {})
            entry.update({"eager_ms_per_step": eager_ms, "libttb_launches_per_step": launches})
            ms, mode = eager_ms, "eager"

            # ---- the same step captured in a CUDA graph (static index buffer, D2D copy of the request) ----
            graph = None
            if graphs and os.environ.get("BENCH_CFG4_GRAPH", "1") != "0":
                try:
                    static = d_idx[0].clone()
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        for _ in range(3):
                            lookup(static, d_off).backward(g)
                    torch.cuda.current_stream().wait_stream(s)
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        static_out = lookup(static, d_off)
                        static_out.backward(g)
                    torch.cuda.synchronize()

                    def step_graph(i):
                        static.copy_(d_idx[i % N_BATCHES])
                        graph.replay()

                    graph_ms = timed(step_graph, steps, warmup)
                    entry["graph_ms_per_step"] = graph_ms
                    if graph_ms < ms:
                        ms, mode = graph_ms, "cuda_graph_replay (static index buffer)"
                except Exception as ex:  # pragma: no cover
                    # a capture that failed half-way leaves the caching allocator routing this stream's allocations to
                    # the dead graph's pool (and aborts the process at exit): hand the stream back
                    try:
                        torch._C._cuda_endAllocateToPool(dev.index, graph.pool())
                        torch._C._cuda_releasePool(dev.index, graph.pool())
                    except Exception:
                        pass
                    try:
                        torch.cuda.synchronize()
                    except Exception:
                        pass
                    graph = None
                    entry["graph_unavailable"] = f"{type(ex).__name__}: {ex}"[:300]
                    if rank == 0:
                        import traceback

                        sys.stderr.write(f"[bench config4] {exch}: graph capture unavailable:\n{traceback.format_exc()}\n")

            # ---- e2e: pinned host indices in, this rank's batch slice out, every step ----
            host_out = torch.empty((bw, len(CARD), D) if world > 1 else (len(CARD), B, D)).pin_memory()
            if graph is not None:
                stage = torch.empty_like(h_idx[0]).pin_memory()
                g2 = torch.cuda.CUDAGraph()
                s_idx, s_off = torch.empty_like(d_idx[0]), torch.empty_like(d_off)
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    s_idx.copy_(stage, non_blocking=True)
                    s_off.copy_(h_off, non_blocking=True)
                    lookup(s_idx, s_off).backward(g)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                with torch.cuda.graph(g2, pool=graph.pool()):
                    s_idx.copy_(stage, non_blocking=True)
                    s_off.copy_(h_off, non_blocking=True)
                    o2 = lookup(s_idx, s_off)
                    o2.backward(g)
                    host_out.copy_(o2.detach(), non_blocking=True)
                torch.cuda.synchronize()

                def step_e2e(i):
                    stage.copy_(h_idx[i % N_BATCHES])
                    g2.replay()
                    torch.cuda.current_stream().synchronize()
            else:
                def step_e2e(i):
                    idx = h_idx[i % N_BATCHES].to(dev, non_blocking=True)
                    off = h_off.to(dev, non_blocking=True)
                    o = lookup(idx, off)
                    o.backward(g)
                    host_out.copy_(o.detach(), non_blocking=True)
                    torch.cuda.current_stream().synchronize()

            e2e_ms = timed(step_e2e, steps, warmup)
            entry.update({"ms_per_step": ms, "mode": mode, "e2e_ms_per_step": e2e_ms,
                          "e2e_mode": "cuda_graph_replay+sync" if graph is not None else "eager+sync",
                          "h2d_bytes_per_step": int(h_idx[0].numel() * 8 + h_off.numel() * 8),
                          "d2h_bytes_per_step": int(host_out.numel() * 4)})
            # ---- per-kernel CUDA-event timing on this rank (roofline pass) ----
            ext.kernel_timing_begin()
            for i in range(min(steps, 20)):
                flush_buf.fill_(i & 0xFF)
                step_eager(i)
            entry["kernel_ms"] = ext.kernel_timing_end()
            entry["tables_per_rank"] = len(local_tables)
            entry["local_nnz"] = NNZ_TABLE * len(local_tables)
            if best is None or ms < res["exchange"][best]["ms_per_step"]:
                best = exch
            if rank == 0:
                sys.stderr.write(f"[bench config4] {exch}: " + repr({k: v for k, v in entry.items() if k != "kernel_ms"}) + "\n")
            del model, fused, graph
            torch.cuda.synchronize()
        except Exception as ex:
            res["errors"][exch] = f"{type(ex).__name__}: {ex}"[:400]
            if rank == 0:
                import traceback

                sys.stderr.write(f"[bench config4] {exch} failed:\n{traceback.format_exc()}\n")
            if world > 1:
                try:
                    torch.cuda.synchronize()
                except Exception:
                    pass
    res["best_exchange"] = best
    if best is not None:
        b = res["exchange"][best]
        res["ms_per_step"] = b["ms_per_step"]
        res["value"] = NNZ_STEP / (b["ms_per_step"] * 1e-3)
        res["e2e_value"] = NNZ_STEP / (b["e2e_ms_per_step"] * 1e-3)
    return res


def cpu_reference_sample(budget_s, max_rows=600_000):
    """--impl reference for config 4: the reference's only CPU-executable path (full_weight() -> embedding_bag ->
    autograd -> SGD; oracle.cpu_reference_step) on the tables small enough to materialise
    (rows <= max_rows), one step each within the budget.  nnz/s = lookups of the sampled tables / their time."""
    import torch

    from oracle import tt_oracle as O

    rng = np.random.RandomState(3)
    R = [1] + RANKS + [1]
    done, secs = 0, 0.0
    tables = sorted((E for E in CARD if int(np.prod(PSHAPE[E])) <= max_rows), key=lambda e: int(np.prod(PSHAPE[e])))
    used = []
    offsets = torch.arange(0, NNZ_TABLE + 1, POOL, dtype=torch.int64)
    grad = torch.rand(B, D) * 0.1
    for E in tables:
        p = PSHAPE[E]
        cores = [torch.randn(1, p[t], R[t] * Q[t] * R[t + 1]) * 0.05 for t in range(3)]
        idx = torch.from_numpy((rng.zipf(ZIPF_A, size=NNZ_TABLE) % E).astype(np.int64))
        t0 = time.perf_counter()
        O.cpu_reference_step(p, Q, RANKS, cores, idx, offsets, grad, LR)
        dt = time.perf_counter() - t0
        secs += dt
        done += NNZ_TABLE
        used.append(E)
        if secs > budget_s:
            break
    return {"value": done / secs, "unit": "nnz/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(),
            "kind": "port",
            "sample": f"{len(used)} of 26 tables (those with <= {max_rows} rows: the path materialises rows x 128 floats "
                      f"per table per step; the 39.9M-row tables would need 20 GB each), one step each, {secs:.1f} s"}, secs
