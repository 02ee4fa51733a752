"""BASELINE config 4: DLRM-style 26 tables (Criteo-Terabyte cardinalities, synthetic zipf indices), D=128,
ranks [64,64], q=[4,4,8], B=4096, one-hot (L=1), table-sharded over the ranks of one box with ONE
all_to_all of pooled rows forward and its mirror backward (fbtt_embedding_b200/sharded.py).
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_config4.py
Prints one JSON line on rank 0 (whole-job nnz/s = 26*B lookups per step / max-over-ranks step time)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

CARD = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10, 2208, 11938, 155, 4,
        976, 14, 39979771, 25641295, 39664984, 585935, 12972, 108, 36]
# p-shapes from the reference's suggested_tt_shapes(E, 3)
PSHAPE = {39884406: [304, 350, 375], 38532951: [320, 325, 375], 39979771: [334, 342, 350], 39664984: [250, 397, 400],
          25641295: [285, 300, 300], 2953546: [130, 142, 160], 585935: [75, 80, 100], 403346: [50, 82, 100],
          39043: [25, 40, 40], 20263: [25, 28, 29], 17289: [24, 25, 30], 12972: [20, 25, 26], 11938: [20, 24, 25],
          7420: [20, 20, 20], 7120: [20, 20, 20], 2208: [10, 13, 17], 1543: [10, 10, 16], 976: [10, 10, 10],
          155: [5, 5, 8], 108: [5, 5, 8], 63: [3, 3, 7], 36: [3, 3, 4], 14: [2, 2, 5], 10: [1, 2, 5], 4: [1, 2, 2],
          3: [1, 1, 3]}
D, RANKS, Q, B, POOL = 128, [64, 64], [4, 4, 8], 4096, int(os.environ.get("POOL", "1"))


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from fbtt_embedding_b200 import OptimType
    from fbtt_embedding_b200.sharded import TableShardedTTEmbeddingBag

    steps, warm = int(os.environ.get("STEPS", "30")), 5
    specs = [dict(num_embeddings=E, embedding_dim=D, tt_ranks=RANKS, tt_p_shapes=PSHAPE[E], tt_q_shapes=Q) for E in CARD]
    torch.manual_seed(0)
    grouped = os.environ.get("CFG4_GROUPED", "0") == "1"  # one host call per phase for the rank's tables
    lanes = int(os.environ.get("CFG4_LANES", "1"))          # internal streams the group spreads its tables over
    fused = os.environ.get("CFG4_FUSED", "0") == "1"      # one plan / forward / backward / sweep launch for the rank's tables
    exchange = os.environ.get("CFG4_EXCHANGE", "nccl")    # "peer": all-to-all folded into the kernels (needs CFG4_FUSED=1)
    if grouped:
        from fbtt_embedding_b200 import tt_embeddings as ext

        ext.group_set_streams(lanes)
    model = TableShardedTTEmbeddingBag(specs, [B * POOL] * len(CARD), grouped=grouped, fused=fused, exchange=exchange, optimizer=OptimType.SGD,
                                       learning_rate=0.1, sparse=True, weight_dist="uniform")
    rng = np.random.RandomState(1)  # same stream on every rank: replicated synthetic inputs
    nnz = B * POOL
    batches = []
    for _ in range(4):
        idx = [torch.as_tensor((rng.zipf(1.2, size=nnz) % E).astype(np.int64)) for E in CARD]
        batches.append([idx[t].to(dev) for t in model.local_tables])
    off = torch.arange(0, nnz + 1, POOL, device=dev)
    offs = [off] * len(model.local_tables)
    if fused:  # table-major (keyed-jagged) inputs, packed once outside the timed region like the index generation
        from fbtt_embedding_b200.fused import pack_table_major

        packed = [pack_table_major(b, offs) for b in batches]
        batches = [p[0] for p in packed]
        offs = packed[0][1]
    g = torch.rand(B // world, len(CARD), D, device=dev) * 0.1

    def step(i):
        out = model(batches[i % 4], offs)
        out.backward(g)

    def timed(fn):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    eager_ms = timed(step)
    # the same step (26 module calls + both all-to-alls) captured once in a CUDA graph: removes the
    # per-table Python / launch overhead, which dominates at 4096 lookups per table
    graph_ms = None
    try:
        if os.environ.get("CFG4_GRAPH", "0") != "1":  # opt-in: capturing NCCL inside the graph hung once on 2 GPUs
            raise RuntimeError("graph mode not requested (set CFG4_GRAPH=1)")
        static = batches[0].clone() if fused else [b.clone() for b in batches[0]]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                model(static, offs).backward(g)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            model(static, offs).backward(g)
        torch.cuda.synchronize()

        def step_graph(i):
            if fused:
                static.copy_(batches[i % 4])
            else:
                for dst, src in zip(static, batches[i % 4]):
                    dst.copy_(src)
            gr.replay()

        graph_ms = timed(step_graph)
    except Exception as ex:  # pragma: no cover
        if rank == 0 and os.environ.get("CFG4_GRAPH", "0") == "1":
            sys.stderr.write(f"[cfg4] graph capture unavailable: {type(ex).__name__}: {ex}\n")
    ms = min(x for x in (eager_ms, graph_ms) if x is not None)
    if rank == 0:
        tot = len(CARD) * nnz
        print(json.dumps({"config": "cfg4_dlrm26", "n_gpus": world, "B": B, "pooling": POOL, "tables": len(CARD),
                          "nnz_per_step": tot, "ms_per_step": float(ms), "nnz_per_s": tot / float(ms) * 1e3,
                          "eager_ms_per_step": eager_ms, "graph_ms_per_step": graph_ms,
                          "tables_per_rank": [len(o) for o in model.owned], "grouped": grouped, "lanes": lanes, "fused": fused, "exchange": exchange,
                          "a2a_bytes_per_rank_fwd": (len(model.local_tables) * B * D * 4),
                          "timing": "CUDA events around %d eager steps, max over ranks" % steps}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
