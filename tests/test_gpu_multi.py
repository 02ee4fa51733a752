"""Multi-GPU (>= 2 devices, NCCL): the table-parallel module end to end on real CUDA tables -- forward
all-to-all of pooled rows, mirrored backward exchange, fused SGD on the owners -- against a single-GPU run
of the same tables.  Skipped on 1-GPU boxes."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _specs(T):
    out = []
    for t in range(T):
        p = [10 + t, 12, 14]
        out.append(dict(num_embeddings=int(np.prod(p)), embedding_dim=64, tt_ranks=[32, 32], tt_p_shapes=p,
                        tt_q_shapes=[4, 4, 4]))
    return out


def _batch(T, B, seed):
    rng = np.random.RandomState(seed)
    idx, off = [], []
    for t in range(T):
        lens = rng.randint(0, 6, size=B)
        E = _specs(T)[t]["num_embeddings"]
        idx.append(rng.randint(0, E, size=int(lens.sum())).astype(np.int64))
        off.append(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
    return idx, off


def _spawn(fn, args, world, timeout_s=300):
    """mp.spawn with a deadline: ranks stuck in a collective (or a rank barrier of the fused exchange) are killed
    instead of hanging the box."""
    import time

    import torch.multiprocessing as mp

    ctx = mp.spawn(fn, args=args, nprocs=world, join=False)
    t0 = time.time()
    while not ctx.join(timeout=5):
        if time.time() - t0 > timeout_s:
            for p in ctx.processes:
                if p.is_alive():
                    p.kill()
            pytest.fail(f"ranks still running after {timeout_s} s: killed")


def _worker(rank, world, port, T, B, ret, mode="nccl"):
    """mode: "nccl" = one module per table + all_to_all_single; "fused" = the rank's tables as one fused
    heterogeneous batch + all_to_all_single; "peer" = fused batch with the exchange folded into the kernels over
    symmetric (peer-mapped) memory."""
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fbtt_embedding_b200 import OptimType, TTEmbeddingBag
        from fbtt_embedding_b200.sharded import TableShardedTTEmbeddingBag

        dev = torch.device("cuda", rank)
        specs = _specs(T)
        idx, off = _batch(T, B, 3)
        torch.manual_seed(0)
        kw = dict(optimizer=OptimType.SGD, learning_rate=0.1, sparse=True, weight_dist="uniform")
        extra = {} if mode == "nccl" else dict(fused=True, exchange="peer" if mode == "peer" else "nccl")
        model = TableShardedTTEmbeddingBag(specs, [len(i) for i in idx], **extra, **kw)
        # identical weights everywhere: table t's cores are seeded by t
        ref_tables = []
        for t in range(T):
            g = torch.Generator(device="cpu").manual_seed(100 + t)
            ref_tables.append([torch.rand(1, specs[t]["tt_p_shapes"][i], [128, 4096, 128][i], generator=g) - 0.5
                               for i in range(3)])
        with torch.no_grad():
            if mode == "nccl":
                for tbl, t in zip(model.tables, model.local_tables):
                    for i in range(3):
                        tbl.tt_cores[i].copy_(ref_tables[t][i])
            else:
                for k, t in enumerate(model.local_tables):
                    model.fused.load_table(k, [c.to(dev) for c in ref_tables[t]])
        li = [torch.as_tensor(idx[t], device=dev) for t in model.local_tables]
        lo = [torch.as_tensor(off[t], device=dev) for t in model.local_tables]
        out = model(li, lo)  # [B/W, T, D]
        bw = B // world
        g_full = torch.rand(B, T, 64, generator=torch.Generator().manual_seed(5)) * 0.1
        out.backward(g_full[rank * bw:(rank + 1) * bw].to(dev))
        # single-GPU reference of every table on this rank's device
        worst = 0.0
        for t in range(T):
            single = TTEmbeddingBag(**specs[t], use_cache=False, **kw)
            with torch.no_grad():
                for i in range(3):
                    single.tt_cores[i].copy_(ref_tables[t][i])
            so = single(torch.as_tensor(idx[t], device=dev), torch.as_tensor(off[t], device=dev))
            want = so[rank * bw:(rank + 1) * bw]
            worst = max(worst, float((out[:, t] - want).abs().max() / want.abs().max().clamp_min(1e-9)))
            so.backward(g_full[:, t].to(dev))
            if t in model.local_tables:
                k = model.local_tables.index(t)
                mine = list(model.tables[k].tt_cores) if mode == "nccl" else model.fused.table_cores(k)
                for i in range(3):
                    d = (mine[i] - single.tt_cores[i]).abs().max() / single.tt_cores[i].abs().max()
                    worst = max(worst, float(d))
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", ["nccl", "fused", "peer"])
def test_table_sharded_two_gpus(mode):
    import torch.multiprocessing as mp

    world, T, B = 2, 5, 32
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn(_worker, (world, _free_port(), T, B, ret, mode), world)
    assert set(ret.keys()) == {0, 1}
    assert max(ret.values()) < 2e-3, dict(ret)  # tf32 path on both sides; atomics order differs


def _replica_worker(rank, world, port, optimizer, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from fbtt_embedding_b200 import OptimType, TTEmbeddingBag
        from fbtt_embedding_b200.replicated import ReplicatedTTEmbeddingBag, shard_bags

        dev = torch.device("cuda", rank)
        p, q, ranks = [20, 22, 25], [4, 4, 4], [32, 32]
        E, D, B = int(np.prod(p)), 64, 96
        kw = dict(tt_p_shapes=p, tt_q_shapes=q, tt_ranks=ranks, optimizer=getattr(OptimType, optimizer),
                  learning_rate=0.1, eps=1e-4, weight_dist="uniform")
        torch.manual_seed(rank)  # replicas start different on purpose: sync_replicas must fix that
        rep = ReplicatedTTEmbeddingBag(E, D, **kw)
        single = TTEmbeddingBag(E, D, use_cache=False, sparse=True, **kw)
        with torch.no_grad():
            for a, b in zip(single.tt_cores, rep.table.tt_cores):
                a.copy_(b)
        rng = np.random.RandomState(11)
        worst = 0.0
        for step in range(3):
            lens = rng.randint(0, 7, size=B)
            idx = torch.as_tensor(rng.randint(0, E, size=int(lens.sum())).astype(np.int64))
            off = torch.as_tensor(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
            g = torch.rand(B, D, generator=torch.Generator().manual_seed(step)) * 0.1
            li, lo = shard_bags(idx, off, rank, world)
            b0 = rank * (B // world)
            out = rep(li.to(dev), lo.to(dev))
            want = single(idx.to(dev), off.to(dev))
            worst = max(worst, float((out - want[b0:b0 + B // world]).abs().max() / want.abs().max()))
            out.backward(g[b0:b0 + B // world].to(dev))
            want.backward(g.to(dev))
            for a, b in zip(rep.table.tt_cores, single.tt_cores):
                worst = max(worst, float((a - b).abs().max() / b.abs().max()))
        # replicas identical bit for bit
        flat = torch.cat([c.detach().flatten() for c in rep.table.tt_cores])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert all(torch.equal(both[0], x) for x in both)
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("optimizer", ["SGD", "EXACT_ADAGRAD"])
def test_replicated_two_gpus_equal_one_gpu_on_the_whole_batch(optimizer):
    import torch.multiprocessing as mp

    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    _spawn(_replica_worker, (world, _free_port(), optimizer, ret), world)
    assert set(ret.keys()) == {0, 1}
    assert max(ret.values()) < (2e-3 if optimizer == "SGD" else 2e-2), dict(ret)
