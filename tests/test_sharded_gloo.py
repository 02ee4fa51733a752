"""CPU, world_size 2, gloo: the table-parallel exchange (fbtt_embedding_b200/sharded.py) -- placement,
packing, forward all-to-all of pooled rows and its mirrored backward -- checked against a single-process
result.  The per-table pooled rows come from the numpy oracle here (no GPU in this container); on the GPU box
the same exchange runs over NCCL with the CUDA tables (tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tt_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tables(T, seed=0):
    rng = np.random.RandomState(seed)
    specs, cores, batches = [], [], []
    B = 8
    for t in range(T):
        p, q, ranks = [3 + t % 3, 4, 5], [2, 2, 3], [3, 2 + t % 2]
        E = int(np.prod(p))
        R = [1] + ranks + [1]
        specs.append(dict(p=p, q=q, ranks=ranks, E=E))
        cores.append([rng.uniform(-1, 1, size=(1, p[i], R[i] * q[i] * R[i + 1])).astype(np.float32) for i in range(3)])
        lens = rng.randint(0, 4, size=B)
        idx = rng.randint(0, E, size=int(lens.sum())).astype(np.int64)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        batches.append((idx, off))
    return specs, cores, batches, B, 12


def _pooled(spec, cores, batch, B, D):
    idx, off = batch
    rowidx, tableidx = O.compute_rowidx(off, 1)
    return O.tt_forward(1, B, D, spec["p"], spec["q"], spec["ranks"], O.make_L(spec["p"]), len(idx), idx, rowidx,
                        tableidx, cores)[0]


def _worker(rank, world, port, T, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fbtt_embedding_b200.sharded import assign_tables, exchange_pooled, tt_lookup_cost

        specs, cores, batches, B, D = _tables(T)
        costs = [tt_lookup_cost(s["q"], s["ranks"], len(b[0])) for s, b in zip(specs, batches)]
        owned = assign_tables(costs, world)
        mine = owned[rank]
        pooled = torch.tensor(np.stack([_pooled(specs[t], cores[t], batches[t], B, D) for t in mine]),
                              requires_grad=True)
        out = exchange_pooled(pooled, owned)  # [B/W, T, D]
        full = np.stack([_pooled(specs[t], cores[t], batches[t], B, D) for t in range(T)])  # [T, B, D]
        bw = B // world
        want = np.transpose(full[:, rank * bw:(rank + 1) * bw, :], (1, 0, 2))
        np.testing.assert_allclose(out.detach().numpy(), want, rtol=0, atol=0)
        # backward: d_out[b, t, :] = (global row id, table id) pattern -> each owner must get its tables' rows
        g = torch.zeros_like(out)
        for b in range(bw):
            for t in range(T):
                g[b, t, :] = 1000.0 * (rank * bw + b) + t
        out.backward(g)
        got = pooled.grad.numpy()
        for li, t in enumerate(mine):
            for b in range(B):
                assert np.all(got[li, b] == 1000.0 * b + t), (rank, t, b)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("T", [2, 5])
def test_table_parallel_exchange_world2(T):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), T, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_assign_tables_balances_by_lookup_cost():
    from fbtt_embedding_b200.sharded import assign_tables

    owned = assign_tables([5, 1, 1, 1, 1, 1, 4, 2], 3)
    assert sorted(t for o in owned for t in o) == list(range(8))
    loads = [sum([5, 1, 1, 1, 1, 1, 4, 2][t] for t in o) for o in owned]
    assert max(loads) - min(loads) <= 1
    assert assign_tables([1.0] * 26, 8) == assign_tables([1.0] * 26, 8)  # deterministic
    assert sorted(len(o) for o in assign_tables([1.0] * 26, 8)) == [3, 3, 3, 3, 3, 3, 4, 4]
