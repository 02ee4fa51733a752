"""CPU, world_size 2, gloo: the data-parallel step of fbtt_embedding_b200/replicated.py -- bag sharding, the
flat all-reduce and the summed-gradient update -- against the single-process oracle step on the whole batch.
Per-rank dense gradients come from the numpy oracle here (no GPU in this container); on the GPU box the same
host code drives ttb_tt_backward(DENSE) / NCCL / ttb_optimizer_step (tests/test_gpu_parity.py, test_gpu_multi.py)."""
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


def _case(seed=0):
    rng = np.random.RandomState(seed)
    p, q, ranks = [4, 5, 6], [2, 2, 3], [3, 4]
    R = [1] + ranks + [1]
    cores = [rng.uniform(-1, 1, size=(1, p[i], R[i] * q[i] * R[i + 1])).astype(np.float32) for i in range(3)]
    B = 9  # odd on purpose: ranks get 5 and 4 bags
    lens = rng.randint(0, 5, size=B)
    lens[3] = 0
    idx = rng.randint(0, int(np.prod(p)), size=int(lens.sum())).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    d_out = rng.uniform(-1, 1, size=(B, int(np.prod(q)))).astype(np.float32)
    return p, q, ranks, cores, idx, off, d_out


def _dense_grads(p, q, ranks, cores, idx, off, d_out):
    rowidx, tableidx = O.compute_rowidx(off, 1)
    return O.tt_backward_dense(d_out.shape[1], p, q, ranks, O.make_L(p), len(idx), idx, rowidx, tableidx,
                               d_out[None], cores)


def _worker(rank, world, port, optim, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fbtt_embedding_b200.replicated import allreduce_and_step, shard_bags

        p, q, ranks, cores, idx, off, d_out = _case()
        lr, eps = 0.05, 1e-6
        li, lo = shard_bags(torch.from_numpy(idx), torch.from_numpy(off), rank, world)
        b0 = sum(9 // world + (1 if r < 9 % world else 0) for r in range(rank))
        nb = lo.numel() - 1
        local = _dense_grads(p, q, ranks, cores, li.numpy(), lo.numpy(), d_out[b0:b0 + nb])
        flat = torch.from_numpy(np.concatenate([g.ravel() for g in local]).astype(np.float32))
        mine = [c.copy() for c in cores]
        state = [np.zeros_like(c) for c in cores]

        def apply_update():
            summed, at = [], 0
            for c in cores:
                summed.append(flat[at:at + c.size].numpy().reshape(c.shape))
                at += c.size
            if optim == "sgd":
                new = O.sgd_step(mine, summed, lr)
            else:
                new, st = O.adagrad_step(mine, state, summed, lr, eps)
                for a, b in zip(state, st):
                    a[...] = b
            for a, b in zip(mine, new):
                a[...] = b

        allreduce_and_step(flat, apply_update)
        # single-process step on the whole batch
        full = _dense_grads(p, q, ranks, cores, idx, off, d_out)
        if optim == "sgd":
            want = O.sgd_step(cores, full, lr)
        else:
            want, _ = O.adagrad_step(cores, [np.zeros_like(c) for c in cores], full, lr, eps)
        for a, b in zip(mine, want):
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
        # replicas agree bit for bit
        mineflat = torch.from_numpy(np.concatenate([c.ravel() for c in mine]))
        gathered = [torch.empty_like(mineflat) for _ in range(world)]
        dist.all_gather(gathered, mineflat)
        assert all(torch.equal(gathered[0], g) for g in gathered)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("optim", ["sgd", "adagrad"])
def test_replicated_step_world2(optim):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), optim, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_bags_covers_the_batch_once():
    from fbtt_embedding_b200.replicated import shard_bags

    _, _, _, _, idx, off, _ = _case(3)
    for world in (1, 2, 3, 4, 9, 12):
        got_idx, bags = [], 0
        for r in range(world):
            li, lo = shard_bags(torch.from_numpy(idx), torch.from_numpy(off), r, world)
            assert int(lo[0]) == 0 and int(lo[-1]) == li.numel()
            got_idx.append(li.numpy())
            bags += lo.numel() - 1
        assert bags == off.size - 1
        assert np.array_equal(np.concatenate(got_idx), idx)
