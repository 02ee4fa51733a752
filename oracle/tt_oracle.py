"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy restatement of the reference algorithm for the TT-EmbeddingBag
hot path (facebookresearch/FBTT-Embedding @ b95947c).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this file; the product (``fbtt_embedding_b200``)
never does and fails loudly when its CUDA library is missing.

Parity pinning: every function below is checked in ``tests/test_oracle.py``
against golden vectors produced by *importing the reference's own Python*
(``tt_matrix_to_full`` + torch autograd, ``tests/golden/make_golden.py``) and
against hash known-answers produced by compiling the reference's
``hashtbl_cuda_utils.cuh`` for the host.  On the GPU box the reference CUDA
extension built by ``oracle/build_ref.sh`` (``oracle/_ref``) is the second,
stronger checker (``tests/test_gpu_vs_reference.py``).

Each function cites the reference file:line it follows (paths relative to the
reference checkout).  Layout facts used everywhere:

* core t is stored ``[num_tables, p_t, r_t*q_t*r_{t+1}]`` fp32; one slice is a
  row-major ``r_t x (q_t*r_{t+1})`` matrix (tt_embeddings_ops.py:513-530 and
  the ``[1,0,2,3]`` permute in :601-611 / :80-127).
* ``L[t] = prod_{s>t} p_s`` (tt_embeddings_ops.py:506-512); index digits are
  ``i_t = idx // L[t]; idx %= L[t]`` (tt_embeddings_cuda.cu:795-799).
"""
from __future__ import annotations

import numpy as np

UNUSED_KEY = -1
MAX_PROBES = 3  # tt_embeddings_cuda.cu:29


# --------------------------------------------------------------------------
# shapes / index decomposition
# --------------------------------------------------------------------------
def make_L(p):
    """tt_embeddings_ops.py:506-512."""
    L, v = [], 1
    for t in range(len(p)):
        L.append(v)
        v *= int(p[len(p) - t - 1])
    L.reverse()
    return np.asarray(L, dtype=np.int64)


def decompose(indices, L):
    """tt_embeddings_cuda.cu:795-799 (and :146-150 for backward): mixed radix digits."""
    idx = np.asarray(indices, dtype=np.int64).copy()
    digits = []
    for t in range(len(L)):
        digits.append(idx // L[t])
        idx = idx % L[t]
    return digits  # list of int64 arrays, one per core


def full_ranks(ranks, T):
    r = [int(x) for x in ranks]
    if len(r) == T - 1:
        r = [1] + r + [1]
    assert len(r) == T + 1
    return r


# --------------------------------------------------------------------------
# dense expansion (the reference's own CPU-executable path)
# --------------------------------------------------------------------------
def tt_matrix_to_full(p, q, ranks, cores, dtype=np.float32):
    """Restates tt_embeddings_ops.py:80-127 with tt_permute=[1,0,2,3] for a
    single table: cores[t] is [p_t, r_t*q_t*r_{t+1}] (table dim squeezed)."""
    T = len(p)
    R = full_ranks(ranks, T)
    cs = []
    for t in range(T):
        c = np.asarray(cores[t], dtype=dtype).reshape(p[t], R[t], q[t], R[t + 1])
        cs.append(np.ascontiguousarray(c.transpose(1, 0, 2, 3)))  # [r_t, p_t, q_t, r_{t+1}]
    res = cs[0]
    for t in range(1, T):
        res = res.reshape(-1, R[t]) @ cs[t].reshape(R[t], -1)
    inter = []
    for t in range(T):
        inter += [p[t], q[t]]
    res = res.reshape(*inter)
    perm = list(range(0, 2 * T, 2)) + list(range(1, 2 * T, 2))
    n_dim = int(np.prod(p))
    k_dim = int(np.prod(q))
    return np.ascontiguousarray(res.transpose(*perm)).reshape(n_dim, k_dim).astype(np.float32)


# --------------------------------------------------------------------------
# per-index chain (what the CUDA path computes)
# --------------------------------------------------------------------------
def tt_rows(p, q, ranks, L, indices, tableidx, cores, dtype=np.float32, keep=False):
    """Rows W[idx] for every lookup via the chain of small GEMMs
    (tt_embeddings_cuda.cu:993-1004 dims, :1039-1055 GEMMs).  cores[t] is
    [num_tables, p_t, S_t].  Returns [nnz, D] (and the intermediates v_t if keep)."""
    T = len(p)
    R = full_ranks(ranks, T)
    digits = decompose(indices, L)
    tbl = np.asarray(tableidx, dtype=np.int64)
    n = len(tbl)
    v = np.asarray(cores[0], dtype=dtype)[tbl, digits[0]].reshape(n, q[0], R[1])
    inter = [v]
    for t in range(1, T):
        c = np.asarray(cores[t], dtype=dtype)[tbl, digits[t]].reshape(n, R[t], q[t] * R[t + 1])
        v = np.matmul(v, c).reshape(n, -1, R[t + 1])
        inter.append(v)
    rows = v.reshape(n, -1)
    return (rows, inter, digits) if keep else rows


def tt_forward(num_tables, B, D, p, q, ranks, L, nnz, indices, rowidx, tableidx, cores,
               dtype=np.float32):
    """tt_embeddings_cuda.cu:964-1075: output[t,b,:] = sum over lookups of the
    bag, accumulated sequentially in index order (reduce_output_kernel :943-961)."""
    out = np.zeros((num_tables, B, D), dtype=dtype)
    if nnz == 0:
        return out
    rows = tt_rows(p, q, ranks, L, indices[:nnz], tableidx[:nnz], cores, dtype)
    np.add.at(out, (np.asarray(tableidx[:nnz]), np.asarray(rowidx[:nnz])), rows)
    return out


def tt_backward_dense(D, p, q, ranks, L, nnz, indices, rowidx, tableidx, d_output, cores,
                      dtype=np.float64):
    """Gradient of every core, dense, core-shaped -- tt_embeddings_cuda.cu:419-652
    with optim == OPTIM_DENSE (K5 recompute, K6 dCore, K7 scatter-add, K8 dPrev).
    Default fp64 accumulation: the reference's atomic order is undefined, so the
    oracle gives the exactly-rounded sum rather than one arbitrary order."""
    T = len(p)
    R = full_ranks(ranks, T)
    grads = [np.zeros(np.asarray(c).shape, dtype=dtype) for c in cores]
    if nnz == 0:
        return grads
    idx = np.asarray(indices[:nnz])
    tbl = np.asarray(tableidx[:nnz], dtype=np.int64)
    row = np.asarray(rowidx[:nnz], dtype=np.int64)
    rows, inter, digits = tt_rows(p, q, ranks, L, idx, tbl, cores, dtype, keep=True)
    n = len(idx)
    dv = np.asarray(d_output, dtype=dtype)[tbl, row].reshape(n, -1, 1)  # [n, m_{T-1}, r_T=1]
    for t in range(T - 1, 0, -1):
        c = np.asarray(cores[t], dtype=dtype)[tbl, digits[t]].reshape(n, R[t], q[t] * R[t + 1])
        prev = inter[t - 1]  # [n, m_{t-1}, r_t]
        dvm = dv.reshape(n, prev.shape[1], q[t] * R[t + 1])
        dcore = np.matmul(prev.transpose(0, 2, 1), dvm)  # [n, r_t, q_t r_{t+1}]   (K6)
        np.add.at(grads[t], (tbl, digits[t]), dcore.reshape(n, -1))  # (K7)
        dv = np.matmul(dvm, c.transpose(0, 2, 1))  # [n, m_{t-1}, r_t]          (K8)
    np.add.at(grads[0], (tbl, digits[0]), dv.reshape(n, -1))
    return grads


def sgd_step(cores, grads, lr):
    """update_tt_cores_sgd_kernel, tt_embeddings_cuda.cu:392 -- applied to EVERY row
    (the reference launch skips rows when p_t > S_t; the oracle follows
    the mathematical definition the reference's own test_backward_sgd asserts,
    tt_embeddings_test.py:243-246)."""
    return [(np.asarray(c, np.float32) - np.float32(lr) * np.asarray(g, np.float32)).astype(np.float32)
            for c, g in zip(cores, grads)]


def adagrad_step(cores, states, grads, lr, eps):
    """update_tt_cores_adagrad_kernel, tt_embeddings_cuda.cu:412-414."""
    new_c, new_s = [], []
    for c, s, g in zip(cores, states, grads):
        g = np.asarray(g, np.float32)
        s2 = (np.asarray(s, np.float32) + g * g).astype(np.float32)
        c2 = np.asarray(c, np.float32) - np.float32(lr) * g / (np.sqrt(s2) + np.float32(eps))
        new_c.append(c2.astype(np.float32))
        new_s.append(s2)
    return new_c, new_s


# --------------------------------------------------------------------------
# CSR -> COO (preprocess_indices_sync, warm-up branch)
# --------------------------------------------------------------------------
def compute_rowidx(offsets, num_tables):
    """compute_rowidx_kernel, tt_embeddings_cuda.cu:1338-1354."""
    offsets = np.asarray(offsets, dtype=np.int64)
    nb = len(offsets) - 1
    B = nb // num_tables
    lens = np.diff(offsets)
    bag = np.repeat(np.arange(nb, dtype=np.int64), lens)
    nnz = int(offsets[-1] - offsets[0])
    rowidx = np.zeros(nnz, np.int64)
    tableidx = np.zeros(nnz, np.int64)
    rowidx[:] = bag % B
    tableidx[:] = bag // B
    return rowidx, tableidx


# --------------------------------------------------------------------------
# hash table / LFU cache (integer state, bit-exact contract)
# --------------------------------------------------------------------------
def _rotl32(x, r):
    x = np.asarray(x, dtype=np.uint32)
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def murmur_hash_3_32_i64(keys, C):
    """hashtbl_cuda_utils.cuh:48-76 (int64 overload): seed 0, low word then high
    word, ``h ^= 2``, fmix32, then Lemire multiply-shift ``(u64)h*C >> 32``."""
    with np.errstate(over="ignore"):
        k = np.asarray(keys, dtype=np.int64).view(np.uint64)
        c1 = np.uint32(0xCC9E2D51)
        c2 = np.uint32(0x1B873593)
        h = np.zeros(k.shape, dtype=np.uint32)
        for word in ((k & np.uint64(0xFFFFFFFF)).astype(np.uint32), (k >> np.uint64(32)).astype(np.uint32)):
            k1 = (word * c1).astype(np.uint32)
            k1 = _rotl32(k1, 15)
            k1 = (k1 * c2).astype(np.uint32)
            h = h ^ k1
            h = _rotl32(h, 13)
            h = (h * np.uint32(5) + np.uint32(0xE6546B64)).astype(np.uint32)
        h = h ^ np.uint32(2)
        h = h ^ (h >> np.uint32(16))
        h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
        h = h ^ (h >> np.uint32(13))
        h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
        h = h ^ (h >> np.uint32(16))
        return ((h.astype(np.uint64) * np.uint64(C)) >> np.uint64(32)).astype(np.int64)


def hashtbl_find(key, hashtbl):
    """hashtbl_cuda_utils.cuh:135-154.  NB: the early exit tests the
    *search key* against UNUSED_KEY, so an empty slot never stops the probe."""
    C = len(hashtbl)
    slot = int(murmur_hash_3_32_i64(np.int64(key), C))
    for _ in range(MAX_PROBES):
        if hashtbl[slot] == key:
            return slot
        if key == UNUSED_KEY:
            return -1
        slot = (slot + 1) % C
    return -1


def update_cache_state(indices, hashtbl, cache_freq):
    """update_cache_state_kernel + hashtbl_insert<accumulate=true>,
    tt_embeddings_cuda.cu:1077-1089, hashtbl_cuda_utils.cuh:102-133, executed in
    index order (one legal serialisation of the racing CAS).  In place."""
    C = len(hashtbl)
    homes = murmur_hash_3_32_i64(np.asarray(indices, np.int64), C)
    dropped = []
    for key, slot in zip(np.asarray(indices, np.int64).tolist(), homes.tolist()):
        placed = False
        for _ in range(MAX_PROBES):
            old = hashtbl[slot]
            if old == UNUSED_KEY:
                hashtbl[slot] = key
                old = UNUSED_KEY
            if old == UNUSED_KEY or old == key:
                cache_freq[slot] += 1
                placed = True
                break
            slot = (slot + 1) % C
        if not placed:
            dropped.append(key)
    return dropped


def cache_populate_state(cache_size, hashtbl, cache_freq, cache_state):
    """Integer half of cache_populate_cuda, tt_embeddings_cuda.cu:1260-1324:
    stable descending sort of slots by frequency carrying the key (K12), then
    mark_popular_colidx_kernel (:1115-1139).  In place; returns the sorted keys
    (first cache_size entries are the cached rows, empty ones replaced by 0)."""
    H = len(hashtbl)
    order = np.argsort(-cache_freq.astype(np.int64), kind="stable")
    sorted_keys = hashtbl[order].copy()
    snapshot = sorted_keys.copy()
    for n in range(H):
        key = int(snapshot[n])
        if key != UNUSED_KEY:
            slot = hashtbl_find(key, hashtbl)
            if n < cache_size:
                cache_state[slot] = n
            else:
                hashtbl[slot] = UNUSED_KEY
                cache_freq[slot] = 0
        elif n < cache_size:
            sorted_keys[n] = 0  # "a hack to use batch gemm", :1135-1138
    return sorted_keys


def cache_lookup(colidx, hashtbl, cache_state):
    """cache_lookup_kernel, tt_embeddings_cuda.cu:1356-1375."""
    n = len(colidx)
    is_tt = np.ones(n, dtype=bool)
    loc = np.zeros(n, dtype=np.int32)
    for i, key in enumerate(np.asarray(colidx, np.int64).tolist()):
        slot = hashtbl_find(key, hashtbl)
        if slot != -1 and cache_state[slot] != -1:
            is_tt[i] = False
            loc[i] = cache_state[slot]
    return is_tt, loc


def partition_flagged(x, flags):
    """cub::DevicePartition::Flagged semantics used at tt_embeddings_cuda.cu:1436-1479:
    selected items first in order, rejected items at the tail in REVERSE order."""
    x = np.asarray(x)
    return np.concatenate([x[flags], x[~flags][::-1]])


def preprocess_indices(colidx, offsets, num_tables, warmup, hashtbl, cache_state):
    """preprocess_indices_sync_cuda, tt_embeddings_cuda.cu:1377-1496."""
    rowidx, tableidx = compute_rowidx(offsets, num_tables)
    if len(rowidx) == 0 or warmup or num_tables != 1:
        return np.asarray(colidx, np.int64), rowidx, tableidx, len(rowidx), None
    is_tt, loc = cache_lookup(colidx, hashtbl, cache_state)
    return (partition_flagged(colidx, is_tt), partition_flagged(rowidx, is_tt), tableidx,
            int(is_tt.sum()), partition_flagged(loc, is_tt))


def cache_forward(cache_locations, rowidx, cache_weight, output):
    """cache_forward_kernel, tt_embeddings_cuda.cu:1498-1538.  output [B,D] in place."""
    np.add.at(output, np.asarray(rowidx, np.int64), cache_weight[np.asarray(cache_locations, np.int64)])


def cache_backward_sgd(grad_output, cache_locations, rowidx, lr, cache_weight):
    """cache_backward_sgd_kernel, tt_embeddings_cuda.cu:1574-1621.  In place."""
    g = (-np.asarray(grad_output, np.float32)[np.asarray(rowidx, np.int64)] * np.float32(lr)).astype(np.float32)
    np.add.at(cache_weight, np.asarray(cache_locations, np.int64), g)


def cache_backward_dense(grad_output, cache_locations, rowidx, cache_weight):
    """cache_backward_dense_kernel, tt_embeddings_cuda.cu:1659-1697."""
    out = np.zeros_like(cache_weight)
    np.add.at(out, np.asarray(cache_locations, np.int64), np.asarray(grad_output, np.float32)[np.asarray(rowidx, np.int64)])
    return out


def cache_backward_rowwise_adagrad_approx(grad_output, cache_locations, rowidx, lr, eps,
                                          state, cache_weight):
    """cache_backward_rowwise_adagrad_approx_kernel, tt_embeddings_cuda.cu:1735-1795,
    serialised in index order (duplicates race in the reference).  In place."""
    go = np.asarray(grad_output, np.float32)
    D = go.shape[1]
    for loc, r in zip(np.asarray(cache_locations).tolist(), np.asarray(rowidx).tolist()):
        g = go[r]
        g2 = np.float32(np.sum(g * g, dtype=np.float32) / np.float32(D))
        old = state[loc]
        state[loc] = old + g2
        mult = np.float32(lr) * np.float32(1.0 / (np.sqrt(np.float32(old + g2)) + np.float32(eps)))
        cache_weight[loc] -= g * mult


# --------------------------------------------------------------------------
# the reference's CPU-executable training step, torch on host
# --------------------------------------------------------------------------
def cpu_reference_step(p, q, ranks, cores_t, indices_t, offsets_t, grad_out_t, lr):
    """full_weight() -> embedding_bag(sum) -> backward through the expansion -> SGD.
    cores_t: list of torch CPU tensors [1,p_t,S_t] (updated in place).  This is the
    path the reference's own tests use as their oracle (tt_embeddings_test.py:95-106,
    161-172, 243-246) restated with torch CPU ops (the reference's function is pure
    torch too, tt_embeddings_ops.py:80-127)."""
    import torch

    T = len(p)
    R = full_ranks(ranks, T)
    leaves = [c.detach().clone().requires_grad_(True) for c in cores_t]
    cs = [leaves[t].view(p[t], R[t], q[t], R[t + 1]).permute(1, 0, 2, 3).contiguous() for t in range(T)]
    res = cs[0]
    for t in range(1, T):
        res = torch.matmul(res.view(-1, R[t]), cs[t].view(R[t], -1))
    inter = []
    for t in range(T):
        inter += [p[t], q[t]]
    perm = list(range(0, 2 * T, 2)) + list(range(1, 2 * T, 2))
    W = res.view(*inter).permute(*perm).contiguous().view(int(np.prod(p)), int(np.prod(q)))
    out = torch.nn.functional.embedding_bag(indices_t, W, offsets_t, mode="sum", include_last_offset=True)
    out.backward(grad_out_t)
    with torch.no_grad():
        for c, leaf in zip(cores_t, leaves):
            c -= lr * leaf.grad
    return out.detach()
