#!/usr/bin/env python3
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the dev container only (needs /root/reference; the GPU box does not have
it, which is why the outputs are committed):

    python tests/golden/make_golden.py

What it executes:
  * the reference's own ``tt_matrix_to_full`` (tt_embeddings_ops.py:80-127), imported
    from /root/reference with a stub ``tt_embeddings`` module (the real one is a
    CUDA extension; the function under test is pure torch and CPU-runnable);
  * torch autograd through that function + ``nn.EmbeddingBag`` -- exactly the oracle
    construction of the reference's tests (tt_embeddings_test.py:95-106, 148-172,
    243-246, 317-333);
  * the reference's ``murmor_hash_3_32(int64)`` (hashtbl_cuda_utils.cuh:48-76),
    compiled for the host by nvcc from the header where it lies.

Outputs: tt_golden_T{2,3,4}.npz, hash_kat.json.  Seeds are fixed.
"""
import json
import os
import subprocess
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REF", "/root/reference")


def import_reference_ops():
    stub = types.ModuleType("tt_embeddings")  # the compiled extension is not needed on CPU
    sys.modules["tt_embeddings"] = stub
    sys.path.insert(0, REF)
    import tt_embeddings_ops as ref_ops  # noqa: E402

    sys.path.pop(0)
    return ref_ops


def make_case(ref_ops, T, seed):
    rng = np.random.RandomState(seed)
    p = [7, 9, 11, 5][:T]
    q = [3, 4, 5, 7][:T]
    ranks = [13, 12, 7][: T - 1]
    R = [1] + ranks + [1]
    E, D = int(np.prod(p)), int(np.prod(q))
    cores = [rng.uniform(-0.5, 0.5, size=(1, p[t], R[t] * q[t] * R[t + 1])).astype(np.float32) for t in range(T)]
    cores_t = [torch.tensor(c, requires_grad=True) for c in cores]
    W = ref_ops.tt_matrix_to_full(p, q, R, cores_t, [1, 0, 2, 3])  # reference code, CPU
    assert W.shape == (E, D)
    # a ragged batch with empty bags and duplicate indices
    B = 24
    lens = rng.randint(0, 6, size=B)
    lens[3] = 0
    lens[B - 1] = 0
    nnz = int(lens.sum())
    indices = rng.randint(0, E, size=nnz).astype(np.int64)
    indices[1] = indices[0]  # duplicate inside one bag
    indices[-1] = indices[0]  # and across bags
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    # same two-step construction as tt_embeddings_test.py:151-172
    emb = torch.nn.EmbeddingBag(E, D, sparse=True, mode="sum", _weight=W.detach().clone(), include_last_offset=True)
    out = emb(torch.tensor(indices), torch.tensor(offsets))
    d_out = rng.uniform(-1, 1, size=(B, D)).astype(np.float32)
    out.backward(torch.tensor(d_out))
    W.backward(emb.weight.grad.to_dense())
    grads = [c.grad.numpy().copy() for c in cores_t]
    lr, eps = 0.1, 1e-4
    sgd = [c - lr * g for c, g in zip(cores, grads)]
    state = [g * g for g in grads]
    adagrad = [c - lr * g / (np.sqrt(s) + eps) for c, g, s in zip(cores, grads, state)]
    rows_sel = np.arange(E) if E <= 1000 else np.sort(rng.choice(E, 256, replace=False))
    d = dict(p=np.array(p), q=np.array(q), ranks=np.array(ranks), indices=indices, offsets=offsets,
             d_out=d_out, out=out.detach().numpy(), rows_sel=rows_sel,
             W_rows=W.detach().numpy()[rows_sel], W_sum=np.float64(W.detach().double().sum().item()),
             lr=np.float32(lr), eps=np.float32(eps))
    for t in range(T):
        d[f"core{t}"] = cores[t]
        d[f"grad{t}"] = grads[t]
        d[f"sgd{t}"] = sgd[t].astype(np.float32)
        d[f"state{t}"] = state[t].astype(np.float32)
        d[f"adagrad{t}"] = adagrad[t].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, f"tt_golden_T{T}.npz"), **d)
    print(f"T={T}: E={E} D={D} nnz={nnz} -> tt_golden_T{T}.npz")


def make_hash_kat():
    keys = [0, 1, 2, 12345, 10999999, 11000000, 2**32, 2**32 + 1, 39884405, -1, 2**40 + 77, 2**62 - 1, -(2**33)]
    sizes = [1000, 1048576, 11000000, 7, 2**31 - 1]
    src = r"""
#include <cstdio>
#include <cstdint>
#include "hashtbl_cuda_utils.cuh"
int main(int argc, char** argv) {
  long long keys[] = {%s};
  int sizes[] = {%s};
  for (long long k : keys) for (int c : sizes)
    printf("%%lld %%d %%u\n", k, c, murmor_hash_3_32((int64_t)k, (int32_t)c));
  return 0;
}
""" % (", ".join(f"{k}LL" for k in keys), ", ".join(str(s) for s in sizes))
    tdir = os.path.dirname(torch.__file__)
    with tempfile.TemporaryDirectory() as td:
        cu = os.path.join(td, "kat.cu")
        with open(cu, "w") as f:
            f.write(src)
        exe = os.path.join(td, "kat")
        import sysconfig

        cmd = ["nvcc", "-std=c++17", "-w", "--expt-relaxed-constexpr", "-I", REF, "-I", f"{tdir}/include",
               "-I", f"{tdir}/include/torch/csrc/api/include", "-I", sysconfig.get_paths()["include"],
               cu, "-o", exe, "-L", f"{tdir}/lib", "-lc10", "-ltorch_cpu", "-Xlinker", f"-rpath={tdir}/lib"]
        subprocess.check_call(cmd)
        txt = subprocess.check_output([exe], env=dict(os.environ, LD_LIBRARY_PATH=f"{tdir}/lib")).decode()
    kat = [[int(a), int(b), int(c)] for a, b, c in (ln.split() for ln in txt.strip().splitlines())]
    with open(os.path.join(HERE, "hash_kat.json"), "w") as f:
        json.dump({"source": "reference hashtbl_cuda_utils.cuh:48-76 compiled for host", "kat": kat}, f)
    print(f"hash_kat.json: {len(kat)} vectors")


if __name__ == "__main__":
    torch.manual_seed(0)
    ref_ops = import_reference_ops()
    for T in (2, 3, 4):
        make_case(ref_ops, T, seed=100 + T)
    make_hash_kat()
