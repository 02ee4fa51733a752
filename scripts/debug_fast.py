"""Debug harness for the bucketed tensor-core path (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import tt_oracle as O
from tests.helpers import make_cores
from fbtt_embedding_b200 import tt_embeddings as ext

def t(x): return torch.as_tensor(np.ascontiguousarray(x), device="cuda:0")
p, q, ranks = [20, 22, 25], [4, 4, 4], [32, 32]
R = [1] + ranks + [1]
rng = np.random.RandomState(0)
E = int(np.prod(p)); B = 8; D = 64
cores = make_cores(rng, 1, p, q, ranks)
nnz = int(sys.argv[1]) if len(sys.argv) > 1 else 40
idx = rng.randint(0, E, size=nnz).astype(np.int64)
row = np.sort(rng.randint(0, B, size=nnz)).astype(np.int64)
tbl = np.zeros(nnz, np.int64)
L = O.make_L(p)
want = O.tt_forward(1, B, D, p, q, ranks, L, nnz, idx, row, tbl, cores)
for path in (ext.PATH_GENERIC, ext.PATH_FAST):
    ext.set_path(path)
    out = ext.tt_forward(1000, 1, B, D, p, q, R, t(L), nnz, t(idx), t(row), t(tbl), [t(c) for c in cores])
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    err = np.abs(o - want).max() / np.abs(want).max()
    print("path", path, "rel err", err, "out[0,0,:8]", o[0, 0, :8], "want", want[0, 0, :8])
    if path == ext.PATH_FAST:
        print("plans cached", len(ext._plan_cache))

# ---- backward
dout = rng.uniform(-1, 1, size=(1, B, D)).astype(np.float32)
g_want = O.tt_backward_dense(D, p, q, ranks, L, nnz, idx, row, tbl, dout, cores)
for path in (ext.PATH_GENERIC, ext.PATH_FAST):
    ext.set_path(path)
    g = ext.tt_dense_backward(1000, D, p, q, R, t(L), nnz, t(idx), t(row), t(tbl), t(dout), [t(c) for c in cores])
    torch.cuda.synchronize()
    for i in range(3):
        gi = g[i].cpu().numpy()
        print("bwd path", path, "core", i, "rel err", np.abs(gi - g_want[i]).max() / np.abs(g_want[i]).max(),
              "nonzero", int((gi != 0).sum()), "/", int((g_want[i] != 0).sum()))
