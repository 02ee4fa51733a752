"""fwd + fused SGD bwd at the README shape with a larger batch (steady-state tile pipeline) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbtt_embedding_b200 import tt_embeddings as ext
P, Q, R = [200, 220, 250], [4, 4, 4], [1, 32, 32, 1]
E, D = 11_000_000, 64
B, pool = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 32
nnz = B * pool
dev = torch.device("cuda:0")
torch.manual_seed(0)
L = torch.tensor([P[1] * P[2], P[2], 1], device=dev)
cs = [((torch.rand(1, P[i], [128, 4096, 128][i]) - 0.5) * 0.2).to(dev) for i in range(3)]
e64 = torch.empty(0, dtype=torch.int64, device=dev); e32 = torch.empty(0, dtype=torch.int32, device=dev)
off = torch.arange(0, nnz + 1, pool, device=dev)
go = torch.rand(1, B, D, device=dev) * 0.1
for i in range(5):
    idx = torch.randint(0, E, (nnz,), device=dev)
    col, row, tbl, n, _ = ext.preprocess_indices_sync(idx, off, 1, True, e64, e32)
    ext.tt_forward(1000, 1, B, D, P, Q, R, L, n, col, row, tbl, cs)
    ext.tt_sgd_backward(1000, D, 0.1, P, Q, R, L, n, col, row, tbl, go, cs)
torch.cuda.synchronize()
print("done")
