"""fwd + fused SGD bwd at the config-5 shape (E=50M, D=128, q=[4,4,8]) for one rank, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbtt_embedding_b200 import tt_embeddings as ext
r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
p5, q5, R5 = [250, 400, 500], [4, 4, 8], [1, r, r, 1]
E5, D5, B, pool = 50_000_000, 128, 1024, 20
nnz = B * pool
dev = torch.device("cuda:0")
torch.manual_seed(0)
L5 = torch.tensor([p5[1] * p5[2], p5[2], 1], device=dev)
S = [4 * r, r * 4 * r, r * 8]
cs = [((torch.rand(1, p5[i], S[i]) - 0.5) * 0.2).to(dev) for i in range(3)]
e64 = torch.empty(0, dtype=torch.int64, device=dev); e32 = torch.empty(0, dtype=torch.int32, device=dev)
off = torch.arange(0, nnz + 1, pool, device=dev)
go = torch.rand(1, B, D5, device=dev) * 0.1
for i in range(3):
    idx = torch.randint(0, E5, (nnz,), device=dev)
    col, row, tbl, n, _ = ext.preprocess_indices_sync(idx, off, 1, True, e64, e32)
    ext.tt_forward(1000, 1, B, D5, p5, q5, R5, L5, n, col, row, tbl, cs)
    ext.tt_sgd_backward(1000, D5, 0.1, p5, q5, R5, L5, n, col, row, tbl, go, cs)
torch.cuda.synchronize()
print("done")
