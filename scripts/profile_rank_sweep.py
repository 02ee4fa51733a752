"""Config-5 shape (E=50M -> p=[250,400,500], D=128, q=[4,4,8], B=1024, pooling 20), fwd + fused SGD bwd, for ncu:
one kernel family per rank -- 8: generic fp32 FFMA, 16: warp-level mma.sync tf32, 128: tcgen05 (64 is captured with
config 4, 32 with the README shape).  Ranks from argv, default 8 16 128."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fbtt_embedding_b200 import tt_embeddings as ext

ranks = [int(a) for a in sys.argv[1:]] or [8, 16, 128]
p5, q5 = [250, 400, 500], [4, 4, 8]
E5, D5, B, pool = 50_000_000, 128, 1024, 20
nnz = B * pool
dev = torch.device("cuda:0")
L5 = torch.tensor([p5[1] * p5[2], p5[2], 1], device=dev, dtype=torch.int64)
e64 = torch.empty(0, dtype=torch.int64, device=dev)
e32 = torch.empty(0, dtype=torch.int32, device=dev)
off = torch.arange(0, nnz + 1, pool, device=dev)
go = torch.rand(1, B, D5, device=dev) * 0.1
for r in ranks:
    R5 = [1, r, r, 1]
    S = [4 * r, r * 4 * r, r * 8]
    g = torch.Generator(device="cpu").manual_seed(r)
    cs = [((torch.rand(1, p5[i], S[i], generator=g) - 0.5) * 0.2).to(dev) for i in range(3)]
    for i in range(2):
        idx = torch.randint(0, E5, (nnz,), device=dev)
        col, row, tbl, n, _ = ext.preprocess_indices_sync(idx, off, 1, True, e64, e32)
        ext.tt_forward(1000, 1, B, D5, p5, q5, R5, L5, n, col, row, tbl, cs)
        ext.tt_sgd_backward(1000, D5, 0.1, p5, q5, R5, L5, n, col, row, tbl, go, cs)
    torch.cuda.synchronize()
    print("rank", r, "done", flush=True)
