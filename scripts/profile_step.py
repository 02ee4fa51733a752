"""Minimal fwd + fused-SGD-bwd loop at the BASELINE shape for ncu (no timing, few steps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbtt_embedding_b200 import OptimType, TTEmbeddingBag
from fbtt_embedding_b200 import tt_embeddings as ext
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
path = sys.argv[2] if len(sys.argv) > 2 else "auto"
ext.set_path({"auto": ext.PATH_AUTO, "generic": ext.PATH_GENERIC}[path])
torch.manual_seed(0)
E, D, B, POOL = 11_000_000, 64, 512, 20
emb = TTEmbeddingBag(E, D, [32, 32], [200, 220, 250], [4, 4, 4], optimizer=OptimType.SGD, learning_rate=0.1,
                     sparse=True, use_cache=False, weight_dist="uniform")
off = torch.arange(0, B * POOL + 1, POOL, device="cuda")
g = torch.rand(B, D, device="cuda") * 0.1
for i in range(steps):
    idx = torch.randint(0, E, (B * POOL,), device="cuda")
    emb(idx, off).backward(g)
torch.cuda.synchronize()
print("done", ext.launch_count())
