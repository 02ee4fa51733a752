"""Phase trace of the tcgen05 forward / backward kernels at the README shape: per-CTA %globaltimer stamps
(ttb_trace_set) -> when each phase starts relative to the first CTA's start, medians over CTAs.
    python scripts/trace_phases.py [nnz] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fbtt_embedding_b200 import OptimType, TTEmbeddingBag
from fbtt_embedding_b200 import tt_embeddings as ext

nnz = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
emb = TTEmbeddingBag(11_000_000, 64, [32, 32], [200, 220, 250], [4, 4, 4], optimizer=OptimType.SGD, learning_rate=0.1,
                     sparse=True, use_cache=False, weight_dist="uniform")
idx = torch.randint(0, 11_000_000, (nnz,), device=dev)
off = torch.arange(0, nnz + 1, nnz // B, device=dev)
g = torch.rand(B, 64, device=dev) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tf = torch.zeros(2048 * 16, dtype=torch.int64, device=dev)
tb = torch.zeros(2048 * 16, dtype=torch.int64, device=dev)
for it in range(6):
    if it == 5:
        ext._lib.ttb_trace_set(tf.data_ptr(), tb.data_ptr())
    flush.fill_(it)
    emb(idx, off).backward(g)
torch.cuda.synchronize()
ext._lib.ttb_trace_set(None, None)
for name, t in (("plan", tf[1024 * 16:]), ("forward", tf[:1024 * 16]), ("backward", tb)):
    a = t.cpu().numpy().reshape(-1, 16).astype(np.int64)
    live = a[:, 0] > 0
    a = a[live]
    t0 = a[:, 0].min()
    print(f"== {name}: {len(a)} CTAs; CTA start spread {np.percentile(a[:, 0] - t0, [0, 50, 100])} ns; "
          f"kernel span {(a.max() - t0) / 1e3:.1f} us")
    for s in range(16):
        col = a[:, s]
        ok = col > 0
        if ok.any():
            rel = (col[ok] - a[ok, 0]) / 1e3
            print(f"   slot {s:2d}: n={int(ok.sum()):4d}  since CTA start  min {rel.min():6.2f}  med {np.median(rel):6.2f}  max {rel.max():6.2f} us")
