"""cProfile of the eager module step (host overhead) at the BASELINE shape."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fbtt_embedding_b200 import OptimType, TTEmbeddingBag
torch.manual_seed(0)
E, D, B, POOL = 11_000_000, 64, 512, 20
emb = TTEmbeddingBag(E, D, [32, 32], [200, 220, 250], [4, 4, 4], optimizer=OptimType.SGD, learning_rate=0.1,
                     sparse=True, use_cache=False, weight_dist="uniform")
off = torch.arange(0, B * POOL + 1, POOL, device="cuda")
g = torch.rand(B, D, device="cuda") * 0.1
reqs = [torch.randint(0, E, (B * POOL,), device="cuda") for _ in range(10)]
def run(n):
    for i in range(n):
        emb(reqs[i % 10], off).backward(g)
run(20); torch.cuda.synchronize()
t0 = time.perf_counter(); run(300); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host time per step {1e6*(t1-t0)/300:.1f} us ; incl. final sync {1e6*(t2-t0)/300:.1f} us")
pr = cProfile.Profile(); pr.enable(); run(300); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
