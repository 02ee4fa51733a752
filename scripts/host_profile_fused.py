"""Host cost of one FusedTTEmbeddingBag training step (26 config-4 tables), measured WITHOUT a GPU: the three device
entry points are stubbed out, tensors live on the CPU, and the output memset is skipped -- what remains is exactly the
Python / ctypes / autograd work a step costs above libttb.  Compare: one per-table module step costs ~190 us of host
time on the B200 box (profiles/r1_host_profile.txt), i.e. ~4.9 ms for 26 tables.
  python scripts/host_profile_fused.py"""
import contextlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fbtt_embedding_b200 import OptimType
from fbtt_embedding_b200 import tt_embeddings as ext
from fbtt_embedding_b200.fused import FusedTTEmbeddingBag, pack_table_major
from bench_config4 import CARD, PSHAPE


class _Stub:
    def __init__(self, real):
        self.real = real

    def __getattr__(self, name):
        return getattr(self.real, name)

    def ttb_preprocess_rowidx(self, *a):
        return 0

    def ttb_tt_forward_het(self, *a):
        return 0

    def ttb_tt_backward_het(self, *a):
        return 0


def main():
    ext._lib = _Stub(ext._lib)
    ext._i64c = lambda t, what: t
    ext._f32c = lambda t, what: t
    ext._cores_inplace = lambda cores, what="tt_cores": [c.data for c in cores]
    ext._DeviceGuard = lambda t: contextlib.nullcontext()
    ext._stream = lambda: 0
    torch.cuda.is_current_stream_capturing = lambda: False
    zeros = torch.zeros
    B, D = 4096, 128
    # small ranks: the cores only have to exist (CPU memory), their size does not enter the host cost
    m = FusedTTEmbeddingBag(CARD, D, [8, 8], [PSHAPE[e] for e in CARD], [4, 4, 8], optimizer=OptimType.SGD,
                            weight_dist="uniform", device="cpu")
    ci, co = pack_table_major([torch.randint(0, e, (B,)) for e in CARD], [torch.arange(0, B + 1)] * len(CARD))
    g = zeros(len(CARD), B, D)
    torch.zeros = lambda *a, **k: torch.empty(*a, **k)  # a 54 MB CPU memset is not what a CUDA step pays

    def step():
        m(ci, co).backward(g)

    for _ in range(50):
        step()
    t0 = time.perf_counter()
    n = 500
    for _ in range(n):
        step()
    print(f"host time per fused step over {len(CARD)} tables: {(time.perf_counter() - t0) / n * 1e6:.0f} us")


if __name__ == "__main__":
    main()
