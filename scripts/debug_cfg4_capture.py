import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_config4 as c4
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
class A: pass
r = c4.run(A(), 0, 0, 1, dev, steps=3, warmup=2, flush_buf=flush, check=True, graphs=True)
print({k: v for k, v in r["exchange"]["nccl"].items() if k != "kernel_ms"})
