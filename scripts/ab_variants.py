"""A/B the opt-in kernel variants (environment switches read by libttb, csrc/ttb_common.cuh `tuning_flag`) on the
bench workload: one child process per variant runs `bench.py --no-cpu --no-refcuda` (same seeds, same timing
rules) after a parity check of that variant against the exact fp32 path, and the table of results is printed.
Round-2 tool: the variants were written without a GPU at hand and stay off by default until measured.

  python scripts/ab_variants.py [--steps 300] [variant ...]      variants: base vecflush pdl vecflush+pdl
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    "base": {},
    "vecflush": {"TTB_BWD_VEC_FLUSH": "1"},
    "pdl": {"TTB_PDL": "1"},
    "vecflush+pdl": {"TTB_BWD_VEC_FLUSH": "1", "TTB_PDL": "1"},
}

PARITY = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from fbtt_embedding_b200 import tt_embeddings as ext
P, Q, R = [200, 220, 250], [4, 4, 4], [1, 32, 32, 1]
E, D, B, pool = 11_000_000, 64, 512, 20
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
cores0 = [((torch.rand(1, P[i], [128, 4096, 128][i], generator=g) - 0.5) * 0.2).to(dev) for i in range(3)]
idx = torch.randint(0, E, (B * pool,), generator=g).to(dev)
off = torch.arange(0, B * pool + 1, pool, device=dev)
go = (torch.rand(1, B, D, generator=g) * 0.1).to(dev)
L = torch.tensor([P[1] * P[2], P[2], 1], device=dev)
e64, e32 = torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int32, device=dev)
res = {}
for path in (ext.PATH_GENERIC, ext.PATH_AUTO):
    ext.set_path(path)
    cs = [c.clone() for c in cores0]
    for _ in range(3):  # several steps: plan reuse, scratch re-zeroing, back-to-back launches
        col, row, tbl, n, _ = ext.preprocess_indices_sync(idx, off, 1, True, e64, e32)
        out = ext.tt_forward(1000, 1, B, D, P, Q, R, L, n, col, row, tbl, cs)
        ext.tt_sgd_backward(1000, D, 0.1, P, Q, R, L, n, col, row, tbl, go, cs)
    torch.cuda.synchronize()
    res[path] = (out.cpu().numpy(), [c.cpu().numpy() for c in cs])
def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
ef = rel(res[ext.PATH_AUTO][0], res[ext.PATH_GENERIC][0])
eb = max(rel(a, b) for a, b in zip(res[ext.PATH_AUTO][1], res[ext.PATH_GENERIC][1]))
print("PARITY", ef, eb)
assert ef < 1e-3 and eb < 1e-2, (ef, eb)
""" % ROOT


def run(name, steps):
    env = dict(os.environ, **VARIANTS[name])
    par = subprocess.run([sys.executable, "-c", PARITY], env=env, capture_output=True, text=True, timeout=900)
    parity = [ln for ln in par.stdout.splitlines() if ln.startswith("PARITY")]
    row = {"variant": name, "env": VARIANTS[name], "parity_ok": par.returncode == 0, "parity": parity[-1] if parity else par.stderr[-300:]}
    if par.returncode == 0:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--no-refcuda", "--steps", str(steps),
                              "--warmup", "10"], env=env, capture_output=True, text=True, timeout=1800)
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if out.returncode == 0 and lines:
            d = json.loads(lines[-1])
            row.update(value=d["value"], ms_per_step=d["ms_per_step"], mode=d.get("value_mode"), eager_ms=d.get("eager_ms_per_step"),
                       e2e_ms=d["e2e"]["ms_per_step"],
                       kernel_us={k: round(v["mean_ms"] * 1e3, 2) for k, v in (d.get("kernel_ms") or {}).items() if v.get("mean_ms")})
        else:
            row["bench_error"] = (out.stderr or out.stdout)[-400:]
    print(json.dumps(row), flush=True)
    return row


def main():
    args = sys.argv[1:]
    steps = 300
    if "--steps" in args:
        i = args.index("--steps")
        steps = int(args[i + 1])
        del args[i:i + 2]
    names = args or list(VARIANTS)
    rows = [run(n, steps) for n in names]
    base = next((r for r in rows if r["variant"] == "base" and "ms_per_step" in r), None)
    for r in rows:
        if "ms_per_step" in r:
            rel = f"{base['ms_per_step'] / r['ms_per_step']:.3f}x" if base else "-"
            print(f"{r['variant']:14s} {r['ms_per_step'] * 1e3:8.2f} us/step  {rel:>8s}  kernels {r.get('kernel_us')}  parity {r['parity']}")
        else:
            print(f"{r['variant']:14s} FAILED: {r.get('bench_error') or r.get('parity')}")


if __name__ == "__main__":
    main()
