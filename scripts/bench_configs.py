"""Secondary configs of BASELINE.json next to the reference CUDA kernels (same inputs, ops called directly).

  s1_nnz   README shape at larger batches (nnz = 10240 .. 262144, fused SGD, no cache): where the tile
           pipeline fills and the tensor-core path is no longer launch-latency bound
  cfg3     config 3 flow: zipf indices, B=2048, nnz=65536, Adagrad, LFU cache of 2^20 rows, hashtbl = E
           (fp32 cores: the reference has no bf16 path)
  cfg5     rank sweep (BASELINE configs[4]): E=50M D=128 B=1024, ranks 8 / 16 / 32 / 64 / 128 -- kernel family per rank,
           TFLOP/s at the benchmark convention 3F and the fraction of the measured tensor peak
Every line carries a `check`: ours against the reference CUDA kernels on the SAME inputs (forward <= 1e-3, dense
core gradients <= 1e-2, max-norm relative -- the north-star bounds) before anything is timed.
Writes one JSON object per line to stdout.  Not the driver's bench (that is bench.py)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fbtt_embedding_b200 import tt_embeddings as ext
from tests.helpers import load_reference_extension

P, Q, R = [200, 220, 250], [4, 4, 4], [1, 32, 32, 1]
E, D = 11_000_000, 64
dev = torch.device("cuda:0")
ref = load_reference_extension()
L = torch.tensor([P[1] * P[2], P[2], 1], device=dev, dtype=torch.int64)
e64 = torch.empty(0, dtype=torch.int64, device=dev)
e32 = torch.empty(0, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
try:
    PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {}


def cores():
    g = torch.Generator(device="cpu").manual_seed(0)
    return [(torch.rand(1, P[i], [128, 4096, 128][i], generator=g) - 0.5).mul_(0.2).to(dev) for i in range(3)]


def timeit(step, steps=30, warm=5):
    for i in range(warm):
        step(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(steps):
        flush.fill_(i & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(i)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    timeit.last = {"min": ts[0], "median": ts[len(ts) // 2], "max": ts[-1]}
    return ts[len(ts) // 2]  # median: robust against a stray slow step


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def check_vs_reference(p, q, Rr, Lt, B, Dd, idx, off, cs, go, bf16=False):
    """ours vs the reference CUDA kernels: forward and dense core gradients on the same inputs."""
    if ref is None:
        return {"unavailable": "oracle/_ref not built"}
    ours_cores = [c.to(torch.bfloat16) for c in cs] if bf16 else cs
    col, row, tbl, n, _ = ref.preprocess_indices_sync(idx, off, 1, True, e64, e32)
    o_ref = ref.tt_forward(1000, 1, B, Dd, p, q, Rr, Lt, n, col, row, tbl, cs)
    g_ref = ref.tt_dense_backward(1000, Dd, p, q, Rr, Lt, n, col, row, tbl, go, cs)
    o = ext.tt_forward(1000, 1, B, Dd, p, q, Rr, Lt, n, col, row, tbl, ours_cores)
    g = ext.tt_dense_backward(1000, Dd, p, q, Rr, Lt, n, col, row, tbl, go, ours_cores)
    res = {"fwd_max_rel": rel(o, o_ref), "dense_grad_max_rel": max(rel(a, b) for a, b in zip(g, g_ref))}
    res["ok"] = bool(res["fwd_max_rel"] < 1e-3 and res["dense_grad_max_rel"] < 1e-2)
    if not res["ok"]:
        raise SystemExit(f"parity check failed: {res}")
    return res


def s1_nnz():
    for B, pool in [(512, 20), (2048, 32), (4096, 64)]:
        nnz = B * pool
        reqs = [torch.randint(0, E, (nnz,), device=dev) for _ in range(4)]
        off = torch.arange(0, nnz + 1, pool, device=dev)
        go = (torch.rand(1, B, D, device=dev) * 0.1)
        res = {"config": "s1_nnz", "B": B, "nnz": nnz,
               "check": check_vs_reference(P, Q, R, L, B, D, reqs[0], off, cores(), go)}
        for name, mod in (("ours", ext), ("reference_cuda", ref)):
            if mod is None:
                continue
            cs = cores()

            def step(i, mod=mod, cs=cs):
                col, row, tbl, n, _ = mod.preprocess_indices_sync(reqs[i % 4], off, 1, True, e64, e32)
                mod.tt_forward(1000, 1, B, D, P, Q, R, L, n, col, row, tbl, cs)
                mod.tt_sgd_backward(1000, D, 0.1, P, Q, R, L, n, col, row, tbl, go, cs)

            ms = timeit(step)
            res[name] = {"ms_per_step": ms, "nnz_per_s": nnz / ms * 1e3, "spread_ms": dict(timeit.last)}
        if "reference_cuda" in res:
            res["speedup"] = res["reference_cuda"]["ms_per_step"] / res["ours"]["ms_per_step"]
        print(json.dumps(res), flush=True)


def cfg3():
    B, pool, C = 2048, 32, 1 << 20
    nnz = B * pool
    rng = np.random.RandomState(0)
    warm = [torch.as_tensor((rng.zipf(1.05, size=nnz) % E).astype(np.int64), device=dev) for _ in range(6)]
    reqs = [torch.as_tensor((rng.zipf(1.05, size=nnz) % E).astype(np.int64), device=dev) for _ in range(6)]
    off = torch.arange(0, nnz + 1, pool, device=dev)
    go = torch.rand(1, B, D, device=dev) * 0.1
    res = {"config": "cfg3", "B": B, "nnz": nnz, "cache_size": C, "zipf_a": 1.05,
           "check_fp32": check_vs_reference(P, Q, R, L, B, D, reqs[0], off, cores(), go),
           "check_bf16": check_vs_reference(P, Q, R, L, B, D, reqs[0], off,
                                            [c.to(torch.bfloat16).float() for c in cores()], go, bf16=True)}
    for name, mod in (("ours", ext), ("ours_async", ext), ("ours_bf16", ext), ("ours_bf16_async", ext),
                      ("reference_cuda", ref)):
        if mod is None:
            continue
        cs = cores()
        st = [torch.zeros_like(c) for c in cs]
        if "bf16" in name:  # BASELINE configs[2]: bf16 cores, fp32 accumulation / state / cache rows
            cs = [c.to(torch.bfloat16) for c in cs]
        hashtbl = torch.full((E,), -1, dtype=torch.int64, device=dev)
        freq = torch.zeros(E, dtype=torch.int64, device=dev)
        cstate = torch.full((E,), -1, dtype=torch.int32, device=dev)
        cw = torch.zeros(C, D, device=dev)
        cst = torch.zeros(C, device=dev)
        for r in warm:  # warm-up batches count frequencies; the timed batches are fresh draws
            mod.update_cache_state(r, hashtbl, freq)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        mod.cache_populate(E, P, Q, R, cs, L, hashtbl, freq, cstate, cw)
        b.record()
        torch.cuda.synchronize()
        populate_ms = a.elapsed_time(b)
        frac = {}

        def step(i, mod=mod):
            idx = reqs[i % 6]
            mod.update_cache_state(idx, hashtbl, freq)
            col, row, tbl, ntt, loc = mod.preprocess_indices_sync(idx, off, 1, False, hashtbl, cstate)
            frac["cached"] = 1.0 - ntt / nnz
            out = mod.tt_forward(1000, 1, B, D, P, Q, R, L, ntt, col, row, tbl, cs)
            mod.cache_forward(B, nnz - ntt, loc[ntt:], row[ntt:], cw, out)
            mod.tt_adagrad_backward(1000, D, 0.1, 1e-10, P, Q, R, L, ntt, col, row, tbl, go, st, cs)
            mod.cache_backward_rowwise_adagrad_approx(nnz - ntt, go, loc[ntt:], row[ntt:], 0.1, 1e-10, cst, cw)

        def step_async(i):
            # one front-end launch, batch kept in order, no D2H count / stream sync
            idx = reqs[i % 6]
            col, row, tbl, loc = ext.cache_frontend(idx, off, 1, hashtbl, freq, cstate)
            out = ext.tt_forward(1000, 1, B, D, P, Q, R, L, nnz, col, row, tbl, cs, cache_locations=loc)
            ext.cache_forward(B, nnz, loc, row, cw, out)
            ext.tt_adagrad_backward(1000, D, 0.1, 1e-10, P, Q, R, L, nnz, col, row, tbl, go, st, cs, cache_locations=loc)
            ext.cache_backward_rowwise_adagrad_approx(nnz, go, loc, row, 0.1, 1e-10, cst, cw)

        if name.endswith("async"):
            step = step_async
            frac["cached"] = None
        ms = timeit(step, steps=20, warm=3)
        res[name] = {"ms_per_step": ms, "nnz_per_s": nnz / ms * 1e3, "cache_populate_ms": populate_ms,
                     "cached_fraction": frac.get("cached"), "spread_ms": dict(timeit.last)}
        if mod is ext:
            ext.kernel_timing_begin()
            for i in range(5):
                step(i)
            kt = ext.kernel_timing_end()
            res[name]["kernel_us_per_step"] = {k: round(v["total_ms"] * 1000 / 5, 1) for k, v in kt.items()}
        del hashtbl, freq, cstate, cw
        torch.cuda.empty_cache()
    if "reference_cuda" in res:
        res["speedup"] = res["reference_cuda"]["ms_per_step"] / res["ours"]["ms_per_step"]
    print(json.dumps(res), flush=True)


def cfg5():
    """rank sweep (BASELINE config 5): E=50M -> p=[250,400,500], D=128, q=[4,4,8], B=1024, pooling 20."""
    p5, q5 = [250, 400, 500], [4, 4, 8]
    E5, D5, B, pool = 50_000_000, 128, 1024, 20
    nnz = B * pool
    L5 = torch.tensor([p5[1] * p5[2], p5[2], 1], device=dev, dtype=torch.int64)
    reqs = [torch.randint(0, E5, (nnz,), device=dev) for _ in range(4)]
    off = torch.arange(0, nnz + 1, pool, device=dev)
    go = torch.rand(1, B, D5, device=dev) * 0.1
    for r in (8, 16, 32, 64, 128):
        R5 = [1, r, r, 1]
        S = [4 * r, r * 4 * r, r * 8]
        F = 2 * (4 * r * 4 * r + 16 * r * 8)
        family = {8: "generic fp32 FFMA"}.get(r, "tcgen05 kind::f16, bf16 hi/lo split")
        if r == 16 and os.environ.get("TTB_LEGACY_R16", "0") == "1":
            family = "warp-level mma.sync tf32"
        g = torch.Generator(device="cpu").manual_seed(r)
        cs0 = [((torch.rand(1, p5[i], S[i], generator=g) - 0.5) * 0.2).to(dev) for i in range(3)]
        res = {"config": "cfg5", "rank": r, "nnz": nnz, "F_fwd_flop_per_nnz": F, "kernel_family": family,
               "check": check_vs_reference(p5, q5, R5, L5, B, D5, reqs[0], off, cs0, go)}
        for name, mod in (("ours", ext), ("reference_cuda", ref)):
            if mod is None:
                continue
            cs = [c.clone() for c in cs0]

            def step(i, mod=mod, cs=cs):
                col, row, tbl, n, _ = mod.preprocess_indices_sync(reqs[i % 4], off, 1, True, e64, e32)
                mod.tt_forward(1000, 1, B, D5, p5, q5, R5, L5, n, col, row, tbl, cs)
                mod.tt_sgd_backward(1000, D5, 0.1, p5, q5, R5, L5, n, col, row, tbl, go, cs)

            ms = timeit(step, steps=10, warm=3)
            res[name] = {"ms_per_step": ms, "nnz_per_s": nnz / ms * 1e3, "tflops_3F": 3 * F * nnz / ms / 1e9}
            if mod is ext:
                peak = PEAKS.get("bf16_tflops", 1661.8) / (2.0 if "tf32" in family else 1.0)  # bf16 MMAs / tf32
                res[name]["frac_of_tensor_peak_3F"] = res[name]["tflops_3F"] / peak if r >= 16 else None
                ext.kernel_timing_begin()
                for i in range(5):
                    step(i)
                kt = ext.kernel_timing_end()
                res[name]["kernel_us_per_step"] = {k: round(v["total_ms"] * 1000 / 5, 1) for k, v in kt.items() if v["count"]}
            del cs
            torch.cuda.empty_cache()
        if "reference_cuda" in res:
            res["speedup"] = res["reference_cuda"]["ms_per_step"] / res["ours"]["ms_per_step"]
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["s1_nnz", "cfg3"]
    if "all" in which:
        which = ["cfg5", "s1_nnz", "cfg3"]
    if "cfg5" in which:
        cfg5()
    if "s1_nnz" in which:
        s1_nnz()
    if "cfg3" in which:
        cfg3()
