import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k: d.get(k) for k in ["value", "ms_per_step", "eager_ms_per_step", "graph_ms_per_step", "back_to_back_ms_per_step", "gpu_launches_per_step", "value_mode"]})
print("e2e", d.get("e2e"))
print({k: (round(v["mean_ms"] * 1000, 2) if v["mean_ms"] else None) for k, v in d.get("kernel_ms", {}).items()}, "us")
print("ref_cuda ms", d.get("reference_cuda", {}).get("ms_per_step"), " roofline", d.get("roofline"))
print("clocks", d.get("clocks"))
