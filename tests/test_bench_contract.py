"""CPU test of bench.py's reference arm: one JSON line with the contract's keys (runs the oracle's CPU port
of the reference's full_weight() path on a tiny row sample)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "1", "--cpu-fraction", "200"], cwd=ROOT, timeout=600).decode()
    lines = [ln for ln in out.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "nnz/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_reference_arm_for_the_table_sharded_config_prints_one_contract_line():
    """--gpus N > 1: the arm's workload is BASELINE configs[3]; its CPU path runs on the tables it can materialise."""
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                                   "--steps", "1", "--warmup", "1", "--cpu-budget-s", "1"], cwd=ROOT, timeout=600).decode()
    lines = [ln for ln in out.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong" and d["value"] > 0
    assert "configs[3]" in d["config"]["workload"] and d["config"]["nnz_per_step"] == 26 * 4096
    assert d["cpu_baseline"]["kind"] == "port" and "tables" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_roofline_traffic_comes_from_the_committed_ncu_reduction():
    """roofline.traffic = mean DRAM read + write bytes per launch of the dominant kernel in profiles/r2/ (one
    `ncu --set full` capture of the bench command, reduced on the GPU box); a few MB for the README shape."""
    sys.path.insert(0, ROOT)
    import bench

    t = bench.ncu_traffic_bytes("x_bwd_kernel")
    assert t is not None and 1e6 < t < 2e7
    assert bench.ncu_traffic_bytes("no_such_kernel") is None
    assert bench.ncu_traffic_bytes("x_bwd_kernel", summary="missing_file.txt") is None
