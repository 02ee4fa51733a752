"""Reduce an `ncu --set full` report (exported with --page raw --csv / --page source --csv) to what profiles/ keeps:
per kernel launch: duration, grid, registers, shared memory, achieved occupancy, DRAM bytes and GB/s, tensor-pipe %,
issue-active %, the top stall reasons; per kernel (source page): the 12 SASS instructions with most stall samples.
    python scripts/ncu_reduce.py report.raw.csv [report.source.csv]"""
import csv
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sectors_op_red.sum", "l2_red_sectors"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h, u = rows[0], rows[1]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki][:110])
        vals = {}
        for col, name in KEEP:
            if col in h:
                i = h.index(col)
                vals[name] = (r[i], u[i])
                print(f"   {name:22s} {r[i]} {u[i]}")
        try:
            dur = float(vals["duration"][0]) * SCALE.get(vals["duration"][1], 1.0)
            byt = sum(float(vals[k][0]) * SCALE.get(vals[k][1], 1.0) for k in ("dram_read", "dram_write"))
            print(f"   {'dram_GB/s':22s} {byt / dur / 1e9:.1f}   (dram bytes {byt / 1e6:.3f} MB)")
        except Exception:
            pass
        stalls = []
        for i, c in enumerate(h):
            if "issue_stalled" in c and c.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), c.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("   stalls/issue          " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]))
    if len(sys.argv) > 2:
        try:
            src = list(csv.reader(open(sys.argv[2])))
        except Exception:
            return
        kernel, hdr, data = None, None, []

        def flush():
            if not data:
                return
            tot = sum(d[0] for d in data) or 1
            print("== source hot spots:", (kernel or "")[:100])
            for s, txt in sorted(data, reverse=True)[:12]:
                print(f"   {100.0 * s / tot:5.1f}%  {txt[:100]}")

        for r in src:
            if r and r[0] == "Kernel Name":
                flush()
                kernel, hdr, data = r[1] if len(r) > 1 else "", None, []
            elif r and r[0] == "Address":
                hdr = r
            elif hdr and len(r) == len(hdr):
                try:
                    data.append((int(r[hdr.index("# Samples")] or 0), r[hdr.index("Source")].strip()))
                except (ValueError, IndexError):
                    pass
        flush()


if __name__ == "__main__":
    main()
