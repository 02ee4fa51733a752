"""Top stall SASS lines from `ncu --page source --csv` output (file arg)."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 6 and r[2].isdigit()]
tot = sum(int(r[2]) for r in rows)
print("total samples", tot, "instructions", len(rows))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(rows, key=lambda r: -int(r[2]))[:n]:
    print(f"{int(r[2]):6d} {100*int(r[2])/max(tot,1):5.1f}%  exec={r[5]:>7s}  {r[1].strip()[:100]}")
