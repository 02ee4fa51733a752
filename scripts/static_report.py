"""Static evidence for profiles/: per-kernel registers / spills (ptxas -v) and the SASS mnemonics that prove which
hardware path a kernel uses (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = TMA bulk copy, HMMA = mma.sync,
RED/ATOM = L2 reductions).  Runs on a CPU-only box: nvcc cross-compiles, cuobjdump reads the built library.
Usage: python scripts/static_report.py > profiles/rN_static_sass_ptxas.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fbtt_embedding_b200", "csrc")
KEYS = ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "HMMA", "RED", "ATOM", "FFMA", "LDG", "LDS", "STS", "SHFL", "BAR")


def demangle(name):
    try:
        d = subprocess.check_output(["c++filt", name]).decode().strip()
    except Exception:
        d = name
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    return d.split("(")[0].replace("void ", "")


def main():
    print("# ptxas -v (nvcc -O3 -gencode arch=compute_100a,code=sm_100a), own kernels only")
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        out = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xptxas", "-v",
                              "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", src, "-o", os.devnull],
                             capture_output=True, text=True).stderr
        cur = None
        for line in out.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = demangle(m.group(1))
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                spill = m.groups()
            m = re.search(r"Used (\d+) registers", line)
            if m and cur and not cur.startswith("cub::"):
                print(f"{cur[:64]:64s} regs={m.group(1):>3s} stack={spill[0]:>3s}B spill_st={spill[1]:>3s}B spill_ld={spill[2]:>3s}B")
                cur = None
    print("\n# SASS mnemonic counts per kernel (cuobjdump -sass libttb.so)")
    sass = subprocess.check_output(["cuobjdump", "-sass", os.path.join(ROOT, "fbtt_embedding_b200", "lib", "libttb.so")]).decode()
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = demangle(f.split("\n", 1)[0].strip())
        if name.startswith("cub::"):
            continue
        c = collections.Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, flags=re.M):
            for k in KEYS:
                if m.group(1).startswith(k):
                    c[k] += 1
        print(f"{name[:64]:64s} " + " ".join(f"{k}={c[k]}" for k in KEYS if c[k]))


if __name__ == "__main__":
    main()
