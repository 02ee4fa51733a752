"""Undefined-name check for files that cannot be executed here (GPU tests, GPU-only scripts): every name a scope reads
as a global must be bound at module level or be a builtin.  No third-party linter is installed in the image.
  python scripts/lint_names.py [files...]      (default: tests/, scripts/, the package, bench.py, __graft_entry__.py)"""
import builtins
import glob
import os
import symtable
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(path):
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    module_names = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    module_names |= set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    bad = []

    def walk(tab):
        for s in tab.get_symbols():
            if s.is_referenced() and s.is_global() and s.get_name() not in module_names:
                bad.append((tab.get_name(), tab.get_lineno(), s.get_name()))
        for ch in tab.get_children():
            walk(ch)

    for s in top.get_symbols():
        if s.is_referenced() and not (s.is_assigned() or s.is_imported() or s.is_namespace()) \
                and s.get_name() not in module_names:
            bad.append(("<module>", 0, s.get_name()))
    for ch in top.get_children():
        walk(ch)
    return bad


def main():
    files = sys.argv[1:]
    if not files:
        for pat in ("tests/*.py", "scripts/*.py", "fbtt_embedding_b200/*.py", "fbtt_embedding_b200/dropin/*.py", "oracle/*.py",
                    "bench.py", "__graft_entry__.py"):
            files += sorted(glob.glob(os.path.join(ROOT, pat)))
    n = 0
    for f in files:
        for scope, line, name in check(f):
            print(f"{os.path.relpath(f, ROOT)}:{line}: in {scope}: undefined name {name!r}")
            n += 1
    print(f"{len(files)} files, {n} undefined names")
    return 1 if n else 0


if __name__ == "__main__":
    sys.exit(main())
