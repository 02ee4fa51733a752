"""The reference's OWN test-suite and benchmark, unmodified, through the drop-in seam (SURVEY 2.1 #16, 8b;
INTEGRATION.md option A / B).  The reference's Python files are staged by oracle/build_ref.sh into the git-ignored
oracle/_ref/py/ (they travel to the GPU box with the snapshot; /root/reference does not exist there).

  option A  reference tt_embeddings_ops.py + OUR `tt_embeddings` (the seam of tt_embeddings_ops.py:14)
  option B  OUR tt_embeddings_ops + OUR tt_embeddings (fbtt_embedding_b200/dropin on PYTHONPATH)
  option R  (benchmark only) reference tt_embeddings_ops.py + the REFERENCE's compiled extension (oracle/_ref, sm_100a
            rebuild): the reference's own number on the same box, recorded beside ours in
            gpurun_out/reference_benchmark_seam.txt (copied to profiles/)

Both run tt_embeddings_test.py:55-525 (six hypothesis property tests, rtol 1.3e-6 / atol 1e-5) on the exact fp32
path (TTB_PATH=generic: those are fp32 FFMA tolerances), and tt_embeddings_benchmark.py:124-215 once."""
import os
import re
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = os.path.join(ROOT, "oracle", "_ref", "py")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "tt_embeddings.cpython-312-x86_64-linux-gnu.so")
RECORD = os.path.join(ROOT, "gpurun_out", "reference_benchmark_seam.txt")
DROPIN = os.path.join(ROOT, "fbtt_embedding_b200", "dropin")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF_PY, "tt_embeddings_test.py")),
                               reason="reference Python files not staged (oracle/build_ref.sh needs /root/reference)")


def seam_dir(tmp_path, option):
    """A directory for PYTHONPATH that resolves `tt_embeddings` (and, for option B, `tt_embeddings_ops`) to ours and
    everything else of the reference (tests, benchmark, option A: its tt_embeddings_ops.py) to the staged files."""
    d = tmp_path / f"seam_{option}"
    d.mkdir()
    if option == "R":
        shutil.copy(REF_SO, d / os.path.basename(REF_SO))
    else:
        shutil.copy(os.path.join(DROPIN, "tt_embeddings.py"), d / "tt_embeddings.py")
    src_ops = os.path.join(DROPIN if option == "B" else REF_PY, "tt_embeddings_ops.py")
    shutil.copy(src_ops, d / "tt_embeddings_ops.py")
    for f in ("tt_embeddings_test.py", "tt_embeddings_benchmark.py"):
        shutil.copy(os.path.join(REF_PY, f), d / f)
    # the dropin shims locate the package relative to their own file: point them at the repo instead
    for f in (("tt_embeddings.py",) if option != "R" else ()) + (("tt_embeddings_ops.py",) if option == "B" else ()):
        txt = (d / f).read_text().replace(
            "_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))", f"_root = {ROOT!r}")
        (d / f).write_text(txt)
    return str(d)


def run(cmd, cwd, path):
    env = dict(os.environ, PYTHONPATH=cwd, TTB_PATH=path, HYPOTHESIS_STORAGE_DIRECTORY=os.path.join(cwd, ".hyp"))
    return subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=840)


@needs_ref
@pytest.mark.parametrize("option", ["A", "B"])
def test_reference_test_suite_passes_unmodified(tmp_path, option):
    d = seam_dir(tmp_path, option)
    r = run([sys.executable, "-m", "pytest", "tt_embeddings_test.py", "-q", "-x", "-p", "no:cacheprovider",
             "-o", "addopts="], d, "generic")
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) == 6, tail  # tt_embeddings_test.py: 6 property tests


@needs_ref
# (("A", "generic") ran too: 0.032 us/nnz, profiles/r2/reference_benchmark_through_seam.txt; dropped to keep the suite short)
@pytest.mark.parametrize("option,path", [("R", "reference"), ("A", "auto"), ("B", "auto")])
def test_reference_benchmark_runs_unmodified(tmp_path, option, path):
    """tt_embeddings_benchmark.py at its defaults = the README shape with use_cache=True never populated."""
    if option == "R" and not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref extension not built")
    d = seam_dir(tmp_path, option)
    r = run([sys.executable, "tt_embeddings_benchmark.py"], d, path)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    m = re.search(r"TTEmbeddingBag FWD-BWD time/nnz:\s*([0-9.]+) usecs", out)
    assert m, out[-3000:]
    us = float(m.group(1))
    what = {"R": "reference ops + reference CUDA extension (sm_100a rebuild)", "A": "reference ops + our tt_embeddings",
            "B": "our ops + our tt_embeddings"}[option]
    line = f"tt_embeddings_benchmark.py defaults, {what}, TTB_PATH={path}: {us} us/nnz (fwd+bwd)"
    print(line)
    try:
        os.makedirs(os.path.dirname(RECORD), exist_ok=True)
        with open(RECORD, "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
    assert 0 < us < 0.416  # the README's own number (README.md:21) is the ceiling for any B200 run
