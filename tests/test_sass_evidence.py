"""Static check of the built library (no GPU needed: cuobjdump reads the sm_100a cubin): every kernel of the tcgen05
family really issues tcgen05.mma / tcgen05.ld (SASS UTCHMMA / LDTM) and no warp-level mma.sync (HMMA) for every rank
it claims (16 / 32 / 64 / 128, fp32 and bf16 cores); the forward variants that stage core-2 slices through shared
memory use the bulk-copy engine (UBLKCP).  The same mnemonics are listed per kernel in
profiles/r2/static_sass_ptxas.txt (scripts/static_report.py)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fbtt_embedding_b200", "lib", "libttb.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not on PATH")
    if not os.path.exists(LIB):
        pytest.skip("libttb.so not built")
    text = subprocess.check_output(["cuobjdump", "-sass", LIB]).decode()
    funcs = {}
    for f in re.split(r"\n\s*Function : ", text)[1:]:
        mangled, body = f.split("\n", 1)
        funcs[mangled.strip()] = body
    names = subprocess.check_output(["c++filt"], input="\n".join(funcs).encode()).decode().splitlines()
    return {n.replace("(anonymous namespace)::", ""): funcs[m] for m, n in zip(funcs, names)}


def count(body, mnemonic):
    return len(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?" + mnemonic + r"\b", body, flags=re.M))


def test_every_rank_of_the_tcgen05_family_issues_tcgen05(sass):
    seen = set()
    for name, body in sass.items():
        m = re.search(r"xk::x_(fwd|bwd)_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (float|__nv_bfloat16)>", name)
        if not m:
            continue
        kind, r, q2, core = m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)
        seen.add((kind, r, q2, core))
        assert count(body, "UTCHMMA") >= 1, f"{name}: no tcgen05.mma"
        assert count(body, r"LDTM[.\w]*") >= 1, f"{name}: no tcgen05.ld"
        assert count(body, r"HMMA[.\w]*") == 0, f"{name}: warp-level mma.sync in a tcgen05 kernel"
        if kind == "bwd":  # three GEMMs per tile (recompute, dA0, dB1^T), each with its split-precision terms
            assert count(body, "UTCHMMA") >= (9 if core == "float" else 5), name
    want = {(k, r, q2, c) for k in ("fwd", "bwd") for r in (16, 32, 64, 128) for q2 in (4, 8)
            for c in ("float", "__nv_bfloat16")}
    assert seen == want, sorted(want - seen)


def test_forward_stages_core2_slices_with_the_bulk_copy_engine_where_they_fit(sass):
    """XCfg::kC2Smem: ranks 16 / 32 (q2 = 4, 8) and rank 64 with q2 = 4, fp32 cores."""
    for name, body in sass.items():
        m = re.search(r"xk::x_fwd_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), float>", name)
        if not m:
            continue
        r, q2 = int(m.group(1)), int(m.group(2))
        fits = r <= 32 or (r == 64 and q2 == 4)
        assert (count(body, r"UBLKCP[.\w]*") >= 1) == fits, name


def test_warp_mma_family_is_the_only_user_of_mma_sync(sass):
    for name, body in sass.items():
        if "cub::" in name:
            continue
        if count(body, r"HMMA[.\w]*"):
            assert "bk::tt_" in name, f"{name}: unexpected mma.sync"
