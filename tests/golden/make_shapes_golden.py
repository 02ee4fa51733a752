"""Generates tests/golden/suggested_shapes.json with the REFERENCE's own suggested_tt_shapes
(tt_embeddings_ops.py:359-418), imported from /root/reference in the build container (it cannot travel to the
GPU box, the JSON does).  Run:  python tests/golden/make_shapes_golden.py
Cases: the 26 Criteo-Terabyte cardinalities + S1/config-5 sizes, random n for d in {2,3,4}, and
allow_round_up=False on embedding dims."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference_ops  # noqa: E402

CRITEO = [39884406, 39043, 17289, 7420, 20263, 3, 7120, 1543, 63, 38532951, 2953546, 403346, 10, 2208, 11938, 155,
          4, 976, 14, 39979771, 25641295, 39664984, 585935, 12972, 108, 36]


def main():
    ref = import_reference_ops()
    rng = np.random.RandomState(7)
    cases = []
    for n in CRITEO + [11_000_000, 50_000_000]:
        cases.append((int(n), 3, True))
    for d in (2, 3, 4):
        for n in rng.randint(1, 3_000_000, size=12).tolist() + [1, 2, 7, 64, 97, 1024, 99991]:
            cases.append((int(n), d, True))
    for n in (16, 32, 64, 128, 192, 256, 100, 81):
        for d in (2, 3, 4):
            cases.append((n, d, False))
    out = []
    for n, d, up in cases:
        out.append({"n": n, "d": d, "allow_round_up": up, "shape": [int(v) for v in ref.suggested_tt_shapes(n, d, up)]})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "suggested_shapes.json")
    json.dump({"source": "reference tt_embeddings_ops.suggested_tt_shapes", "cases": out}, open(path, "w"), indent=0)
    print(len(out), "cases ->", path)


if __name__ == "__main__":
    main()
