"""CPU tests of the constructor-side helpers: pure host code, no kernels."""
import json
import os
import time

import numpy as np
import torch

from fbtt_embedding_b200.tt_embeddings_ops import approx_uniform_cores, suggested_tt_shapes, tt_matrix_to_full

G = os.path.join(os.path.dirname(__file__), "golden")


def test_suggested_tt_shapes_equals_the_reference_on_its_own_outputs():
    """tests/golden/suggested_shapes.json was written by the reference's suggested_tt_shapes
    (tests/golden/make_shapes_golden.py): Criteo-Terabyte cardinalities, S1, config 5, random n, d in {2,3,4}."""
    cases = json.load(open(os.path.join(G, "suggested_shapes.json")))["cases"]
    assert len(cases) > 100
    t0 = time.time()
    for c in cases:
        assert suggested_tt_shapes(c["n"], c["d"], c["allow_round_up"]) == c["shape"], c
    assert time.time() - t0 < 30  # the reference needs ~28 s for these, ~6 s for 11M alone


def test_suggested_tt_shapes_invariants():
    rng = np.random.RandomState(3)
    for n in rng.randint(1, 10_000_000, size=40).tolist():
        for d in (2, 3, 4):
            s = suggested_tt_shapes(n, d)
            assert len(s) == d and all(v >= 1 for v in s)
            assert n <= int(np.prod(s)) < 10 * max(n, 10)  # rounded UP, by less than one decimal digit
            exact = suggested_tt_shapes(n, d, allow_round_up=False)
            assert int(np.prod(exact)) == n


def test_approx_uniform_cores_structure_and_spread():
    p, q, R = [20, 22, 25], [4, 4, 4], [1, 32, 32, 1]
    E = int(np.prod(p))
    g = torch.Generator().manual_seed(0)
    cores = approx_uniform_cores(E, p, q, R, g)
    assert [tuple(c.shape) for c in cores] == [(1, 20, 128), (1, 22, 4096), (1, 25, 128)]
    assert all(c.dtype == torch.float32 and c.is_contiguous() for c in cores)
    scale = E ** (-1.0 / 6.0)
    # tail [p2, r2, q2]: per (row digit, column digit) exactly one entry off the sigma background, at an odd rank
    tail = (cores[2][0].view(25, 32, 4) / scale).abs()
    big = tail > 0.05  # background is N(0, 0.01^2); saw-tooth values are multiples of 1/15 (j = 0 stays small)
    assert int(big.sum(dim=1).max()) <= 1
    assert bool((big.nonzero()[:, 1] % 2 == 1).all())
    # middle [p1, r1, q1, r2]: one damped even output-rank column per (row digit, column digit)
    mid = cores[1][0].view(22, 32, 4, 32) / scale
    col_mid = mid.median(dim=1).values  # over the input rank: ~1/sqrt(32) on live columns, ~0 on the damped one
    damped = col_mid.abs() < 0.08
    assert bool((damped.sum(dim=-1) == 1).all())
    assert bool((damped.nonzero()[:, 2] % 2 == 0).all())
    # the materialised table is spread over about [-1, 1] / sqrt(E) with a near-uniform second moment
    W = tt_matrix_to_full(p, q, R, cores, [1, 0, 2, 3]).numpy().ravel() * np.sqrt(E)
    assert -1.15 < W.min() and W.max() < 1.15
    assert abs(W.std() - 1 / np.sqrt(3)) < 0.06
    hist, _ = np.histogram(W, bins=10, range=(-1, 1))
    assert hist.min() > 0.03 * W.size  # no empty stretch
    # same generator state -> same draw
    again = approx_uniform_cores(E, p, q, R, torch.Generator().manual_seed(0))
    assert all(torch.equal(a, b) for a, b in zip(cores, again))
