"""CPU: the oracle against tests/golden/search_<name>.npz -- search nodes recorded by the B200 kernels at BASELINE
sizes (queens 14..16, the 10 000-sudoku batch, schedule, wcet, 3-SAT n=200) and replayed through the compiled
reference (tests/golden/make_search_samples.py). Pins the oracle on the states a real search visits; also the
oracle's tree partition used to produce tests/golden/tree_counts.json."""
import json
import os

import numpy as np
import pytest

import csolve_b200 as cb
import search_samples as S
import util
from csolve_b200 import instances as I

FIXTURES = [n for n in S.SAMPLED if os.path.exists(S.fixture_path(n))]


def test_the_fixtures_are_there():
    assert len(FIXTURES) >= 8, FIXTURES


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_equals_reference_on_search_sampled_nodes(name):
    fx = S.load_samples(S.fixture_path(name))
    m = cb.Model(S.SAMPLED[name]["text"]())
    # a bounded, evenly spread subset keeps the CPU suite short; the GPU suite checks every record
    idx = np.arange(len(fx["var"]))[:: max(1, len(fx["var"]) // 1500)]
    sub = {k: (v[idx] if getattr(v, "ndim", 0) >= 1 and len(v) == len(fx["var"]) else v) for k, v in fx.items()}
    n, nonfailed, bad = S.check_against_oracle(m, sub)
    assert not bad, bad[:3]
    assert n >= min(len(fx["var"]), 1000)
    if name != "schedule":
        assert int(((fx["flags"] & S.FAILED) == 0).sum()) >= 10000     # >= 10 k non-failed nodes per instance


@pytest.mark.parametrize("text,order", [(I.queens(9), 0), (I.queens(8), 1), (I.random_3sat(50, seed=1), 0),
                                        (I.random_3sat(20, seed=1, objective="ALL"), 0)])
def test_oracle_tree_partition_adds_up(text, order):
    m = cb.Model(text)
    o = util.Oracle(m)
    w, _ = o.solve_tree(order)
    for n_parts, split in ((5, 0), (7, 1), (48, 8)):
        tot = np.zeros(3, np.int64)
        for k in range(n_parts):
            r = o.solve_tree_part(order, k, n_parts, split)
            tot += [r.solutions, r.calls, r.cuts]
        assert tot.tolist() == [w.solutions, w.calls, w.cuts]


def test_tree_counts_small_entries_rederived():
    tree = json.load(open(os.path.join(util.GOLDEN, "tree_counts.json")))
    for key, order in (("queens11/none", 0), ("queens11/largest-domain", 2)):
        o, _ = util.Oracle(cb.Model(I.queens(11))).solve_tree(order)
        assert (o.solutions, o.calls, o.cuts) == (tree[key]["solutions"], tree[key]["nodes"], tree[key]["cuts"])
    # OEIS A000170
    for n, cnt in ((11, 2680), (12, 14200), (13, 73712), (14, 365596), (15, 2279184), (16, 14772512)):
        if "queens%d/none" % n in tree:
            assert tree["queens%d/none" % n]["solutions"] == cnt
    rows = tree["sudoku_batch200_seed20261018/smallest-domain"]["per_root"]
    g = I.sudoku_batch(3, seed=20261018)
    for k in range(3):
        o, _ = util.Oracle(cb.Model(I.sudoku(g[k]))).solve_tree(1)
        assert [o.solutions, o.calls, o.cuts] == rows[k]
