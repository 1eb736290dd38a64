"""The K-variables-per-lane form on all-different style networks (every pair clause is x_i != x_j: sudoku, latin
squares, graph colouring): the kernel reads such a network through an adjacency bit matrix (DevModel::lov_adj) instead
of the 64-bit offset sets. Generated networks with 33..100 variables, random walks, lanes emulated on the host
(tests/harness) against the oracle; networks with an offset somewhere must keep the offset sets."""
import ctypes as C
import random

import numpy as np
import pytest

import csolve_b200 as cb
import util


def _network(seed, offsets=False):
    rng = random.Random(seed)
    n = rng.randint(33, 100)
    hi = rng.randint(4, 12)
    lines = ["ALL;"]
    for g in range(rng.randint(6, 30)):
        members = rng.sample(range(n), rng.randint(2, min(hi, 9)))
        args = ["X%d" % v for v in members]
        if offsets and g == 0:
            args[0] += "+1"
        lines.append("all_different(%s);" % ", ".join(args))
    for v in range(n):
        lines.append("1 <= X%d; X%d <= %d;" % (v, v, hi))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(40))
def test_adjacency_form_against_the_oracle(seed):
    m = cb.Model(_network(seed))
    hc = util.harness_lib()
    assert hc.hc_load(C.byref(m.flat), 1) == 0, hc.hc_error()
    assert hc.hc_lov_adj_only() == 1
    orc = util.Oracle(m)
    rng = random.Random(1000 + seed)
    V = m.n_vars
    checked = 0
    for _walk in range(6):
        dom = m.root_domains.copy()
        order = list(range(V))
        rng.shuffle(order)
        for v in order:
            lo, hi = int(dom[2 * v]), int(dom[2 * v + 1])
            val = rng.randint(lo, hi)
            eo, ef = orc.node(dom, v, val)
            out = np.empty_like(dom)
            lf = hc.hc_node_lov(util.p32(np.ascontiguousarray(dom, np.int32)), v, val, util.p32(out))
            assert lf >= 0
            assert bool(lf) == bool(ef), (seed, v, val)
            checked += 1
            if ef:
                break
            assert np.array_equal(out, eo), (seed, v, val)
            dom = eo
    assert checked >= 6


@pytest.mark.parametrize("seed", range(5))
def test_networks_with_an_offset_keep_the_offset_sets(seed):
    m = cb.Model(_network(seed, offsets=True))
    hc = util.harness_lib()
    assert hc.hc_load(C.byref(m.flat), 1) == 0, hc.hc_error()
    assert hc.hc_lov_adj_only() == 0
