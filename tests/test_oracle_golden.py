"""Oracle (and the product's host-run contractors) against fixtures produced by the COMPILED
reference (tests/golden/make_golden.py): node transitions, leaf verdicts, CLI counters."""
import glob
import json
import os

import numpy as np
import pytest

import csolve_b200 as cb
import util
from make_instances import instance_table, random_table

INST = instance_table()
REPLAY = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(util.GOLDEN, "replay_*.npz")) if "random" not in p)


def _check_nodes(g, node_fn, leaf_fn):
    g = {k: g[k] for k in (g.files if hasattr(g, "files") else g)}   # NpzFile decompresses on every access
    bad = 0
    for i in range(len(g["var"])):
        out, f = node_fn(g["dom_in"][i], int(g["var"][i]), int(g["val"][i]), int(g["best"][i]))
        if bool(f) != bool(g["failed"][i]) or (not f and not np.array_equal(out, g["dom_out"][i])):
            bad += 1
    for i in range(len(g["leaf"])):
        if bool(leaf_fn(g["leaf"][i])) != bool(g["leaf_true"][i]):
            bad += 1
    return bad


@pytest.mark.parametrize("name", REPLAY)
def test_oracle_node_transitions(name):
    g = np.load(os.path.join(util.GOLDEN, "replay_%s.npz" % name))
    m = cb.Model(INST[name])
    assert np.array_equal(m.root_domains, g["root"])
    o = util.Oracle(m)
    assert len(g["var"]) > 100
    assert _check_nodes(g, o.node, o.leaf_true) == 0


@pytest.mark.parametrize("name", REPLAY)
@pytest.mark.parametrize("specialise", [0, 1, 2])
def test_product_contractors_on_host(name, specialise):
    """contract.cuh (what each lane runs) with a sequential worklist: interpreter only (0), the compiled
    watch records the kernels use (1), specialised per-clause records (2)"""
    g = np.load(os.path.join(util.GOLDEN, "replay_%s.npz" % name))
    m = cb.Model(INST[name])
    hc = util.harness_lib()
    assert hc.hc_load(m.flat, specialise) == 0, hc.hc_error()

    def node(dom, var, val, best):
        dom = np.ascontiguousarray(dom, np.int32)
        out = np.empty_like(dom)
        f = hc.hc_node(util.p32(dom), var, val, best, util.p32(out))
        return out, f

    def leaf(dom):
        dom = np.ascontiguousarray(dom, np.int32)
        return hc.hc_leaf_true(util.p32(dom))

    assert _check_nodes(g, node, leaf) == 0


@pytest.mark.parametrize("name", [n for n in REPLAY if n.startswith("queens") or n.startswith("sudoku")])
def test_lane_owns_variable_form_on_host(name):
    """the register-resident forms (N-queens: one variable per lane; sudoku: three per lane, forbidden-value sets)
    with the warp emulated on the host"""
    g = np.load(os.path.join(util.GOLDEN, "replay_%s.npz" % name))
    m = cb.Model(INST[name])
    hc = util.harness_lib()
    assert hc.hc_load(m.flat, 1) == 0, hc.hc_error()
    # sudoku is an all-different style network (every pair clause has offset 0): the K-per-lane kernel reads it through
    # the adjacency bit matrix; the diagonals of N-queens carry offsets, so they keep the 64-bit offset sets
    assert hc.hc_lov_adj_only() == (1 if name.startswith("sudoku") else 0)

    def node(dom, var, val, best):
        dom = np.ascontiguousarray(dom, np.int32)
        out = np.empty_like(dom)
        f = hc.hc_node_lov(util.p32(dom), var, val, util.p32(out))
        assert f >= 0, "pure NOT(EQ) networks must be eligible"
        return out, f

    assert _check_nodes(g, node, lambda d: hc.hc_leaf_true(util.p32(np.ascontiguousarray(d, np.int32)))) == 0


def test_random_instances_node_transitions():
    z = np.load(os.path.join(util.GOLDEN, "replay_random.npz"))
    g = {k: z[k] for k in z.files}
    rnd = random_table()
    names = sorted({k.split("/")[0] for k in g})
    assert len(names) >= 50
    hc = util.harness_lib()
    total = 0
    for name in names:
        sub = {k.split("/")[1]: g[k] for k in g if k.startswith(name + "/")}
        m = cb.Model(rnd[name])
        o = util.Oracle(m)
        assert _check_nodes(sub, o.node, o.leaf_true) == 0, name
        assert hc.hc_load(m.flat, 1) == 0

        def node(dom, var, val, best):
            dom = np.ascontiguousarray(dom, np.int32)
            out = np.empty_like(dom)
            return out, hc.hc_node(util.p32(dom), var, val, best, util.p32(out))
        assert _check_nodes(sub, node, lambda d: hc.hc_leaf_true(util.p32(np.ascontiguousarray(d, np.int32)))) == 0, name
        total += len(sub["var"])
    assert total > 500


COUNTS = json.load(open(os.path.join(util.GOLDEN, "ref_counts.json")))


@pytest.mark.parametrize("name", sorted(k for k in COUNTS if "nocf" in COUNTS[k]))
def test_oracle_reference_mode_counters(name):
    """solve() restated: CALLS / CUTS / PROPS / SOLUTIONS / BEST equal to the reference CLI (-c false -r 0)"""
    exp = COUNTS[name]["nocf"]
    m = cb.Model(INST[name])
    r, sol = util.Oracle(m).solve_reference()
    assert (r.calls, r.cuts, r.props, r.solutions) == (exp["calls"], exp["cuts"], exp["props"], exp["solutions"])
    if m.obj_var >= 0:
        assert r.best == exp["best"]
    assert bool(r.has_solution) == (not exp["no_solution"])
    if exp["last_solution"] and m.objective != cb.OBJ_ALL:
        # ANY / MIN / MAX: the printed assignment is the one the oracle returns
        names = m.var_names
        printed = dict(kv.split(" = ") for kv in exp["last_solution"].split(", "))
        assert [int(printed[n]) for n in names] == sol.tolist()


def test_oracle_tree_mode_counts():
    """the traversal-independent tree (what the device explores) finds the same solutions"""
    for name, nsol in (("queens4", 2), ("queens6", 4), ("queens8", 92), ("sudoku", 1), ("sat20all", 9), ("sat50all", 23)):
        for order in range(5):
            r, _ = util.Oracle(cb.Model(INST[name])).solve_tree(order)
            assert r.solutions == nsol, (name, order)
    r, _ = util.Oracle(cb.Model(INST["sat50"])).solve_tree(0)
    assert r.solutions == 0
    r, _ = util.Oracle(cb.Model(INST["wcet"])).solve_tree(0, max_calls=3000000)
    assert (r.best, r.hit_limit) == (COUNTS["wcet"]["known_optimum"], 0)
