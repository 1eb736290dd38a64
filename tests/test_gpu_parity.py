"""Parity tests proper (run on the B200 with -m gpu): the CUDA path, called through the C ABI,
against the oracle, the reference-generated golden fixtures and size-independent properties."""
import glob
import json
import os

import numpy as np
import pytest

import csolve_b200 as cb
import util
from csolve_b200 import instances as I
from make_instances import instance_table, random_table

pytestmark = pytest.mark.gpu

INST = instance_table()
REPLAY = sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(util.GOLDEN, "replay_*.npz")) if "random" not in p)
COUNTS = json.load(open(os.path.join(util.GOLDEN, "ref_counts.json")))


def _load(name):
    z = np.load(os.path.join(util.GOLDEN, "replay_%s.npz" % name))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", REPLAY)
def test_node_transitions_match_reference_fixtures(name):
    """bit-exact post-fixpoint domains and fail flags on nodes replayed through the reference"""
    g = _load(name)
    m = cb.Model(INST[name])
    p = cb.GpuProblem(m)
    out, failed = p.propagate_batch(g["dom_in"], g["var"], g["val"], g["best"])
    assert np.array_equal(failed.astype(bool), g["failed"].astype(bool))
    ok = ~g["failed"].astype(bool)
    assert np.array_equal(out[ok], g["dom_out"][ok])


def test_node_transitions_random_instances():
    z = np.load(os.path.join(util.GOLDEN, "replay_random.npz"))
    g = {k: z[k] for k in z.files}
    rnd = random_table()
    n = 0
    for name in sorted({k.split("/")[0] for k in g}):
        sub = {k.split("/")[1]: g[k] for k in g if k.startswith(name + "/")}
        p = cb.GpuProblem(cb.Model(rnd[name]))
        out, failed = p.propagate_batch(sub["dom_in"], sub["var"], sub["val"], sub["best"])
        assert np.array_equal(failed.astype(bool), sub["failed"].astype(bool)), name
        ok = ~sub["failed"].astype(bool)
        assert np.array_equal(out[ok], sub["dom_out"][ok]), name
        n += len(failed)
    assert n > 500


def test_node_transitions_against_oracle_on_seeded_walks():
    """fresh seeded walks (not in the fixtures), oracle as checker, batch of ~20k nodes"""
    import random
    for name in ("queens10", "sudoku", "wcet", "sat100"):
        m = cb.Model(INST[name])
        orc = util.Oracle(m)
        p = cb.GpuProblem(m)
        rng = random.Random(99)
        V = m.n_vars
        root = m.root_domains
        doms, vs, xs, bs, eo, ef = [], [], [], [], [], []
        while len(doms) < 5000:
            dom = root.copy()
            un = list(range(V)); rng.shuffle(un)
            best = 2**31 - 1 if m.objective == 2 else (-2**31 if m.objective == 3 else 0)
            if m.obj_var >= 0 and rng.random() < 0.7:
                best = rng.randint(int(root[2 * m.obj_var]), int(root[2 * m.obj_var + 1]))
            while un:
                x = un.pop()
                v = rng.randint(int(dom[2 * x]), int(dom[2 * x + 1]))
                o, f = orc.node(dom, x, v, best)
                if not f and m.obj_var >= 0 and o[2 * m.obj_var] > o[2 * m.obj_var + 1]:
                    f = 1
                doms.append(dom.copy()); xs.append(x); vs.append(v); bs.append(best); eo.append(o); ef.append(f)
                if f:
                    break
                dom = o
        out, failed = p.propagate_batch(np.array(doms), xs, vs, bs)
        ef = np.array(ef, bool)
        assert np.array_equal(failed.astype(bool), ef), name
        assert np.array_equal(out[~ef], np.array(eo)[~ef]), name


ORDERS = [cb.ORDER_NONE, cb.ORDER_SMALLEST_DOMAIN, cb.ORDER_LARGEST_DOMAIN, cb.ORDER_SMALLEST_VALUE, cb.ORDER_LARGEST_VALUE]


@pytest.mark.parametrize("name", ["queens4", "queens6", "queens8", "queens10", "sudoku", "sat20all", "sat50all", "sat50"])
@pytest.mark.parametrize("order", ORDERS)
def test_all_solutions_counters_equal_oracle(name, order):
    """ALL mode: solutions, nodes and cuts are traversal-independent -> identical to the oracle's tree"""
    text = INST[name] if name != "sat50" else I.random_3sat(50, seed=1, objective="ALL")
    m = cb.Model(text)
    r = cb.GpuProblem(m).solve(order=order)
    o, _ = util.Oracle(m).solve_tree(order)
    assert (r.solutions, r.nodes, r.cuts) == (o.solutions, o.calls, o.cuts)
    if name in COUNTS and m.objective == cb.OBJ_ALL:
        assert r.solutions == COUNTS[name]["nocf"]["solutions"]     # the reference CLI's count


@pytest.mark.parametrize("split", [1, 64, 4096, 0])
def test_counters_do_not_depend_on_the_split(split):
    m = cb.Model(I.queens(9))
    o, _ = util.Oracle(m).solve_tree(0)
    r = cb.GpuProblem(m).solve(split_target=split, slice_ms=1)
    assert (r.solutions, r.nodes, r.cuts) == (o.solutions, o.calls, o.cuts) == (352, o.calls, o.cuts)


@pytest.mark.parametrize("general", [False, True])
@pytest.mark.parametrize("n,split,slice_ms", [(11, 1, 1), (12, 1, 1), (12, 40, 1), (13, 1, 2)])
def test_ticket_queue_hands_out_the_whole_tree(n, split, slice_ms, general, monkeypatch):
    """split_target=1: the expansion stops at the first level with a frame or more, so almost every warp gets its
    work through the donation ring (tickets, served slots, tickets given back at the many slice ends); the counters
    must not notice"""
    if general:
        monkeypatch.setenv("CSOLVE_NO_LOV", "1")
    m = cb.Model(I.queens(n))
    p = cb.GpuProblem(m)
    base = p.solve()
    for _ in range(3):
        r = p.solve(split_target=split, slice_ms=slice_ms)
        assert (r.solutions, r.nodes, r.cuts) == (base.solutions, base.nodes, base.cuts)
    assert base.solutions == {11: 2680, 12: 14200, 13: 73712}[n]


def test_stored_solutions_are_valid_and_distinct():
    m = cb.Model(I.queens(8))
    r = cb.GpuProblem(m).solve(max_solutions=200)
    assert r.solutions == 92 and len(r.assignments) == 92
    orc = util.Oracle(m)
    seen = set()
    for a in r.assignments:
        dom = np.repeat(np.array(a, np.int32), 2)
        assert orc.leaf_true(dom)
        seen.add(tuple(a))
    assert len(seen) == 92


@pytest.mark.parametrize("name", ["queens8any", "sudoku_any", "sat100"])
def test_any_returns_a_valid_assignment(name):
    m = cb.Model(INST[name])
    r = cb.GpuProblem(m).solve()
    assert r.has_solution == 1 and r.solutions == 1 and len(r.assignments) == 1
    assert util.Oracle(m).leaf_true(np.repeat(np.array(r.assignments[0], np.int32), 2))
    if name == "sudoku_any":   # unique solution: must be the one the reference prints
        printed = dict(kv.split(" = ") for kv in COUNTS[name]["nocf"]["last_solution"].split(", "))
        assert [int(printed[n]) for n in m.var_names] == r.assignments[0]


def test_unsat_reports_no_solution():
    r = cb.GpuProblem(cb.Model(INST["sat50"])).solve()
    assert r.has_solution == 0 and r.solutions == 0 and COUNTS["sat50"]["nocf"]["no_solution"]


@pytest.mark.parametrize("order", ORDERS)
def test_schedule_optimum(order):
    m = cb.Model(INST["schedule"])
    r = cb.GpuProblem(m).solve(order=order)
    assert r.best == COUNTS["schedule"]["nocf"]["best"] == 11 and r.has_solution
    w = dict(zip(m.var_names, r.assignments[-1]))
    assert w["end"] == 11 and w["<obj>"] == 11
    assert util.Oracle(m).leaf_true(np.repeat(np.array(r.assignments[-1], np.int32), 2))


@pytest.mark.parametrize("order", [cb.ORDER_NONE, cb.ORDER_SMALLEST_DOMAIN])
def test_wcet_optimum(order):
    m = cb.Model(INST["wcet"])
    r = cb.GpuProblem(m).solve(order=order)
    assert r.best == COUNTS["wcet"]["known_optimum"] == 1560
    w = dict(zip(m.var_names, r.assignments[-1]))
    assert w["<obj>"] == 1560
    assert util.Oracle(m).leaf_true(np.repeat(np.array(r.assignments[-1], np.int32), 2))


def test_random_optimisation_instances_match_oracle():
    """MIN/MAX on random instances: optimum (or infeasibility) equals the oracle's reference-mode run"""
    rnd = random_table()
    n = 0
    for name, text in sorted(rnd.items()):
        if not (text.startswith("MIN") or text.startswith("MAX")):
            continue
        try:
            m = cb.Model(text)
        except cb.CsolveError:
            continue
        o, _ = util.Oracle(m).solve_reference(max_calls=200000)
        if o.hit_limit:
            continue
        r = cb.GpuProblem(m).solve()
        assert bool(r.has_solution) == bool(o.has_solution), name
        if o.has_solution:
            assert r.best == o.best, name
        n += 1
    assert n >= 10


def test_random_all_instances_match_oracle():
    rnd = random_table()
    n = 0
    for name, text in sorted(rnd.items()):
        if not text.startswith("ALL"):
            continue
        try:
            m = cb.Model(text)
        except cb.CsolveError:
            continue
        o, _ = util.Oracle(m).solve_tree(0, max_calls=300000)
        if o.hit_limit:
            continue
        r = cb.GpuProblem(m).solve()
        assert (r.solutions, r.nodes, r.cuts) == (o.solutions, o.calls, o.cuts), name
        n += 1
    assert n >= 15


@pytest.mark.parametrize("parts", [2, 3, 8])
@pytest.mark.parametrize("split", [200, 1000000])
def test_partitions_are_disjoint_and_complete(parts, split):
    """multi-GPU partitioning, emulated on one GPU: the parts sum to the whole tree exactly.
    split=200: the frontier is dealt to the ranks by path hash; split=1000000: the breadth-first
    expansion exhausts the tree on every rank and only rank 0 may report it."""
    m = cb.Model(I.queens(10))
    p = cb.GpuProblem(m)
    whole = p.solve(split_target=split)
    tot = [0, 0, 0]
    per = []
    for r in range(parts):
        x = p.solve(part_rank=r, part_count=parts, split_target=split)
        tot[0] += x.solutions; tot[1] += x.nodes; tot[2] += x.cuts
        per.append(x.solutions)
    assert tuple(tot) == (whole.solutions, whole.nodes, whole.cuts) == (724, whole.nodes, whole.cuts)
    if split == 200:
        assert max(per) < 724 and min(per) > 0


def test_edge_cases():
    # one variable, domain of one value
    r = cb.GpuProblem(cb.Model("ALL; x = 3;")).solve(max_solutions=4)
    assert (r.solutions, r.nodes, r.assignments) == (1, 1, [[3]])
    # fails at the first level
    r = cb.GpuProblem(cb.Model("ALL; 0 <= x; x <= 2; 0 <= y; y <= 2; x != y; x + y = 4 | x + y = 0;")).solve()
    assert (r.solutions, r.has_solution, r.nodes, r.cuts) == (0, 0, 3, 3)
    # everything fixed at root: empty watch lists (the reference's fuzz/inputs/sat.txt shape)
    t = "ANY; (!x1|!x2|!x3) & (!x2|!x3|!x4) & (!x2|!x2|x3) & (x2|x2|x2); 0<=x1;x1<=1; 0<=x2;x2<=1; 0<=x3;x3<=1; 0<=x4;x4<=1;"
    r = cb.GpuProblem(cb.Model(t)).solve()
    assert r.assignments == [[0, 1, 1, 0]] and r.nodes == 4
    # a huge domain is bisected, never enumerated breadth-first
    r = cb.GpuProblem(cb.Model("MIN x; 5 <= x; x <= 2000000000; y = x + 1;")).solve()
    assert r.best == 5
    # negative domains, multiplication, both objective directions
    t = "-4 <= a; a <= 4; -4 <= b; b <= 4; a * b = 6; a < b;"
    o, _ = util.Oracle(cb.Model("ALL;" + t)).solve_tree(0)
    r = cb.GpuProblem(cb.Model("ALL;" + t)).solve()
    assert r.solutions == o.solutions == 2
    assert cb.GpuProblem(cb.Model("MAX a - b;" + t)).solve().best == -1
    assert cb.GpuProblem(cb.Model("MIN a + b;" + t)).solve().best == -5


def test_time_limit_reports_timeout():
    # 17-queens needs several hundred ms on one B200; a 20 ms budget must cut it short
    r = cb.GpuProblem(cb.Model(I.queens(17))).solve(time_limit_ms=20, slice_ms=5)
    assert r.timed_out == 1 and 0 < r.solutions < 95815104


def test_batched_sudoku_instances():
    """config 2 in miniature: generated unique-solution puzzles, one search root each"""
    for g in I.sudoku_batch(12, seed=20261018):
        m = cb.Model(I.sudoku(g))
        r = cb.GpuProblem(m).solve(order=cb.ORDER_SMALLEST_DOMAIN, max_solutions=2)
        assert r.solutions == 1
        sol = dict(zip(m.var_names, r.assignments[0]))
        for rr in range(9):
            for cc in range(9):
                if g[rr * 9 + cc] != ".":
                    assert sol[I._cell(rr, cc)] == int(g[rr * 9 + cc])


def test_batched_roots_share_one_network():
    """config 2: many sudokus over ONE resident network, root phase and search batched on the device"""
    grids = I.sudoku_batch(200, seed=20261018)
    grids.append("11" + "." * 79)                       # infeasible at root: two 1s in the first row
    grids.append(I.SUDOKU_EXAMPLE)
    m = cb.Model(I.sudoku("." * 81))
    p = cb.GpuProblem(m)
    roots = I.sudoku_roots(m.var_names, grids)
    r, counts, failed = p.solve_batch(roots, order=cb.ORDER_SMALLEST_DOMAIN, max_solutions=len(grids) + 8)
    assert failed.tolist() == [0] * 200 + [1, 0]
    assert counts.tolist() == [1] * 200 + [0, 1]
    assert r.solutions == 201
    names = m.var_names
    seen = set()
    for rid, vals in r.assignments:
        sol = dict(zip(names, vals))
        g = grids[rid]
        for rr in range(9):
            for cc in range(9):
                if g[rr * 9 + cc] != ".":
                    assert sol[I._cell(rr, cc)] == int(g[rr * 9 + cc])
        assert util.Oracle(m).leaf_true(np.repeat(np.array(vals, np.int32), 2))
        seen.add(rid)
    assert len(seen) == 201
    # the example's unique solution equals what the reference prints for examples/sudoku.txt
    printed = dict(kv.split(" = ") for kv in COUNTS["sudoku"]["nocf"]["last_solution"].split(", "))
    ex = [vals for rid, vals in r.assignments if rid == 201][0]
    assert [int(printed[n]) for n in names] == ex
    # per-root counters equal independent per-instance searches (oracle on the per-instance model)
    for g in grids[:5]:
        o, _ = util.Oracle(cb.Model(I.sudoku(g))).solve_tree(cb.ORDER_SMALLEST_DOMAIN)
        assert o.solutions == 1


def test_queens_full_size_counts():
    """BASELINE config 3 sizes; counts from OEIS A000170 (== the reference CLI's)"""
    for n, cnt in ((12, 14200), (13, 73712), (14, 365596)):
        r = cb.GpuProblem(cb.Model(I.queens(n))).solve()
        assert r.solutions == cnt
        # mirror symmetry X -> n+1-X: the tree differs, the count cannot
    r = cb.GpuProblem(cb.Model(I.queens(14))).solve(order=cb.ORDER_SMALLEST_DOMAIN)
    assert r.solutions == 365596


def test_exchange_callback_injects_incumbent_and_stops_any():
    """the per-slice exchange hook (multi-GPU incumbent all-reduce), emulated in one process"""
    m = cb.Model(INST["wcet"])
    p = cb.GpuProblem(m)
    base = p.solve(slice_ms=2)
    calls = []

    def other_rank_has_1559(best, found, local_done):
        calls.append((best, local_done))
        return max(best, 1559), found, local_done      # MAX model: another rank already holds 1559

    p.set_exchange(other_rank_has_1559)
    r = p.solve(slice_ms=2)
    p.set_exchange(None)
    assert r.best == 1560 and len(calls) >= 1 and calls[-1][1] == 1
    # with 1559 known from the first slice on, only 1560 itself can still be accepted as an improvement
    assert 1 <= r.solutions < base.solutions
    # ANY: another rank reports a solution -> this rank stops without one
    m2 = cb.Model(I.random_3sat(200, seed=1))
    p2 = cb.GpuProblem(m2)
    p2.set_exchange(lambda best, found, done: (best, 1, 1))
    r2 = p2.solve(slice_ms=1)
    p2.set_exchange(None)
    assert r2.has_solution == 0 and r2.nodes < 21000000


@pytest.mark.parametrize("name,order,general", [("queens15", cb.ORDER_NONE, False), ("queens13", cb.ORDER_NONE, True),
                                                ("queens9", cb.ORDER_SMALLEST_DOMAIN, True), ("sat50", cb.ORDER_NONE, True)])
def test_rebalance_export_import_keeps_the_counts(name, order, general, monkeypatch):
    """the frontier-rebalancing hook (csolve_gpu_export_frames / import_frames), emulated in one process: after every
    slice frames are split off the busy warps and shipped -- back to the same rank, or held for one slice as if they
    came from another rank. Every node must still be searched exactly once: the counters equal those of the plain
    search (which test_all_solutions_counters_equal_oracle ties to the oracle's tree)."""
    text = I.random_3sat(50, seed=1, objective="ALL") if name == "sat50" else I.queens(int(name[6:]))
    if general:
        monkeypatch.setenv("CSOLVE_NO_LOV", "1")       # general kernel: slow enough for several 1 ms slices
    m = cb.Model(text)
    p = cb.GpuProblem(m)
    base = p.solve(order=order)
    moved = []
    held = []

    def exchange(best, found, local_done):
        return best, found, 1 if (local_done and not held) else 0      # not done while frames are in transit

    def rebalance(prob, n_idle, n_busy, fw):
        got = 0
        if held:                                    # frames "received from another rank"
            got += prob.import_frames(held.pop())
        if n_busy > 0:
            fr = prob.export_frames(48)
            assert fr.shape[1] == fw
            moved.append(fr.shape[0])
            if fr.shape[0]:
                if len(moved) % 2:
                    held.append(fr.copy())
                else:
                    got += prob.import_frames(fr)
        return got

    p.set_exchange(exchange)
    p.set_rebalance(rebalance)
    r = p.solve(order=order, slice_ms=1, split_target=64)
    p.set_rebalance(None)
    p.set_exchange(None)
    assert (r.solutions, r.nodes, r.cuts) == (base.solutions, base.nodes, base.cuts)
    assert not held
    if name == "queens15":
        assert r.solutions == 2279184
    if name in ("queens15", "queens13"):
        assert sum(moved) > 0                       # the search was long enough to be rebalanced at least once


def test_conflict_learning_keeps_results_and_learns_sound_nogoods():
    """-c true on the device (src/conflict.c): identical counts / status, and every learned nogood is implied
    by the model (adding its literals as constraints leaves no solution)."""
    import re
    for name, nsol in (("sat20all", 9), ("sat50all", 23)):
        m = cb.Model(INST[name])
        p = cb.GpuProblem(m)
        r = p.solve(create_conflicts=True, split_target=1)   # no breadth-first phase: nogoods are learned depth-first
        assert r.solutions == nsol                       # the reference over-counts here with -c true
        assert r.conflicts > 0
        ngs = p.nogoods()
        assert len(ngs) == r.conflicts
        names = m.var_names
        checked = 0
        for ng in ngs[:: max(1, len(ngs) // 40)]:
            assert all(val in (0, 1) for _, val in ng) and len({v for v, _ in ng}) == len(ng)
            text = INST[name] + "".join("%s = %d;\n" % (names[v], val) for v, val in ng)
            try:
                o, _ = util.Oracle(cb.Model(text)).solve_tree(0)
                assert o.solutions == 0, (name, ng)
            except cb.CsolveError as e:
                assert e.code == -3                      # already infeasible at root
            checked += 1
        assert checked >= 5
    # status on the decision versions, with and without the failure-driven order
    for name, sat in (("sat50", 0), ("sat100", 1)):
        for pf in (False, True):
            r = cb.GpuProblem(cb.Model(INST[name])).solve(create_conflicts=True, prefer_failing=pf, split_target=1)
            assert r.has_solution == sat
    r = cb.GpuProblem(cb.Model(I.random_3sat(200, seed=1))).solve(create_conflicts=True, prefer_failing=True)
    assert r.has_solution == 0 and r.conflicts > 0
    # models on the NOT(EQ) kernels never produce a nogood (neither does the reference: CONFL 0 in BASELINE.md)
    r = cb.GpuProblem(cb.Model(I.queens(8))).solve(create_conflicts=True)
    assert (r.solutions, r.conflicts) == (92, 0)
    # general kernel, non-binary values: every analysis is abandoned, results unchanged
    r = cb.GpuProblem(cb.Model(INST["wcet"])).solve(create_conflicts=True, split_target=1)
    assert r.best == 1560 and r.conflicts == 0 and r.conflicts_abandoned > 0
