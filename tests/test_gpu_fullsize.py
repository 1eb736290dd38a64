"""GPU parity at BASELINE.json's full sizes, where the oracle would need minutes to hours: the
results are checked through properties that do not need it -- OEIS A000170 counts (the compiled reference's counts
in BASELINE.md are the same numbers), partitions that add up, puzzles whose solution is unique by construction,
satisfying assignments evaluated directly on the CNF, the z3-confirmed SAT/UNSAT status of the seeded instances."""
import numpy as np
import pytest

import csolve_b200 as cb
from csolve_b200 import instances as I

pytestmark = pytest.mark.gpu

A000170 = {14: 365596, 15: 2279184, 16: 14772512}


@pytest.mark.parametrize("n", [15, 16])
def test_config3_queens_all_solutions(n):
    """config 3 (N-queens 14..16, all solutions); 14 is covered with the counters in test_gpu_parity.py"""
    p = cb.GpuProblem(cb.Model(I.queens(n)))
    r = p.solve()
    assert r.solutions == A000170[n]
    assert r.nodes - r.cuts > r.solutions          # every solution hangs off successful inner nodes
    # the tree is a property of the model: a different frontier / slice schedule must not change a single counter
    r2 = p.solve(split_target=100000, slice_ms=3)
    assert (r2.solutions, r2.nodes, r2.cuts) == (r.solutions, r.nodes, r.cuts)
    # value symmetry X -> n+1-X maps solutions to solutions: largest-domain order gives another tree, the same count
    if n == 15:
        r3 = p.solve(order=cb.ORDER_LARGEST_DOMAIN)
        assert r3.solutions == A000170[n] and r3.nodes != r.nodes


def test_config3_partition_over_8_ranks_adds_up():
    """what each of 8 GPUs does (one rank's share emulated on this device): shares are disjoint and complete"""
    p = cb.GpuProblem(cb.Model(I.queens(15)))
    whole = p.solve()
    tot = np.zeros(3, np.int64)
    for rk in range(8):
        r = p.solve(part_rank=rk, part_count=8)
        assert 0 < r.solutions < whole.solutions
        tot += np.array([r.solutions, r.nodes, r.cuts], np.int64)
    assert tot.tolist() == [whole.solutions, whole.nodes, whole.cuts]


def test_config2_ten_thousand_sudokus_one_network():
    """config 2: 10 000 generated unique-solution puzzles over one resident network, one search root each"""
    grids = I.sudoku_batch(10000, base=100)
    assert len(set(grids)) > 9900                   # distinct instances (symmetry transforms of 100 dug puzzles)
    m = cb.Model(I.sudoku("." * 81))
    p = cb.GpuProblem(m)
    roots = I.sudoku_roots(m.var_names, grids)
    r, counts, failed = p.solve_batch(roots, order=cb.ORDER_SMALLEST_DOMAIN, max_solutions=10000)
    assert not failed.any() and counts.tolist() == [1] * 10000 and r.solutions == 10000
    where = {I._cell(rr, cc): (rr, cc) for rr in range(9) for cc in range(9)}
    pos = [where[nm] for nm in m.var_names]
    seen = set()
    for rid, vals in r.assignments:
        seen.add(rid)
        g = grids[rid]
        grid = np.zeros((9, 9), np.int32)
        for k, (rr, cc) in enumerate(pos):
            grid[rr, cc] = vals[k]
        for i in range(9):                          # a valid sudoku ...
            assert sorted(grid[i, :]) == list(range(1, 10)) and sorted(grid[:, i]) == list(range(1, 10))
            b = grid[3 * (i // 3):3 * (i // 3) + 3, 3 * (i % 3):3 * (i % 3) + 3]
            assert sorted(b.reshape(-1)) == list(range(1, 10))
        for k, ch in enumerate(g):                  # ... that extends the clues
            assert ch == "." or grid[k // 9, k % 9] == int(ch)
    assert len(seen) == 10000


@pytest.mark.parametrize("seed,sat", [(1, False), (2, True), (3, True)])
def test_config5_random_3sat_n200(seed, sat):
    """config 5: uniform random 3-SAT n=200, m=852; status as z3 reports it, model checked on the CNF"""
    cnf = I.random_3sat_cnf(200, seed=seed)
    m = cb.Model(I.cnf_to_csolve(200, cnf))
    r = cb.GpuProblem(m).solve(prefer_failing=True, time_limit_ms=120000)
    assert r.timed_out == 0 and r.has_solution == (1 if sat else 0)
    if sat:
        val = dict(zip(m.var_names, r.assignments[0]))
        assert set(val.values()) <= {0, 1}
        for cl in cnf:
            assert any((val["x%d" % abs(l)] == 1) == (l > 0) for l in cl), cl


# status of seeds 1..11 (z3 for 1-3; the compiled reference CLI for all of them, profiles/r2_sat_seeds.md)
SAT200_STATUS = {1: False, 2: True, 3: True, 4: True, 5: False, 6: True, 7: True, 8: True, 9: False, 10: True, 11: False}


@pytest.mark.parametrize("seed", sorted(SAT200_STATUS))
def test_config5_with_luby_restarts(seed):
    """-r 100 (the reference's default, src/main.c:109-113): the search goes back to the root on the Luby schedule and
    expands it again in the order of the failure-driven priorities learned so far (src/csolve.c:76-83,264-276,380-385).
    Restarts change the tree, never the answer: SAT / UNSAT as without them, every model checked on the CNF."""
    cnf = I.random_3sat_cnf(200, seed=seed)
    m = cb.Model(I.cnf_to_csolve(200, cnf))
    p = cb.GpuProblem(m)
    restarts = 0
    for rf in (100, 5):
        r = p.solve(prefer_failing=True, restart_frequency=rf, time_limit_ms=120000)
        assert r.timed_out == 0 and bool(r.has_solution) == SAT200_STATUS[seed], (seed, rf)
        restarts += r.restarts
        if r.has_solution:
            val = dict(zip(m.var_names, r.assignments[0]))
            for cl in cnf:
                assert any((val["x%d" % abs(l)] == 1) == (l > 0) for l in cl), cl
    if not SAT200_STATUS[seed]:
        assert restarts > 0                      # an unsatisfiable instance outlasts the first thresholds


def test_restarts_leave_all_mode_and_later_searches_alone():
    """restarts are for ANY models only (is_restartable(), src/csolve.c:212-215): ALL-mode counts are the tree's; and
    the static order a restart rewrote on the device is the model's own again for the next search"""
    m = cb.Model(I.random_3sat(20, seed=1, objective="ALL"))
    p = cb.GpuProblem(m)
    a = p.solve(prefer_failing=True)
    b = p.solve(prefer_failing=True, restart_frequency=1)
    assert b.solutions == a.solutions > 0 and b.restarts == 0
    m = cb.Model(I.random_3sat(100, seed=4))
    p = cb.GpuProblem(m)
    before = p.solve()                                   # static order: a deterministic tree when unsatisfiable
    r = p.solve(prefer_failing=True, restart_frequency=1)
    after = p.solve()
    assert bool(r.has_solution) == bool(before.has_solution)
    if not before.has_solution:
        assert (after.nodes, after.cuts) == (before.nodes, before.cuts)


def test_config4_optima_with_and_without_learning():
    """config 4: schedule (MIN, optimum 11) and wcet (MAX, optimum 1560, confirmed with z3 in BASELINE.md)"""
    for text, best in ((I.schedule(), 11), (I.wcet(), 1560)):
        m = cb.Model(text)
        p = cb.GpuProblem(m)
        for kw in ({}, {"create_conflicts": True}, {"slice_ms": 1}, {"split_target": 1}):
            r = p.solve(**kw)
            assert r.has_solution == 1 and r.best == best
            assert r.assignments[-1][m.obj_var] == best
