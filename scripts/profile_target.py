"""The ncu targets, one script: one search of a fixed workload, warm-up excluded by ncu's own kernel filter.
    profile_target.py queens [N] [order]     N-queens ALL (k_search_lov<false,true>), default 16
    profile_target.py sudoku [n_roots]       batched sudokus (k_search_lovk<false,3>), default 10000
    profile_target.py wcet                   examples/wcet.txt MAX (k_search<false,false,LIN>)
    profile_target.py sat [seed]             random 3-SAT n=200 (k_search_sat), -f true -r 100
    profile_target.py sat-general [seed]     the same on the general kernel (CSOLVE_NO_SAT=1)
Environment: SLICE_MS, SPLIT (split_target)."""
import os
import sys

kind = sys.argv[1] if len(sys.argv) > 1 else "queens"
if kind == "sat-general":
    os.environ["CSOLVE_NO_SAT"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I

kw = dict(slice_ms=int(os.environ.get("SLICE_MS", "0")), split_target=int(os.environ.get("SPLIT", "0")))
if kind == "queens":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    r = cb.GpuProblem(cb.Model(I.queens(n))).solve(order=sys.argv[3] if len(sys.argv) > 3 else "none", **kw)
elif kind == "sudoku":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    grids = I.sudoku_batch(n, base=100)
    m = cb.Model(I.sudoku("." * 81))
    r, counts, failed = cb.GpuProblem(m).solve_batch(I.sudoku_roots(m.var_names, grids), order="smallest-domain")
    assert counts.tolist() == [1] * len(grids)
elif kind == "wcet":
    r = cb.GpuProblem(cb.Model(I.wcet())).solve(**kw)
else:
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    r = cb.GpuProblem(cb.Model(I.random_3sat(200, seed=seed))).solve(prefer_failing=True, restart_frequency=100, **kw)
print(kind, r, "launches", r.kernel_launches, "expand_ms %.3f" % r.expand_ms, "visits", r.clause_visits,
      "Mnodes/s %.1f" % (r.nodes / max(r.kernel_ms + r.expand_ms, 1e-9) / 1e3))
