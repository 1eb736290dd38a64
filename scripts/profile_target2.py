"""ncu targets for the general kernel: `sudoku` = 4000 batched sudokus, `sat` = random 3-SAT n=200 seed 1 (UNSAT)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I

kind = sys.argv[1] if len(sys.argv) > 1 else "sudoku"
if kind == "sudoku":
    grids = I.sudoku_batch(4000, base=50)
    m = cb.Model(I.sudoku("." * 81))
    p = cb.GpuProblem(m)
    r, counts, failed = p.solve_batch(I.sudoku_roots(m.var_names, grids), order="smallest-domain")
    assert counts.tolist() == [1] * len(grids)
else:
    m = cb.Model(I.random_3sat(200, seed=1))
    p = cb.GpuProblem(m)
    r = p.solve(prefer_failing=True)
print(kind, r, "launches", r.kernel_launches, "expand_ms %.3f" % r.expand_ms, "Mnodes/s %.1f" % (r.nodes / (r.kernel_ms + r.expand_ms) / 1e3))
