"""ncu target for the K-variables-per-lane kernel: 10 000 batched sudokus (config 2), one solve_batch call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I

grids = I.sudoku_batch(10000, base=100)
m = cb.Model(I.sudoku("." * 81))
p = cb.GpuProblem(m)
roots = I.sudoku_roots(m.var_names, grids)
r, counts, failed = p.solve_batch(roots, order="smallest-domain")
assert counts.tolist() == [1] * len(grids)
print("sudoku", r, "launches", r.kernel_launches)
