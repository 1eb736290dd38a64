"""Small fixed workload for ncu: one all-solutions search (default 13-queens, slices of 2 ms)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I

n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
order = sys.argv[2] if len(sys.argv) > 2 else "none"
kind = sys.argv[3] if len(sys.argv) > 3 else "queens"
text = I.queens(n) if kind == "queens" else (I.sudoku(I.SUDOKU_EXAMPLE) if kind == "sudoku" else I.wcet())
p = cb.GpuProblem(cb.Model(text))
r = p.solve(order=order, slice_ms=int(os.environ.get("SLICE_MS", "0")), split_target=int(os.environ.get("SPLIT", "0")))
print(r, "launches", r.kernel_launches, "expand_ms %.3f" % r.expand_ms, "visits", r.clause_visits,
      "Mnodes/s %.1f" % (r.nodes / (r.kernel_ms + r.expand_ms) / 1e3))
