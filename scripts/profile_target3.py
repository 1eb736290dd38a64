"""ncu target for the general kernel: `wcet` or `sat` (random 3-SAT n=200 seed 1), one solve."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I

kind = sys.argv[1] if len(sys.argv) > 1 else "sat"
if kind == "wcet":
    r = cb.GpuProblem(cb.Model(I.wcet())).solve(slice_ms=int(os.environ.get("SLICE_MS", "0")))
else:
    r = cb.GpuProblem(cb.Model(I.random_3sat(200, seed=1))).solve(prefer_failing=True, slice_ms=int(os.environ.get("SLICE_MS", "0")))
print(kind, r, "launches", r.kernel_launches)
