"""general-kernel timings: 3-SAT n=200 seeds 1, 5, 9 (UNSAT: whole tree), wcet, schedule"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I
for seed in (1, 5, 9):
    p = cb.GpuProblem(cb.Model(I.random_3sat(200, seed=seed)))
    p.solve(prefer_failing=True)
    rs = [p.solve(prefer_failing=True) for _ in range(3)]
    print("sat seed %d ms=%s nodes=%d" % (seed, ["%.1f" % (r.kernel_ms + r.expand_ms) for r in rs], rs[0].nodes), flush=True)
p = cb.GpuProblem(cb.Model(I.wcet()))
p.solve()
rs = [p.solve() for _ in range(4)]
print("wcet ms=%s best=%d" % (["%.1f" % (r.kernel_ms + r.expand_ms) for r in rs], rs[0].best), flush=True)
os.environ["CSOLVE_NO_LOV"] = "1"
p = cb.GpuProblem(cb.Model(I.queens(12)))
p.solve()
rs = [p.solve() for _ in range(3)]
print("queens12 general ms=%s sols=%d" % (["%.2f" % (r.kernel_ms + r.expand_ms) for r in rs], rs[0].solutions), flush=True)
