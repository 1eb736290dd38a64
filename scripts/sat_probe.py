import sys,os
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I
for seed in (2,3,4,5,6,7,8,9,10,11):
    m=cb.Model(I.random_3sat(200, seed=seed))
    p=cb.GpuProblem(m)
    p.solve(prefer_failing=True)
    rs=[p.solve(prefer_failing=True) for _ in range(3)]
    print("seed %d sat=%d ms=%s nodes=%s"%(seed, rs[0].has_solution, ["%.1f"%(r.kernel_ms+r.expand_ms) for r in rs], [r.nodes for r in rs]), flush=True)
