import sys, time
sys.path.insert(0,'/root/repo')
import csolve_b200 as cb
from csolve_b200 import instances as I
for seed in (1,2,3):
    m=cb.Model(I.random_3sat(200,seed=seed)); p=cb.GpuProblem(m)
    for pf in (True,):
        for order in ("none","smallest-domain"):
            t=time.time(); r=p.solve(order=order, prefer_failing=pf, time_limit_ms=60000); dt=time.time()-t
            print("seed",seed,"order",order,"pf",pf,"sat",r.has_solution,"timeout",r.timed_out,"nodes",r.nodes,"%.2fs"%dt, flush=True)
m=cb.Model(I.wcet()); p=cb.GpuProblem(m)
for pf in (False,True):
    t=time.time(); r=p.solve(prefer_failing=pf); print("wcet pf",pf,r.best,r.nodes,"%.3fs"%(time.time()-t))
