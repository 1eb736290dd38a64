import sys, time
sys.path.insert(0,'/root/repo')
import csolve_b200 as cb
from csolve_b200 import instances as I
for seed in (1,2,3):
    m=cb.Model(I.random_3sat(200,seed=seed)); p=cb.GpuProblem(m)
    for learn in (False, True):
        t=time.time(); r=p.solve(prefer_failing=True, create_conflicts=learn, time_limit_ms=60000); dt=time.time()-t
        print("seed",seed,"learn",learn,"sat",r.has_solution,"timeout",r.timed_out,"nodes",r.nodes,"confl",r.conflicts,"abandoned",r.conflicts_abandoned,"%.3fs"%dt, flush=True)
