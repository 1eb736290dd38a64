"""Config 5 over seeds 1..11 on one GPU: device time and nodes of three runs each, without restarts and with the
reference's default restart frequency (-r 100). usage: sat_probe.py [restart frequencies ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I

freqs = [int(x) for x in sys.argv[1:]] or [0, 100]
for seed in range(1, 12):
    m = cb.Model(I.random_3sat(200, seed=seed))
    p = cb.GpuProblem(m)
    for rf in freqs:
        p.solve(prefer_failing=True, restart_frequency=rf)
        rs = [p.solve(prefer_failing=True, restart_frequency=rf) for _ in range(3)]
        print("seed %d r=%d sat=%d ms=%s nodes=%s restarts=%s" % (seed, rf, rs[0].has_solution, ["%.1f" % (r.kernel_ms + r.expand_ms) for r in rs],
                                                                 [r.nodes for r in rs], [r.restarts for r in rs]), flush=True)
