import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, csolve_b200 as cb
from csolve_b200 import instances as I
for name, text in (("schedule", I.schedule()), ("wcet", I.wcet())):
    g = np.load('/root/repo/tests/golden/replay_%s.npz' % name)
    m = cb.Model(text)
    p = cb.GpuProblem(m)
    out, failed = p.propagate_batch(g["dom_in"], g["var"], g["val"], g["best"])
    bad = 0
    for i in range(len(g['var'])):
        if bool(failed[i]) != bool(g['failed'][i]) or (not failed[i] and not np.array_equal(out[i], g['dom_out'][i])):
            bad += 1
            if bad <= 3:
                dom = g['dom_in'][i]
                print(name, i, m.var_names[g['var'][i]], g['val'][i], g['best'][i], "ref failed", g['failed'][i], "gpu", failed[i])
                print(" in ", dict(zip(m.var_names, zip(dom[0::2].tolist(), dom[1::2].tolist()))))
                print(" ref", dict(zip(m.var_names, zip(g['dom_out'][i][0::2].tolist(), g['dom_out'][i][1::2].tolist()))))
                print(" gpu", dict(zip(m.var_names, zip(out[i][0::2].tolist(), out[i][1::2].tolist()))))
    print(name, "n", len(g['var']), "bad", bad)
