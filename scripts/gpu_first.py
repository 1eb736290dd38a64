"""First contact with the GPU: counts, optima and node-transition parity (development script)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I


def run(name, text, **kw):
    m = cb.Model(text)
    p = cb.GpuProblem(m)
    t = time.time()
    r = p.solve(**kw)
    dt = time.time() - t
    print("%-14s %s wall=%.3fs launches=%d expand_ms=%.2f nodes/s=%.3g" % (
        name, r, dt, r.kernel_launches, r.expand_ms, r.nodes / max(dt, 1e-9)), flush=True)
    return r


def replay_parity(name, text, n_walks=200, seed=7):
    ref_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle/_ref/libcsolve_ref.so")
    if not os.path.exists(ref_path):
        print("no reference replay library; skipped")
        return
    ref = C.CDLL(ref_path)
    path = "/tmp/_gpu_first_%s.txt" % name
    open(path, "w").write(text)
    n = ref.ref_load(path.encode(), 0, 1)
    assert n > 0
    I32P = C.POINTER(C.c_int32)
    m = cb.Model(text)
    p = cb.GpuProblem(m)
    V = n
    ov = m.obj_var
    import random
    rng = random.Random(seed)
    root = np.zeros(2 * V, np.int32)
    ref.ref_get_domains(root.ctypes.data_as(I32P))
    assert np.array_equal(root, m.root_domains)
    doms, vars_, vals, bests, exp_out, exp_fail = [], [], [], [], [], []
    for w in range(n_walks):
        dom = root.copy()
        un = list(range(V)); rng.shuffle(un)
        best = 2**31 - 1 if m.objective == 2 else (-2**31 if m.objective == 3 else 0)
        if ov >= 0 and rng.random() < 0.7:
            lo, hi = int(root[2 * ov]), int(root[2 * ov + 1])
            best = rng.randint(lo, min(hi, lo + 5000))
        while un:
            x = un.pop()
            lo, hi = int(dom[2 * x]), int(dom[2 * x + 1])
            val = rng.randint(lo, hi) if hi - lo < 50 else rng.choice([lo, hi, lo + 1, hi - 1, rng.randint(lo, lo + 20)])
            o = np.zeros(2 * V, np.int32)
            f = ref.ref_replay(dom.ctypes.data_as(I32P), x, val, best, o.ctypes.data_as(I32P), None)
            if not f and ov >= 0 and o[2 * ov] > o[2 * ov + 1]:
                f = 1   # documented deviation: empty <obj> is a failure on the device
            doms.append(dom.copy()); vars_.append(x); vals.append(val); bests.append(best)
            exp_out.append(o); exp_fail.append(f)
            if f:
                break
            dom = o
    out, failed = p.propagate_batch(np.array(doms), vars_, vals, bests)
    bad = 0
    for i in range(len(doms)):
        if bool(failed[i]) != bool(exp_fail[i]) or (not exp_fail[i] and not np.array_equal(out[i], exp_out[i])):
            bad += 1
            if bad < 4:
                print("MISMATCH", name, i, vars_[i], vals[i], bests[i], failed[i], exp_fail[i])
                print(" in ", doms[i].tolist()); print(" ref", exp_out[i].tolist()); print(" gpu", out[i].tolist())
    print("%-14s replay parity: %d nodes, %d failed, %d mismatches" % (name, len(doms), int(sum(exp_fail)), bad), flush=True)
    return bad


if __name__ == "__main__":
    bad = 0
    for nm, txt in [("queens8", I.queens(8)), ("sudoku", I.sudoku(I.SUDOKU_EXAMPLE)), ("schedule", I.schedule()),
                    ("wcet", I.wcet()), ("sat50", I.random_3sat(50))]:
        bad += replay_parity(nm, txt) or 0
    exp = {4: 2, 6: 4, 8: 92, 10: 724, 12: 14200}
    for n, e in exp.items():
        r = run("queens%d" % n, I.queens(n))
        print("   expected", e, "OK" if r.solutions == e else "WRONG")
    r = run("queens8-dom", I.queens(8), order="smallest-domain"); print("   expected 92", r.solutions == 92)
    r = run("sudoku", I.sudoku(I.SUDOKU_EXAMPLE)); print("   expected 1", r.solutions == 1)
    r = run("sudoku-dom", I.sudoku(I.SUDOKU_EXAMPLE), order="smallest-domain"); print("   expected 1", r.solutions == 1)
    r = run("schedule", I.schedule()); print("   expected best 11", r.best)
    r = run("sat20", I.random_3sat(20, seed=1, objective="ALL")); print("   expected 9 models", r.solutions)
    r = run("sat50", I.random_3sat(50, seed=1)); print("   expected UNSAT", r.has_solution)
    r = run("sat100", I.random_3sat(100, seed=1)); print("   expected SAT", r.has_solution)
    r = run("wcet", I.wcet(), time_limit_ms=120000); print("   expected best 1560", r.best)
    r = run("queens13", I.queens(13)); print("   expected 73712", r.solutions)
    r = run("queens14", I.queens(14)); print("   expected 365596", r.solutions)
    print("replay mismatches:", bad)
