"""Run by tests/test_gpu_zz_backjump.py in a process of its own (a kernel fault or a search that does not end must not
take the test session with it): csolve_solve_options.backjump on the device against the chronological search and the
oracle. Prints one JSON line per case and a final {"done": true}."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import csolve_b200 as cb
import util
from csolve_b200 import instances as I

LIMIT_MS = 10000


def check_model(m, cnf, assignment):
    val = dict(zip(m.var_names, assignment))
    return set(val.values()) <= {0, 1} and all(any((val["x%d" % abs(l)] == 1) == (l > 0) for l in cl) for cl in cnf)


def min_true_vars(n, cnf):
    """MIN x1 + ... + xn over the clauses of a CNF: 0/1 nogoods under an objective bound that is not 0/1"""
    text = I.cnf_to_csolve(n, cnf, "MIN " + " + ".join("x%d" % i for i in range(1, n + 1)))
    return text


def main():
    out = []
    # 1. decision instances: same status as the chronological search, models valid, the search does jump
    for n, ratio, seed in ((40, 4.26, 1), (40, 4.26, 2), (60, 4.26, 3), (60, 4.6, 4), (80, 4.26, 5), (80, 4.0, 6), (100, 4.26, 7)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(I.cnf_to_csolve(n, cnf))
        p = cb.GpuProblem(m)
        plain = p.solve(time_limit_ms=LIMIT_MS)
        for kw in ({"split_target": 1}, {}, {"prefer_failing": True}, {"slice_ms": 1, "split_target": 1}):
            r = p.solve(create_conflicts=True, backjump=True, time_limit_ms=LIMIT_MS, **kw)
            ok = r.timed_out == 0 and plain.timed_out == 0 and r.has_solution == plain.has_solution
            if ok and r.has_solution:
                ok = check_model(m, cnf, r.assignments[0])
            out.append({"case": "sat n=%d seed=%d %s" % (n, seed, kw), "ok": bool(ok), "sat": int(r.has_solution),
                        "nodes": int(r.nodes), "nodes_plain": int(plain.nodes), "conflicts": int(r.conflicts),
                        "backjumps": int(r.backjumps), "ms": r.kernel_ms + r.expand_ms})
            print(json.dumps(out[-1]), flush=True)
    # 2. optimisation: the optimum is the oracle's, with a witness that satisfies the clauses
    for n, ratio, seed in ((16, 3.0, 11), (20, 3.2, 12), (24, 3.5, 13)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(min_true_vars(n, cnf))
        o, _ = util.Oracle(m).solve_tree(0)
        p = cb.GpuProblem(m)
        for kw in ({"split_target": 1}, {}):
            r = p.solve(create_conflicts=True, backjump=True, max_solutions=16, time_limit_ms=LIMIT_MS, **kw)
            ok = r.timed_out == 0 and r.has_solution == o.has_solution and (not o.has_solution or r.best == o.best)
            if ok and r.has_solution:
                w = dict(zip(m.var_names, r.assignments[-1]))
                xs = {k: v for k, v in w.items() if k.startswith("x")}
                ok = sum(xs.values()) == r.best and all(any((xs["x%d" % abs(l)] == 1) == (l > 0) for l in cl) for cl in cnf)
            out.append({"case": "min n=%d seed=%d %s" % (n, seed, kw), "ok": bool(ok), "best": int(r.best), "best_oracle": int(o.best),
                        "conflicts": int(r.conflicts), "backjumps": int(r.backjumps)})
            print(json.dumps(out[-1]), flush=True)
    # 3. ALL models never jump: the counters stay the tree's
    cnf = I.random_3sat_cnf(30, 3.6, 21)
    m = cb.Model(I.cnf_to_csolve(30, cnf, "ALL"))
    p = cb.GpuProblem(m)
    a = p.solve(split_target=1)
    b = p.solve(create_conflicts=True, backjump=True, split_target=1)
    o, _ = util.Oracle(m).solve_tree(0)
    ok = (a.solutions, a.nodes, a.cuts) == (o.solutions, o.calls, o.cuts) and b.solutions == a.solutions and b.backjumps == 0
    out.append({"case": "all n=30", "ok": bool(ok), "solutions": int(b.solutions), "backjumps": int(b.backjumps)})
    print(json.dumps(out[-1]), flush=True)
    # 4. the chronological learning search is untouched by the option being off
    r = cb.GpuProblem(cb.Model(I.random_3sat(60, seed=3))).solve(create_conflicts=True, split_target=1)
    out.append({"case": "learning without backjump", "ok": r.backjumps == 0, "backjumps": int(r.backjumps)})
    print(json.dumps(out[-1]), flush=True)
    print(json.dumps({"done": True, "cases": len(out), "failed": [c["case"] for c in out if not c["ok"]],
                      "backjumps": sum(c.get("backjumps", 0) for c in out)}), flush=True)


if __name__ == "__main__":
    main()
