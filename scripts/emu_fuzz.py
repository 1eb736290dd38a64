"""Differential fuzz of the search kernel SOURCE on the CPU: csolve_b200/csrc/kernels.cu (k_search) runs under the SIMT
emulator of tests/harness/simt_emu.h and is compared with the oracle -- ALL: (solutions, nodes, cuts) of the tree;
ANY: status and a model that satisfies the CNF; MIN / MAX: the optimum. With --backjump the back-jumping build
(csrc/kernels_bj.cu) is used for the learning runs.   python scripts/emu_fuzz.py [--seeds N] [--seed0 S] [--backjump]"""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import csolve_b200 as cb
import gen_random
import util
from csolve_b200 import instances as I


def ranks_search(m, kw, rng):
    """the search as `world` ranks run it in ALL mode: everybody expands the root, the frames are dealt by path hash, the
    counters add up (the expansion is reported by rank 0)"""
    world = rng.choice([1, 1, 2, 3])
    kw = dict(kw, split_target=rng.choice([1, 1, 8, 64, 500]))
    if world > 1 and rng.random() < 0.4 and not kw.get("sink_headroom"):
        # ... with frames shipped from the busiest rank to ranks that ran dry at the slice boundaries
        r, moved = util.emu_search_exchange(m, world, order=kw["order"], n_blocks=kw["n_blocks"], split_target=max(kw["split_target"], 2),
                                            slice_clock=kw["slice_clock"] or 5000, general=kw["general"])
        return r, [], dict(kw, world=world, exchange=moved)
    tot = None
    for rank in range(world):
        r, sols = util.emu_search(m, part_rank=rank, part_count=world, **kw)
        if tot is None:
            tot = r
        else:
            tot.solutions += r.solutions; tot.nodes += r.nodes; tot.cuts += r.cuts
    return tot, sols, dict(kw, world=world)


def sink_check(m, o, kw, rng):
    """ALL models now and then through a bounded solution buffer that is drained between slices (the solution sink of the
    drop-in): every solution exactly once, each of them a leaf the oracle accepts"""
    import numpy as np
    # the head room capi.cu reserves (sink_headroom()): every warp may add a solution per node until its next poll
    kw = dict(kw, sink_headroom=64 * 8 * kw["n_blocks"] + rng.choice([8, 64, 256]), sink_rows=int(o.solutions) + 8)
    r, sols = util.emu_search(m, **kw)
    uniq = {tuple(s) for s in sols}
    orc = util.Oracle(m)
    ok = len(sols) == o.solutions == len(uniq) == r.solutions
    ok = ok and all(orc.leaf_true(np.array([x for v in s for x in (v, v)], np.int32)) for s in list(uniq)[:64])
    return ok, kw, r


def sat_case(rng, backjump):
    n = rng.randint(10, 70)
    ratio = rng.choice([3.0, 3.8, 4.26, 4.26, 4.8])
    cnf = I.random_3sat_cnf(n, ratio, rng.randint(1, 10**6))
    obj = rng.choice(["ANY", "ANY", "ALL", "MIN"])
    if obj == "ALL" and n > 34:
        obj = "ANY"
    if obj == "MIN":
        n = min(n, 26)
        cnf = [cl for cl in cnf if all(abs(l) <= n for l in cl)] or [[1, 2, 3]]
        text = I.cnf_to_csolve(n, cnf, "MIN " + " + ".join("x%d" % i for i in range(1, n + 1)))
    else:
        text = I.cnf_to_csolve(n, cnf, obj)
    m = cb.Model(text)
    o, _ = util.Oracle(m).solve_tree(0)
    learn = rng.random() < 0.7
    kw = dict(order=0 if obj == "ALL" else rng.randint(0, 4), learn=learn, backjump=learn and backjump,
              prefer_failing=obj != "ALL" and rng.random() < 0.4, n_blocks=rng.choice([1, 1, 2, 3]), max_solutions=16,
              general=rng.random() < 0.5, slice_clock=rng.choice([0, 0, 2000, 10000, 50000]))
    if obj == "ALL":
        o, _ = util.Oracle(m).solve_tree(kw["order"])
    if obj == "ALL" and not learn and o.solutions < 20000 and rng.random() < 0.5:
        ok, kw, r = sink_check(m, o, kw, rng)
        return ok, "sat n=%d ALL sink %s expected %s" % (n, kw, o.solutions), r
    if obj == "ANY" and kw["prefer_failing"] and not learn and rng.random() < 0.6:
        kw["restart_frequency"] = rng.choice([1, 2, 5, 20])         # Luby restarts (-r), ANY models without learning
    if obj == "ALL" and not learn:
        r, sols, kw = ranks_search(m, kw, rng)
    else:
        r, sols = util.emu_search(m, split_target=rng.choice([1, 1, 8, 64]), **kw)
    if obj == "ALL":
        # learned nogoods cut nodes the plain tree expands: with learning only the solution count is the tree's
        ok = r.solutions == o.solutions if learn else (r.solutions, r.nodes, r.cuts) == (o.solutions, o.calls, o.cuts)
    elif obj == "ANY":
        ok = r.has_solution == (1 if o.solutions > 0 else 0)
        if ok and r.has_solution:
            val = dict(zip(m.var_names, sols[0]))
            ok = all(any((val["x%d" % abs(l)] == 1) == (l > 0) for l in cl) for cl in cnf)
    else:
        ok = r.has_solution == o.has_solution and (not o.has_solution or r.best == o.best)
        if ok and r.has_solution:
            val = dict(zip(m.var_names, sols[0]))
            xs = [val["x%d" % i] for i in range(1, n + 1)]
            ok = sum(xs) == r.best and all(any((val["x%d" % abs(l)] == 1) == (l > 0) for l in cl) for cl in cnf)
    return ok, "sat n=%d %s %s expected %s" % (n, obj, kw, (o.solutions, o.calls, o.cuts, o.best)), r


def queens_case(rng, backjump):
    """N-queens / sudoku: the lane-owns-variable and K-per-lane kernels"""
    if rng.random() < 0.6:
        n = rng.randint(4, 9)
        text = I.queens(n, rng.choice(["ALL", "ALL", "ANY"]))
    else:
        text = I.sudoku(I.sudoku_puzzle(random.Random(rng.randint(0, 10**6)), rng.randint(24, 32)), rng.choice(["ALL", "ANY"]))
    m = cb.Model(text)
    order = rng.randint(0, 4)
    o, _ = util.Oracle(m).solve_tree(order)
    kw = dict(order=order, n_blocks=rng.choice([1, 2, 3]), general=rng.random() < 0.25, slice_clock=rng.choice([0, 0, 2000, 10000, 50000]))
    if text.startswith("ALL") and rng.random() < 0.4:
        ok, kw, r = sink_check(m, o, kw, rng)
        return ok, "queens/sudoku ALL sink %s expected %s" % (kw, o.solutions), r
    if text.startswith("ALL"):
        r, sols, kw = ranks_search(m, kw, rng)
    else:
        r, sols = util.emu_search(m, split_target=rng.choice([1, 1, 8, 64]), **kw)
    if text.startswith("ALL"):
        ok = (r.solutions, r.nodes, r.cuts) == (o.solutions, o.calls, o.cuts)
    else:
        ok = r.has_solution == (1 if o.solutions > 0 else 0)
        if ok and r.has_solution:
            # the assignment satisfies the model: as the only root domain it leaves exactly one solution
            orc = util.Oracle(m)
            dom = [x for v in sols[0] for x in (v, v)]
            import numpy as np
            ok = bool(orc.leaf_true(np.array(dom, np.int32)))
    return ok, "queens/sudoku %s %s" % (text.split(";")[0], kw), r


def comm_case(rng, backjump):
    """ANY / MIN models on 2-3 emulated GPUs of a csolve_gpu_comm (util.emu_search_comm)"""
    n = rng.randint(10, 60)
    cnf = I.random_3sat_cnf(n, rng.choice([3.0, 3.8, 4.26, 4.8]), rng.randint(1, 10**6))
    obj = rng.choice(["ANY", "MIN"])
    if obj == "MIN":
        n = min(n, 24)
        cnf = [cl for cl in cnf if all(abs(l) <= n for l in cl)] or [[1, 2, 3]]
        text = I.cnf_to_csolve(n, cnf, "MIN " + " + ".join("x%d" % i for i in range(1, n + 1)))
    else:
        text = I.cnf_to_csolve(n, cnf, obj)
    m = cb.Model(text)
    o, _ = util.Oracle(m).solve_tree(0)
    kw = dict(world=rng.choice([2, 2, 3, 4]), order=rng.randint(0, 4), prefer_failing=rng.random() < 0.3, n_blocks=rng.choice([1, 1, 2]),
              split_target=rng.choice([1, 8, 32, 128]), slice_clock=rng.choice([0, 2000, 10000, 50000]), general=rng.random() < 0.5)
    world = kw.pop("world")
    if rng.random() < 0.4:
        kw["learn"] = True                      # -c under -j N: every rank learns into a pool of its own
        kw["backjump"] = backjump
    r, w = util.emu_search_comm(m, world, **kw)
    if obj == "ANY":
        ok = r.has_solution == (1 if o.solutions > 0 else 0) and (w is not None) == bool(r.has_solution)
    else:
        ok = r.has_solution == o.has_solution and (not o.has_solution or (r.best == o.best and w is not None))
    if ok and w is not None:
        val = dict(zip(m.var_names, w))
        ok = all(any((val["x%d" % abs(l)] == 1) == (l > 0) for l in cl) for cl in cnf)
        if obj == "MIN":
            ok = ok and sum(val["x%d" % i] for i in range(1, n + 1)) == r.best
    return ok, "comm world=%d n=%d %s %s expected %s" % (world, n, obj, kw, (o.solutions, o.best)), r


def generic_case(rng, backjump):
    text = gen_random.gen_instance(rng.randint(0, 10**7))
    try:
        m = cb.Model(text)
    except cb.CsolveError:
        return True, "rejected", None
    hc = util.harness_lib()
    if hc.hc_load(m.flat, 1) != 0:
        return True, "unsupported", None
    obj = text.split(";")[0].split()[0]
    order = rng.randint(0, 4)
    o, _ = util.Oracle(m).solve_tree(order)
    if o.hit_limit:
        return True, "too large", None
    learn = rng.random() < 0.3
    kw = dict(order=order, learn=learn, backjump=learn and backjump, n_blocks=rng.choice([1, 2]), max_solutions=16,
              general=rng.random() < 0.5, slice_clock=rng.choice([0, 0, 2000, 10000, 50000]))
    if obj == "ALL" and not learn:
        r, sols, kw = ranks_search(m, kw, rng)
    else:
        r, sols = util.emu_search(m, split_target=rng.choice([1, 1, 8, 64]), **kw)
    if obj == "ALL":
        ok = r.solutions == o.solutions if learn else (r.solutions, r.nodes, r.cuts) == (o.solutions, o.calls, o.cuts)
    elif obj == "ANY":
        ok = r.has_solution == (1 if o.solutions > 0 else 0)
    else:
        ok = r.has_solution == o.has_solution and (not o.has_solution or r.best == o.best)
    return ok, "generic %s %s expected %s\n%s" % (obj, kw, (o.solutions, o.calls, o.cuts, o.best), text), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=200)
    ap.add_argument("--seed0", type=int, default=0)
    ap.add_argument("--backjump", action="store_true")
    ap.add_argument("--kind", default="both")
    a = ap.parse_args()
    t0 = time.time()
    bad = 0
    tot = {"nodes": 0, "conflicts": 0, "backjumps": 0, "cases": 0}
    for s in range(a.seed0, a.seed0 + a.seeds):
        rng = random.Random(s)
        fn = ({"sat": sat_case, "generic": generic_case, "queens": queens_case, "comm": comm_case}.get(a.kind)
              or (sat_case, generic_case, queens_case, sat_case, comm_case)[s % 5])
        ok, what, r = fn(rng, a.backjump)
        if r is not None:
            tot["cases"] += 1; tot["nodes"] += r.nodes; tot["conflicts"] += r.conflicts; tot["backjumps"] += r.backjumps
        if not ok:
            bad += 1
            print("MISMATCH seed %d: %s -> %s" % (s, what, (r.solutions, r.nodes, r.cuts, r.best, r.has_solution)), flush=True)
    print("emu_fuzz: %d seeds, %d searches, %d nodes, %d nogoods, %d back-jumps, %d mismatches, %.0f s"
          % (a.seeds, tot["cases"], tot["nodes"], tot["conflicts"], tot["backjumps"], bad, time.time() - t0), flush=True)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
