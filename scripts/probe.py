"""Development probe: timings of the search kernels on one GPU, including one rank's share of an N-way partition
(part_count = N on a single device reproduces exactly what each rank of an N-GPU run does).
usage: probe.py [queens N] [parts P] [sat] [wcet] [sudoku]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import instances as I

args = sys.argv[1:]


def run(name, text, reps=3, **kw):
    m = cb.Model(text)
    p = cb.GpuProblem(m)
    p.solve(**kw)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = p.solve(**kw)
        w = (time.perf_counter() - t0) * 1e3
        if best is None or r.kernel_ms + r.expand_ms < best[0].kernel_ms + best[0].expand_ms:
            best = (r, w)
    r, w = best
    print("%-28s sols=%d nodes=%d cuts=%d best=%d search_ms=%.2f expand_ms=%.2f wall_ms=%.2f launches=%d Gnodes/s=%.3f" % (
        name, r.solutions, r.nodes, r.cuts, r.best, r.kernel_ms, r.expand_ms, w, r.kernel_launches,
        r.nodes / max(r.kernel_ms + r.expand_ms, 1e-9) / 1e6), flush=True)
    return r


n = int(args[args.index("queens") + 1]) if "queens" in args else 16
run("queens%d" % n, I.queens(n))
if "parts" in args:
    P = int(args[args.index("parts") + 1])
    worst = 0.0
    for rk in range(P):
        r = run("queens%d part %d/%d" % (n, rk, P), I.queens(n), part_rank=rk, part_count=P)
        worst = max(worst, r.kernel_ms + r.expand_ms)
    print("slowest rank of %d: %.2f ms" % (P, worst))
if "orders" in args:
    for o in ("smallest-domain", "largest-domain"):
        run("queens%d -o %s" % (n - 2, o), I.queens(n - 2), order=o)
if "wcet" in args:
    run("wcet", I.wcet())
    run("schedule", I.schedule())
if "sat" in args:
    for seed in (1, 2, 3):
        run("sat200 seed %d" % seed, I.random_3sat(200, seed=seed), prefer_failing=True)
if "satstatic" in args:
    run("sat200 seed 1 static", I.random_3sat(200, seed=1), reps=1)
if "sudoku" in args:
    grids = I.sudoku_batch(10000, base=50)
    m = cb.Model(I.sudoku("." * 81))
    p = cb.GpuProblem(m)
    roots = I.sudoku_roots(m.var_names, grids)
    for _ in range(3):
        t0 = time.perf_counter()
        r, counts, failed = p.solve_batch(roots, order="smallest-domain")
        w = (time.perf_counter() - t0) * 1e3
        print("sudoku x10000 nodes=%d search_ms=%.2f expand_ms=%.2f wall_ms=%.2f ok=%s" % (
            r.nodes, r.kernel_ms, r.expand_ms, w, counts.tolist() == [1] * len(grids)), flush=True)
if "split" in args:
    P = int(args[args.index("split") + 1])
    for mult in (8, 16, 32, 64, 128, 256):
        st = 4736 * mult
        run("queens%d split %dx" % (n, mult), I.queens(n), split_target=st)
        if P > 1:
            run("queens%d split %dx part 0/%d" % (n, mult, P), I.queens(n), split_target=st, part_rank=0, part_count=P)
if "wcetvar" in args:
    for kw in ({"slice_ms": 1}, {"slice_ms": 2}, {"slice_ms": 3}, {"slice_ms": 5}, {"slice_ms": 10}):
        for _ in range(2):
            run("wcet %s" % kw, I.wcet(), reps=2, **kw)
