"""Every BASELINE.json configuration on one GPU, with the reference CPU solver timed beside it on a
bounded sample (oracle/_ref/csolve_ref when it was built, else skipped). Writes a markdown table to stdout."""
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csolve_b200 as cb
from csolve_b200 import instances as I

REF = os.path.join(ROOT, "oracle", "_ref", "csolve_ref")


def ref_run(text, flags=(), timeout=600):
    if not os.path.exists(REF) or "--no-cpu" in sys.argv:      # --no-cpu: GPU columns only (a quick refresh)
        return None
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
    t0 = time.perf_counter()
    try:
        out = subprocess.run([REF, "-s", "0", *flags, f.name], capture_output=True, text=True, timeout=timeout).stdout
    except subprocess.TimeoutExpired:
        os.unlink(f.name)
        return dict(secs=timeout, calls=None, sols=None, best=None, timeout=True)
    dt = time.perf_counter() - t0
    os.unlink(f.name)
    m = re.search(r"CALLS: (\d+).*SOLUTIONS: (\d+)", out)
    best = re.findall(r"BEST: (-?\d+)", out)
    return dict(secs=dt, calls=int(m.group(1)) if m else None, sols=int(m.group(2)) if m else None,
                best=int(best[-1]) if best else None, timeout=False, unsat="NO SOLUTION FOUND" in out)


def gpu_run(text, reps=3, **kw):
    m = cb.Model(text)
    p = cb.GpuProblem(m)
    p.solve(**kw)                      # warm-up
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = p.solve(**kw)
        wall = time.perf_counter() - t0
        if best is None or wall < best[1]:
            best = (r, wall)
    return best


rows = []


def row(name, gpu, cpu, note=""):
    r, wall = gpu
    dev_ms = r.kernel_ms + r.expand_ms
    c = "-"
    sp = "-"
    if cpu:
        c = "%.3f s%s" % (cpu["secs"], " (timeout)" if cpu.get("timeout") else "")
        if cpu["calls"]:
            c += ", %d CALLS" % cpu["calls"]
        sp = "%.0fx" % (cpu["secs"] / wall)
    rows.append("| %s | %d | %d | %s | %.2f | %.2f | %.3g | %s | %s | %s |" % (
        name, r.solutions, r.nodes, r.best if r.has_solution and r.best else "-", dev_ms, wall * 1e3,
        r.nodes / max(dev_ms, 1e-6) * 1e3, c, sp, note))


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    # config 1
    row("queens8 ALL (config 1)", gpu_run(I.queens(8)), ref_run(I.queens(8)))
    # config 3
    for n in (14, 15, 16):
        cpu = ref_run(I.queens(n)) if (n == 14 and not quick) else None
        row("queens%d ALL (config 3)" % n, gpu_run(I.queens(n)), cpu,
            "" if cpu else "CPU: see BASELINE.md (133.5 s / 939.5 s / 6629.7 s in the build container)")
    # config 2: 10 000 sudokus, one resident network
    n_inst = 1000 if quick else 10000
    grids = I.sudoku_batch(n_inst, base=100)
    m = cb.Model(I.sudoku("." * 81))
    p = cb.GpuProblem(m)
    roots = I.sudoku_roots(m.var_names, grids)
    p.solve_batch(roots, order="smallest-domain")          # warm-up (workspace allocation, lazy kernel loading)
    wall = None
    for _ in range(3):
        t0 = time.perf_counter()
        r_, counts, failed = p.solve_batch(roots, order="smallest-domain")
        w_ = time.perf_counter() - t0
        if wall is None or w_ < wall:
            r, wall = r_, w_
    assert counts.tolist() == [1] * n_inst and not failed.any()
    sample = 20
    cpu_secs = 0.0
    cpu_calls = 0
    for g in grids[:sample]:
        c = ref_run(I.sudoku(g))
        if c is None:
            break
        cpu_secs += c["secs"]; cpu_calls += c["calls"]
    cpu = dict(secs=cpu_secs / sample * n_inst, calls=cpu_calls * n_inst // sample) if cpu_secs else None
    row("%d sudokus batched, ALL, -o smallest-domain (config 2)" % n_inst, (r, wall), cpu,
        "CPU: %d instances timed (%.1f ms each incl. process start), extrapolated" % (sample, 1e3 * cpu_secs / sample) if cpu else "")
    # config 4
    row("schedule MIN (config 4)", gpu_run(I.schedule()), ref_run(I.schedule()))
    row("wcet MAX (config 4)", gpu_run(I.wcet()), None if quick else ref_run(I.wcet()), "optimum 1560")
    # config 5
    for seed in (1, 2, 3):
        text = I.random_3sat(200, seed=seed)
        g = gpu_run(text, reps=2, time_limit_ms=120000, prefer_failing=True, restart_frequency=100)
        cpu = None if quick else ref_run(text, ["-c", "false"], timeout=120)
        row("3-SAT n=200 m=852 seed %d ANY (config 5)" % seed, g, cpu,
            ("SAT" if g[0].has_solution else ("TIMEOUT" if g[0].timed_out else "UNSAT")) + "; GPU: -f true -r 100 (the reference's defaults) on k_search_sat; CPU run with -c false (the default -c true needs 1383 s on seed 1, BASELINE.md)")
    print("| instance | solutions | nodes | best | device ms | wall ms | nodes/s (device) | reference CPU (this host, 1 thread) | speed-up (wall) | note |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    print("\n".join(rows))
