#!/usr/bin/env python3
"""Counters of the search TREE (solutions, nodes, cuts) from the oracle (oracle/csolve_oracle.c: orc_solve_tree),
at the sizes where tests cannot afford to run it: queens 11..14 in all five variable orders, queens 15 / 16 in the
static order (about an hour and a half of CPU, spread over the host's cores by root-level value), random 3-SAT
n=200 seed 1 (unsatisfiable: the whole tree) and the per-root counters of the first 200 generated sudokus.

    python tests/golden/make_tree_counts.py [--only queens|sat|sudoku] [--max-queens 16] [--jobs 7]

Writes tests/golden/tree_counts.json (merged with what is already there). The `-m gpu` tests assert that the
search kernels report exactly these counters (tests/test_gpu_tree_counts.py); the CPU suite re-derives the small
entries.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(HERE, "tree_counts.json")
ORDER_NAMES = ["none", "smallest-domain", "largest-domain", "smallest-value", "largest-value"]


def _part(job):
    text, order, part, n_parts, split_level = job
    import csolve_b200 as cb
    import util
    m = cb.Model(text)
    r = util.Oracle(m).solve_tree_part(order, part, n_parts, split_level)
    return [int(r.solutions), int(r.calls), int(r.cuts), int(r.props)]


def tree(pool, text, order, n_parts, split_level=0):
    t0 = time.time()
    parts = pool.map(_part, [(text, order, k, n_parts, split_level) for k in range(n_parts)], chunksize=1)
    tot = [sum(p[i] for p in parts) for i in range(3)]
    return dict(solutions=tot[0], nodes=tot[1], cuts=tot[2], oracle_seconds=round(time.time() - t0, 1))


def _sudoku(job):
    grid, order = job
    import csolve_b200 as cb
    import util
    from csolve_b200 import instances as I
    m = cb.Model(I.sudoku(grid))
    r, _ = util.Oracle(m).solve_tree(order)
    return [int(r.solutions), int(r.calls), int(r.cuts)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--max-queens", type=int, default=16)
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    a = ap.parse_args()
    from csolve_b200 import instances as I
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}

    def save():
        json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)

    with mp.Pool(a.jobs) as pool:
        if a.only in ("", "sudoku"):
            grids = I.sudoku_batch(200, seed=20261018)
            for order in (1,):
                rows = pool.map(_sudoku, [(g, order) for g in grids], chunksize=4)
                data["sudoku_batch200_seed20261018/%s" % ORDER_NAMES[order]] = dict(
                    per_root=rows, solutions=sum(r[0] for r in rows), nodes=sum(r[1] for r in rows), cuts=sum(r[2] for r in rows))
                save()
                print("sudoku batch", ORDER_NAMES[order], data["sudoku_batch200_seed20261018/%s" % ORDER_NAMES[order]]["nodes"], flush=True)
        if a.only in ("", "queens"):
            for n in range(11, a.max_queens + 1):
                for order in range(5):
                    if n >= 15 and order != 0:
                        continue
                    key = "queens%d/%s" % (n, ORDER_NAMES[order])
                    if key in data:
                        continue
                    data[key] = tree(pool, I.queens(n), order, 4 * n, 1)
                    save()
                    print(key, data[key], flush=True)

        if a.only in ("", "sat"):
            for seed in (1,):
                key = "sat200_seed%d/none" % seed
                data[key] = tree(pool, I.random_3sat(200, seed=seed), 0, 48, 8)
                save()
                print(key, data[key], flush=True)


if __name__ == "__main__":
    main()
