"""Search nodes sampled out of the real search kernels (csolve_solve_options.sample_mod) and how they are checked.

BASELINE.md §4.5 / north_star: "identical post-fixpoint domains on sampled nodes replayed through the reference
propagate()". The search kernels (k_search, k_search_lov, k_search_lovk) record every node -- executed or counted by
one of their bulk shortcuts -- whose identity hash is 0 modulo sample_mod: parent domains, decision, incumbent, fail
flag, post-fixpoint domains. Three checkers look at such records:

  * the oracle, live on the GPU box          (tests/test_gpu_search_samples.py)
  * the COMPILED REFERENCE in the build container: tests/golden/make_search_samples.py replays records collected on a
    B200 (scripts/collect_search_samples.py) through oracle/_ref's ref_replay() and commits its answers as
    tests/golden/search_<name>.npz -- for the deterministic trees (ALL models, static tree) the kernel must reproduce
    those record sets exactly, for the others the fixture parents are pushed through csolve_gpu_propagate_batch
  * the oracle against those fixtures on the CPU (tests/test_search_sample_fixtures.py), which pins the oracle on
    search-sampled nodes at BASELINE sizes.
"""
import os

import numpy as np

import csolve_b200 as cb
from csolve_b200 import instances as I

FAILED, COUNTED, LEAF = cb.host.SAMPLE_FAILED, cb.host.SAMPLE_COUNTED, cb.host.SAMPLE_LEAF
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# name -> how the instance is searched when its nodes are sampled.
#   text        csolve input            batch     number of generated sudoku roots searched over the one network
#   mod, fkeep  sampling rate           determ    the tree (hence the sampled set) does not depend on timing
#   env         development switches    solve     extra solve() arguments
SAMPLED = {
    "queens14":    dict(text=lambda: I.queens(14), mod=150, fkeep=12, determ=True),
    "queens14_sd": dict(text=lambda: I.queens(14), mod=150, fkeep=12, determ=True, solve=dict(order=cb.ORDER_SMALLEST_DOMAIN)),
    "queens15":    dict(text=lambda: I.queens(15), mod=900, fkeep=12, determ=True),
    "queens16":    dict(text=lambda: I.queens(16), mod=6000, fkeep=12, determ=True),
    "queens12_general": dict(text=lambda: I.queens(12), mod=12, fkeep=4, determ=True, env={"CSOLVE_NO_LOV": "1"}),
    "sudoku10k":   dict(text=lambda: I.sudoku("." * 81), batch=10000, mod=40, fkeep=4, determ=True,
                        solve=dict(order=cb.ORDER_SMALLEST_DOMAIN)),
    "schedule":    dict(text=I.schedule, mod=1, fkeep=1, determ=False),
    "wcet":        dict(text=I.wcet, mod=2, fkeep=8, determ=False),
    # static order (no failure-driven priorities): seed 1 is unsatisfiable -- the whole tree, 451 592 233 nodes;
    # seeds 2 and 3 are satisfiable -- ANY is a race, seed 2 is cut short by the time limit
    "sat200_s1":   dict(text=lambda: I.random_3sat(200, seed=1), mod=12000, fkeep=2, determ=True),
    "sat200_s2":   dict(text=lambda: I.random_3sat(200, seed=2), mod=8000, fkeep=2, determ=False, solve=dict(time_limit_ms=1500)),
    "sat200_s3":   dict(text=lambda: I.random_3sat(200, seed=3), mod=3000, fkeep=2, determ=False, solve=dict(time_limit_ms=4000)),
    "sat200_s1_pf": dict(text=lambda: I.random_3sat(200, seed=1), mod=300, fkeep=2, determ=False, solve=dict(prefer_failing=True)),
}


def sudoku_grids(n):
    return I.sudoku_batch(n, base=100) if n > 1000 else I.sudoku_batch(n, seed=20261018)


def run_sampled(name, cap=1 << 18, mod=None):
    """search instance `name` with sampling on; returns (model, result, samples dict)"""
    cfg = SAMPLED[name]
    old = {k: os.environ.get(k) for k in cfg.get("env", {})}
    os.environ.update(cfg.get("env", {}))
    try:
        m = cb.Model(cfg["text"]())
        p = cb.GpuProblem(m)
        kw = dict(cfg.get("solve", {}))
        kw.update(sample_mod=mod or cfg["mod"], sample_failed_keep=cfg["fkeep"], sample_cap=cap)
        if cfg.get("batch"):
            roots = I.sudoku_roots(m.var_names, sudoku_grids(cfg["batch"]))
            r, _, _ = p.solve_batch(roots, **kw)
        else:
            r = p.solve(**kw)
        s = p.samples(cap)
        p.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return m, r, s


def check_against_oracle(model, s, oracle=None, limit=None):
    """every record against the oracle's node transition; returns (n_checked, n_nonfailed, list of mismatches)"""
    import util
    orc = oracle or util.Oracle(model)
    ov = model.obj_var
    bad = []
    n = len(s["var"]) if limit is None else min(limit, len(s["var"]))
    nonfailed = 0
    for i in range(n):
        out, f = orc.node(s["parent"][i], int(s["var"][i]), int(s["val"][i]), int(s["best"][i]))
        if not f and ov >= 0 and out[2 * ov] > out[2 * ov + 1]:
            f = 1                           # the device fails a node whose <obj> became empty
        kf = bool(s["flags"][i] & FAILED)
        if kf != bool(f):
            bad.append((i, "fail flag", kf, bool(f)))
            continue
        if f:
            continue
        nonfailed += 1
        if not np.array_equal(out, s["child"][i]):
            bad.append((i, "domains", s["child"][i].tolist(), out.tolist()))
        if s["flags"][i] & LEAF:
            if not (orc.leaf_true(out) and np.array_equal(out[0::2], out[1::2])):
                bad.append((i, "leaf", None, None))
    return n, nonfailed, bad


def canonical(s):
    """records as a sorted array of unique rows (flags, var, val, best, parent..., child... with failed children zeroed)"""
    child = s["child"].copy()
    child[(s["flags"] & FAILED) != 0] = 0
    rows = np.concatenate([s["flags"][:, None], s["var"][:, None], s["val"][:, None], s["best"][:, None],
                           s["parent"], child], axis=1).astype(np.int64)
    rows = np.unique(rows, axis=0)
    return rows


def _small(a):
    a = np.asarray(a)
    if a.size and a.min() >= -128 and a.max() <= 127:
        return a.astype(np.int8)
    if a.size and a.min() >= -32768 and a.max() <= 32767:
        return a.astype(np.int16)
    return a.astype(np.int32)


def save_samples(path, s, **extra):
    child = s["child"].copy()
    child[(s["flags"] & FAILED) != 0] = 0
    np.savez_compressed(path, flags=_small(s["flags"]), var=_small(s["var"]), val=_small(s["val"]), best=np.asarray(s["best"], np.int32),
                        parent=_small(s["parent"]), child=_small(child), **extra)


def load_samples(path):
    z = np.load(path)
    s = {k: z[k] for k in z.files}
    for k in ("flags", "var", "val", "best", "parent", "child"):
        s[k] = s[k].astype(np.int32)
    return s


def fixture_path(name):
    return os.path.join(GOLDEN, "search_%s.npz" % name)
