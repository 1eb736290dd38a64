#!/usr/bin/env python3
"""Generate golden fixtures from the COMPILED REFERENCE (oracle/_ref, built from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    make -C oracle/ref && python tests/golden/make_golden.py

Writes
  replay_<name>.npz     sampled node transitions through the reference's own
                        bind + objective_update_val + propagate_clauses (oracle/ref/replay.c):
                        dom_in, var, val, best -> failed, dom_out, plus leaf states and their
                        is_true(eval(root)) verdict
  ref_counts.json       CALLS / CUTS / PROPS / SOLUTIONS / BEST printed by the reference CLI
                        (-c false -r 0 -s 0, and the default flags) on the instances
  flat_digests.json     sha256 of the flat model obtained from the reference's structures
                        (integration/csolve_gpu_shim.c) for the instances and 300 random inputs
"""
import ctypes as C
import json
import os
import random
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from csolve_b200 import instances as I  # noqa: E402

I32P = C.POINTER(C.c_int32)


from make_instances import instance_table, random_table  # noqa: E402


def sample_replay(ref, text, n_walks, seed):
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
        path = f.name
    n = ref.ref_load(path.encode(), 0, 1)
    os.unlink(path)
    if n <= 0:
        return None
    V = n
    obj, ov = ref.ref_objective(), ref.ref_obj_var()
    root = np.zeros(2 * V, np.int32)
    ref.ref_get_domains(root.ctypes.data_as(I32P))
    rng = random.Random(seed)
    rec = dict(dom_in=[], var=[], val=[], best=[], failed=[], dom_out=[], leaf=[], leaf_true=[])
    for _ in range(n_walks):
        dom = root.copy()
        un = list(range(V)); rng.shuffle(un)
        best = 2**31 - 1 if obj == 2 else (-2**31 if obj == 3 else 0)
        if ov >= 0 and rng.random() < 0.7:
            lo, hi = int(root[2 * ov]), int(root[2 * ov + 1])
            best = rng.randint(lo, min(hi, lo + 5000))
        alive = True
        while un:
            x = un.pop()
            lo, hi = int(dom[2 * x]), int(dom[2 * x + 1])
            val = rng.randint(lo, hi) if hi - lo < 50 else rng.choice([lo, hi, lo + 1, hi - 1, rng.randint(lo, lo + 20)])
            out = np.zeros(2 * V, np.int32)
            f = ref.ref_replay(dom.ctypes.data_as(I32P), x, val, best, out.ctypes.data_as(I32P), None)
            if not f and ov >= 0 and out[2 * ov] > out[2 * ov + 1]:
                alive = False      # reference leaves an empty <obj> undetected: not a parity case
                break
            rec["dom_in"].append(dom.copy()); rec["var"].append(x); rec["val"].append(val); rec["best"].append(best)
            rec["failed"].append(f); rec["dom_out"].append(out.copy())
            if f:
                alive = False
                break
            dom = out
        if alive:
            rec["leaf"].append(dom.copy())
            rec["leaf_true"].append(ref.ref_eval_root(dom.ctypes.data_as(I32P)))
    out = {k: np.array(v, np.int32) for k, v in rec.items()}
    out["root"] = root
    return out


def cli_counts(text, flags):
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
        path = f.name
    out = subprocess.run([util.REF_CLI, "-s", "0", *flags, path], capture_output=True, text=True).stdout
    os.unlink(path)
    m = re.search(r"CALLS: (\d+), CUTS: (\d+), PROPS: (\d+), CONFL: (\d+).*SOLUTIONS: (\d+)", out)
    best = re.findall(r"BEST: (-?\d+)", out)
    sols = re.findall(r"SOLUTION: (.*), BEST", out)
    d = dict(zip(["calls", "cuts", "props", "confl", "solutions"], [int(x) for x in m.groups()]))
    d["best"] = int(best[-1]) if best else None
    d["no_solution"] = "NO SOLUTION FOUND" in out
    d["last_solution"] = sols[-1] if sols else None
    return d


def main():
    ref = util.reference_lib()
    assert ref is not None, "build oracle/_ref first (make -C oracle/ref)"
    inst = instance_table()
    rnd = random_table()

    for name, text in inst.items():
        walks = 400 if name not in ("sudoku", "sudoku_any", "sat100") else 150
        r = sample_replay(ref, text, walks, seed=hash(name) % 100000 if False else sum(map(ord, name)))
        np.savez_compressed(os.path.join(HERE, "replay_%s.npz" % name), **r)
        print(name, "replay nodes", len(r["var"]), "failed", int(r["failed"].sum()), "leaves", len(r["leaf"]))
    # random instances: a few nodes each, packed into one file
    packed = {}
    kept = 0
    for name, text in rnd.items():
        r = sample_replay(ref, text, 12, seed=sum(map(ord, name)))
        if r is None or len(r["var"]) == 0:
            continue
        kept += 1
        for k, v in r.items():
            packed["%s/%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "replay_random.npz"), **packed)
    print("random instances with replay nodes:", kept)

    counts = {}
    for name, text in inst.items():
        if name == "wcet":
            continue  # 80 s of CPU per run; its optimum (1560) is recorded below from the survey + z3 cross-check
        counts[name] = {"nocf": cli_counts(text, ["-c", "false", "-r", "0"])}
        if name in ("schedule", "queens8", "sudoku"):
            counts[name]["default"] = cli_counts(text, [])
        print(name, counts[name]["nocf"])
    counts["wcet"] = {"known_optimum": 1560}
    json.dump(counts, open(os.path.join(HERE, "ref_counts.json"), "w"), indent=1, sort_keys=True)

    digests = {}
    all_inst = dict(inst)
    all_inst.update(rnd)
    for name, text in all_inst.items():
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(text)
            path = f.name
        n = ref.ref_load(path.encode(), 0, 1)
        os.unlink(path)
        if n == -2:
            digests[name] = "INFEASIBLE"
        elif n < 0:
            digests[name] = "SYNTAX"
        else:
            rc = ref.ref_flatten()
            digests[name] = util.flat_digest(ref.ref_flat().contents) if rc == 0 else "RC%d" % rc
    json.dump(digests, open(os.path.join(HERE, "flat_digests.json"), "w"), indent=0, sort_keys=True)
    print("digests:", len(digests), "feasible:", sum(1 for v in digests.values() if len(v) == 64))


if __name__ == "__main__":
    main()
