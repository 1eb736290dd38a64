"""The warp-cooperative contractor for linear clauses (x_obj == konst + SUM (+-) k_i * x_i, contract.cuh:
lin_lane_load / lin_lane_apply; on the device one term per lane, here the lanes are emulated by the host harness):
its fixpoint must be the one of the reference's nested propagate_eq / add / neg / mul calls, which the oracle
restates -- fail flags and post-fixpoint domains on seeded random walks over generated MIN / MAX models."""
import ctypes as C
import random

import numpy as np
import pytest

import csolve_b200 as cb
import util


def linear_model(rng):
    n = rng.randint(3, 9)
    names = ["v%d" % i for i in range(n)]
    lines = []
    terms = []
    for _ in range(rng.randint(2, 12)):
        x = rng.choice(names)
        k = rng.choice([-7, -3, -2, -1, 0, 1, 1, 2, 3, 4, 5, 9])
        form = rng.randint(0, 4)
        if form == 0:
            t = x
        elif form == 1:
            t = "%d*%s" % (k, x)
        elif form == 2:
            t = "%s*%d" % (x, k)
        elif form == 3:
            t = "%d" % rng.randint(-20, 20)
        else:
            t = "%d*%s" % (abs(k), x)
        terms.append((rng.choice("+-") if terms else "", t))
    expr = " ".join((s + " " + t).strip() for s, t in terms)
    lines.append("%s %s;" % (rng.choice(["MIN", "MAX"]), expr))
    for x in names:
        lo = rng.randint(-6, 3)
        lines.append("%d <= %s; %s <= %d;" % (lo, x, x, lo + rng.randint(0, 12)))
    for _ in range(rng.randint(0, 4)):
        a, b = rng.sample(names, 2)
        lines.append(rng.choice(["%s != %s;", "%s <= %s;", "%s + %s <= 9;", "%s < %s + 3;"]) % (a, b))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(12))
def test_linear_clause_fixpoint_equals_oracle(seed):
    rng = random.Random(1000 + seed)
    hc = util.harness_lib()
    checked = with_linear = 0
    for _ in range(12):
        text = linear_model(rng)
        try:
            m = cb.Model(text)
        except cb.CsolveError:
            continue                                  # infeasible at root
        assert hc.hc_load(C.byref(m.flat), 1) == 0, hc.hc_error()
        if hc.hc_n_linear() == 0:
            continue
        with_linear += 1
        orc = util.Oracle(m)
        V = m.n_vars
        for _walk in range(20):
            dom = m.root_domains.copy()
            best = 2**31 - 1 if m.objective == cb.OBJ_MIN else -2**31
            free = [v for v in range(V) if v != m.obj_var]
            rng.shuffle(free)
            for v in free:
                lo, hi = int(dom[2 * v]), int(dom[2 * v + 1])
                val = rng.randint(lo, hi)
                if rng.random() < 0.3:                # an incumbent from somewhere: objective tightening in the node
                    best = rng.randint(-60, 60)
                eo, ef = orc.node(dom, v, val, best)
                out = np.empty_like(dom)
                hf = hc.hc_node(util.p32(np.ascontiguousarray(dom, np.int32)), v, val, best, util.p32(out))
                assert bool(hf) == bool(ef), (text, dom.tolist(), v, val, best)
                checked += 1
                if ef:
                    break
                assert np.array_equal(out, eo), (text, dom.tolist(), v, val, best, out.tolist(), eo.tolist())
                dom = eo
    assert with_linear >= 3 and checked > 100
