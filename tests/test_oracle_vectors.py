"""The reference's own unit-test value tables (tests/golden/ref_unit_vectors.json, extracted from
/root/reference/test/test_arith.c, test_eval.c, test_propagate.c) against
  * the oracle (oracle/csolve_oracle.c) and
  * the product's per-lane contractor code (csolve_b200/csrc/contract.cuh) run on the host.
"""
import numpy as np
import pytest

import util

VEC = util.load_vectors()
DMIN, DMAX = util.DMIN, util.DMAX


@pytest.mark.parametrize("case", VEC["arith"], ids=lambda c: "%s%s" % (c["op"], c["args"]))
def test_arith(case):
    orc, hc = util.oracle_lib(), util.harness_lib()
    f = {"neg": orc.orc_neg, "add": orc.orc_add, "mul": orc.orc_mul, "min": orc.orc_min, "max": orc.orc_max}[case["op"]]
    assert f(*case["args"]) == case["expect"]
    g = {"neg": hc.hc_sneg, "add": hc.hc_sadd, "mul": hc.hc_smul}.get(case["op"])
    if g is not None:
        assert g(*case["args"]) == case["expect"]


def test_arith_exhaustive_corners():
    """oracle (bit tricks of arith.c) == product (64-bit clamp) on all corner pairs"""
    orc, hc = util.oracle_lib(), util.harness_lib()
    pts = [DMIN, DMIN + 1, DMIN + 2, -65536, -3, -2, -1, 0, 1, 2, 3, 46341, 65536, DMAX - 2, DMAX - 1, DMAX]
    for a in pts:
        assert orc.orc_neg(a) == hc.hc_sneg(a)
        for b in pts:
            assert orc.orc_add(a, b) == hc.hc_sadd(a, b), (a, b)
            assert orc.orc_mul(a, b) == hc.hc_smul(a, b), (a, b)


@pytest.mark.parametrize("case", VEC["eval"], ids=lambda c: "%s-%s-%s-%s" % (c["test"], c["op"], c["l"], c["r"]))
def test_eval(case):
    doms = [case["l"]] + ([case["r"]] if case["r"] is not None else [])
    hm = util.HandModel(case["op"], len(doms), doms)
    dom = np.array([x for d in doms for x in d], np.int32)
    assert util.Oracle(hm).eval_root(dom) == case["expect"]
    hc = util.harness_lib()
    assert hc.hc_load(hm.flat, 0) == 0, hc.hc_error()
    out = np.zeros(2, np.int32)
    hc.hc_eval_root(util.p32(dom), util.p32(out))
    assert out.tolist() == case["expect"]


def _expected_domains(case, names):
    """apply the bind() calls the reference test expects, in order"""
    doms = {n: list(case["terms"][n]) for n in names}
    for b in case["binds"]:
        doms[b["term"]] = list(b["val"])
    return doms


@pytest.mark.parametrize("case", VEC["propagate"],
                         ids=lambda c: "%s-%s-%s-%s" % (c["test"], c["op"], c["terms"], c["val"]))
def test_propagate(case):
    names = [case["l"]] + ([case["r"]] if case["r"] else [])
    if len(names) == 2 and names[0] == names[1]:
        pytest.skip("same terminal on both sides")
    doms = [case["terms"][n] for n in names]
    hm = util.HandModel(case["op"], len(names), doms)
    dom = np.array([x for d in doms for x in d], np.int32)
    out, res = util.Oracle(hm).prop_root(dom, case["val"][0], case["val"][1])
    hc = util.harness_lib()
    assert hc.hc_load(hm.flat, 0) == 0, hc.hc_error()
    out2 = np.zeros_like(dom)
    res2 = hc.hc_prop_root(util.p32(dom), case["val"][0], case["val"][1], util.p32(out2))
    assert (res == -1) == (case["result"] == -1)
    assert (res2 == -1) == (case["result"] == -1)
    if case["result"] == -1:
        return
    # In the reference tests only terminals with an env are "variables" (bind() is mocked and
    # expected explicitly); anonymous interval terminals are narrowed silently. The expected
    # value of a variable is what the expected bind() calls leave behind.
    exp = _expected_domains(case, names)
    for i, n in enumerate(names):
        if n in case["vars"]:
            assert out[2 * i:2 * i + 2].tolist() == exp[n], (n, out.tolist())
            assert out2[2 * i:2 * i + 2].tolist() == exp[n], (n, out2.tolist())
    # change count: every terminal is a variable here, so the count can only be >= the reference's
    if all(n in case["vars"] or case["terms"][n][0] == case["terms"][n][1] for n in names):
        assert res == case["result"]


def test_nogood_propagation_vectors():
    """PropagateConfl.* of the reference (test/test_propagate.c:1011-1155), value = true: the nogood
    {A = 0, B = 1} against (A, B) domains -> expected change of B / no change"""
    hc = util.harness_lib()
    lits = np.array([(0 << 1) | 0, (1 << 1) | 1], np.int32)
    cases = [
        ([0, 0, 0, 1], [0, 0, 0, 0], 1),      # Basic: A matches, B open -> B loses 1
        ([0, 0, 0, 0], [0, 0, 0, 0], 0),      # NonConfl: B is 0 != 1 -> nothing
        ([0, 1, 0, 1], [0, 1, 0, 1], 0),      # TwoVars: two open variables -> nothing
        ([1, 1, 0, 1], [1, 1, 0, 1], 0),      # A is 1 != 0 -> nothing
        ([0, 0, 1, 1], [0, 0, 1, 1], 0),      # everything matches: NOT an error (src/propagate.c:459-471)
        ([0, 0, 0, 5], [0, 0, 0, 5], 0),      # the recorded value sits on neither bound of B
        ([0, 0, 1, 5], [0, 0, 2, 5], 1),      # ... on the lower bound
    ]
    for dom, exp, changed in cases:
        d = np.array(dom, np.int32)
        out = np.zeros_like(d)
        r = hc.hc_prop_nogood(util.p32(d), 2, util.p32(lits), 2, util.p32(out))
        assert r == changed and out.tolist() == exp, (dom, out.tolist(), r)
