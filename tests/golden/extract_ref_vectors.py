#!/usr/bin/env python3
"""Extract the known-answer tables of the reference's own unit tests into JSON.

Reads /root/reference/test/{test_arith.c,test_eval.c,test_propagate.c,test_csolve.c,test_objective.c}
(googletest sources that cannot be built here: gtest/gmock are absent) and writes
tests/golden/ref_unit_vectors.json. Only value tables are taken; no code is copied.

  arith      EXPECT_EQ(expected, op(a, b))                          test/test_arith.c:6-58
  eval       X = CONSTRAINT_EXPR(OP, &L, &R); EXPECT_EQ(v, eval_op(&X))   test/test_eval.c:37-330
  propagate  X = ...; [EXPECT_CALL bind(&e, v, NULL)]; EXPECT_EQ(r, propagate_op(&X, v, NULL))
                                                                     test/test_propagate.c:175-994
Run in the build container (the GPU box has no /root/reference); the JSON is committed.
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DMIN, DMAX = -2**31, 2**31 - 1


def num(s):
    s = s.strip()
    s = s.replace("DOMAIN_MIN", str(DMIN)).replace("DOMAIN_MAX", str(DMAX))
    if not re.fullmatch(r"[-+0-9 ()*]+", s):
        raise ValueError(s)
    return int(eval(s))


def val(s):
    s = s.strip()
    m = re.fullmatch(r"VALUE\((.*)\)", s)
    if m:
        v = num(m.group(1))
        return [v, v]
    m = re.fullmatch(r"INTERVAL\((.*),(.*)\)", s)
    if m:
        return [num(m.group(1)), num(m.group(2))]
    raise ValueError(s)


def tests_of(text):
    for m in re.finditer(r"^TEST\((\w+), (\w+)\) \{\n(.*?)^\}", text, re.M | re.S):
        yield m.group(1), m.group(2), m.group(3)


def arith():
    text = open(os.path.join(REF, "test/test_arith.c")).read()
    out = []
    for m in re.finditer(r"EXPECT_EQ\(([^,]+), (neg|add|mul|min|max)\(([^)]*)\)\);", text):
        args = [num(a) for a in m.group(3).split(",")]
        out.append({"op": m.group(2), "args": args, "expect": num(m.group(1))})
    return out


def terms_of(body):
    terms = {}
    for m in re.finditer(r"struct constr_t (\w+) = CONSTRAINT_TERM\((.*)\);", body):
        terms[m.group(1)] = val(m.group(2))
    envs = {}
    for m in re.finditer(r"struct env_t (\w+) = \{ \.key = NULL, \.val = &(\w+)", body):
        envs[m.group(1)] = m.group(2)
    return terms, envs


def evals():
    text = open(os.path.join(REF, "test/test_eval.c")).read()
    out = []
    for suite, name, body in tests_of(text):
        if not suite.startswith("Eval") or suite in ("EvalWand", "EvalConfl", "EvalTerm"):
            continue
        try:
            terms, _ = terms_of(body)
        except ValueError:
            continue
        for m in re.finditer(r"X = CONSTRAINT_EXPR\((\w+), &(\w+), (?:&(\w+)|NULL)\);\s*EXPECT_EQ\((.*?), eval_\w+\(&X\)\);", body):
            op, l, r, exp = m.group(1), m.group(2), m.group(3), m.group(4)
            if l not in terms or (r and r not in terms):
                continue
            out.append({"test": "%s.%s" % (suite, name), "op": op, "l": terms[l], "r": terms[r] if r else None,
                        "expect": val(exp)})
    return out


def props():
    text = open(os.path.join(REF, "test/test_propagate.c")).read()
    out = []
    for suite, name, body in tests_of(text):
        if suite in ("PropagateTerm", "PropagateWand", "PropagateConfl", "Propagate"):
            continue
        try:
            terms, envs = terms_of(body)
        except ValueError:
            continue
        for blk in re.split(r"MockProxy = new Mock\(\);", body)[1:]:
            mx = re.search(r"X = CONSTRAINT_EXPR\((\w+), &(\w+), (?:&(\w+)|NULL)\);", blk)
            mr = re.search(r"EXPECT_EQ\(([^,]+), propagate_\w+\(&X, (.*), NULL\)\);", blk)
            if not mx or not mr:
                continue
            op, l, r = mx.group(1), mx.group(2), mx.group(3)
            if l not in terms or (r and r not in terms):
                continue
            binds = []
            for mb in re.finditer(r"bind\(&(\w+), (.*?), NULL\)\)", blk):
                binds.append({"term": envs[mb.group(1)], "val": val(mb.group(2))})
            res = mr.group(1).strip()
            res = -1 if res == "PROP_ERROR" else (0 if res == "PROP_NONE" else int(res))
            out.append({"test": "%s.%s" % (suite, name), "op": op, "l": l, "r": r,
                        "terms": {k: terms[k] for k in ([l] + ([r] if r else []))},
                        "vars": sorted(set(envs.values())), "val": val(mr.group(2)), "result": res, "binds": binds})
    return out


def misc():
    # Luby thresholds (test/test_csolve.c:305-337) and step_val order (test/test_csolve.c:628-657)
    text = open(os.path.join(REF, "test/test_csolve.c")).read()
    luby = [int(x) for x in re.findall(r"EXPECT_EQ\((\d+), _fail_threshold\);", text)]
    return {"luby": luby}


if __name__ == "__main__":
    data = {"source": "jeuneS2/csolve test/*.c value tables", "arith": arith(), "eval": evals(), "propagate": props()}
    data.update(misc())
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_unit_vectors.json")
    with open(dst, "w") as f:
        json.dump(data, f, indent=0, sort_keys=True)
    print("arith %d, eval %d, propagate %d, luby %d -> %s" % (
        len(data["arith"]), len(data["eval"]), len(data["propagate"]), len(data["luby"]), dst))
