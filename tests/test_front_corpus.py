"""Front-end hardening: the reference's fuzz seeds and examples mutated with its fuzz dictionary,
every mutant labelled by the COMPILED REFERENCE CLI (tests/golden/make_front_corpus.py -> front_corpus.json).
The built-in front end (csolve_model_parse: lexer, grammar actions, root propagate / normalize / propagate,
env_generate) must classify every input the same way, report the reference's error text where the engine produces
it (ERROR_MSG_UNBOUNDED_VARIABLE, ERROR_MSG_INVALID_OPERATION, src/csolve.h:513-544), and hand the search a model
whose result is the reference's."""
import json
import os

import pytest

import csolve_b200 as cb
import util

CORPUS = json.load(open(os.path.join(util.GOLDEN, "front_corpus.json")))
KIND = {-2: "syntax", -3: "infeasible", -4: "unbounded", -1: "fatal"}


def _parse(text):
    try:
        return "ok", None, cb.Model(text)
    except cb.CsolveError as e:
        return KIND.get(e.code, "code %d" % e.code), e.message, None


def test_corpus_is_substantial():
    kinds = {}
    for c in CORPUS:
        kinds[c["kind"]] = kinds.get(c["kind"], 0) + 1
    assert len(CORPUS) >= 2000 and kinds["ok"] >= 200 and kinds["syntax"] >= 500 and kinds["unbounded"] >= 100 and kinds["infeasible"] >= 30


def test_every_input_is_classified_like_the_reference():
    for c in CORPUS:
        kind, msg, m = _parse(c["text"])
        assert kind == c["kind"], (c["text"], c, kind, msg)
        if kind in ("unbounded", "fatal"):
            assert msg == c["message"], (c["text"], msg)               # the engine's own text
        if kind == "syntax":
            # message format of ERROR_MSG_LEXER_ERROR / ERROR_MSG_PARSER_ERROR and the line number
            assert msg.split(" in line ")[-1] == c["message"].split(" in line ")[-1], (c["text"], msg, c["message"])
            assert msg.startswith("invalid input `") == c["message"].startswith("invalid input `")


def test_accepted_inputs_search_to_the_reference_result():
    """what the front end hands over is what the reference's parser hands to solve(): same result from the oracle's
    reference-mode search (ALL: count; MIN / MAX: optimum; ANY: a solution exists or NO SOLUTION FOUND)"""
    n = 0
    for c in CORPUS:
        if c["kind"] != "ok":
            continue
        m = cb.Model(c["text"])
        o, _ = util.Oracle(m).solve_reference(max_calls=400000)
        if o.hit_limit:
            continue
        assert bool(o.has_solution) == (not c["no_solution"]), c["text"]
        if m.objective == cb.OBJ_ALL:
            assert o.solutions == c["solutions"], c["text"]
        elif m.objective in (cb.OBJ_MIN, cb.OBJ_MAX) and o.has_solution:
            assert o.best == c["best"], c["text"]
        n += 1
    assert n >= 150
