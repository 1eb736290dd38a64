"""The drop-in end to end (needs oracle/_ref, i.e. a checkout where /root/reference was available at build
time): the reference's own main.c + option handling + root phase, with solve() provided by
integration/csolve_gpu_shim.c on top of libcsolve_b200.so, against the reference CLI."""
import os
import re
import subprocess
import tempfile

import pytest

import csolve_b200 as cb
import util
from csolve_b200 import instances as I

pytestmark = pytest.mark.gpu

GPU_CLI = os.path.join(util.ROOT, "oracle", "_ref", "csolve_gpu")


def _run(binary, text, *flags):
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
    try:
        return subprocess.run([binary, "-s", "0", *flags, f.name], capture_output=True, text=True, timeout=300)
    finally:
        os.unlink(f.name)


@pytest.mark.skipif(not (os.path.exists(GPU_CLI) and os.path.exists(util.REF_CLI)), reason="oracle/_ref not built")
def test_dropin_prints_the_same_solutions_as_the_reference():
    for text in (I.queens(6), I.queens(8), I.sudoku(I.SUDOKU_EXAMPLE)):
        a = _run(GPU_CLI, text)
        b = _run(util.REF_CLI, text, "-c", "false")
        assert a.returncode == 0, a.stderr
        sa = sorted(re.findall(r"SOLUTION: (.*), BEST", a.stdout))
        sb = sorted(re.findall(r"SOLUTION: (.*), BEST", b.stdout))
        assert sa == sb and len(sa) > 0


@pytest.mark.skipif(not (os.path.exists(GPU_CLI) and os.path.exists(util.REF_CLI)), reason="oracle/_ref not built")
def test_dropin_optimisation_and_unsat():
    a = _run(GPU_CLI, I.schedule())
    b = _run(util.REF_CLI, I.schedule())
    last = lambda s: re.findall(r"SOLUTION: (.*BEST: -?\d+)", s)[-1]
    assert last(a.stdout) == last(b.stdout)          # same optimal schedule, BEST: 11
    a = _run(GPU_CLI, I.random_3sat(50, seed=1))
    assert "NO SOLUTION FOUND" in a.stdout
    a = _run(GPU_CLI, "ALL; 0 <= x; x <= 3; x > 5;")
    assert "INFEASIBLE PROBLEM" in a.stdout          # printed by the reference's own front end


def test_front_corpus_inputs_solve_like_the_reference_cli():
    """tests/golden/front_corpus.json (the reference's fuzz seeds mutated with its fuzz dictionary, labelled by the
    compiled reference CLI): every accepted input, searched on the device, gives the reference's result"""
    import json
    corpus = json.load(open(os.path.join(util.GOLDEN, "front_corpus.json")))
    n = 0
    for c in corpus:
        if c["kind"] != "ok":
            continue
        m = cb.Model(c["text"])
        r = cb.GpuProblem(m).solve(time_limit_ms=5000)
        assert r.timed_out == 0
        assert bool(r.has_solution) == (not c["no_solution"]), c["text"]
        if m.objective == cb.OBJ_ALL:
            assert r.solutions == c["solutions"], c["text"]
        elif m.objective in (cb.OBJ_MIN, cb.OBJ_MAX) and r.has_solution:
            assert r.best == c["best"], c["text"]
        n += 1
    assert n >= 300
