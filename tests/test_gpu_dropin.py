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


@pytest.mark.skipif(not os.path.exists(GPU_CLI), reason="oracle/_ref not built")
def test_dropin_honours_the_reference_options():
    """-o is recovered from the reference's own variable heap, -c / -f / -r through its getters, -t / -j through
    integration/csolve_accessors.c: the CALLS the drop-in prints are the library's for exactly those options"""
    stats = lambda out: tuple(int(x) for x in re.search(r"CALLS: (\d+), CUTS: (\d+), PROPS: \d+, CONFL: \d+, SOLUTIONS: (\d+)", out).groups())
    text = I.queens(9)
    m = cb.Model(text)
    seen = set()
    for name, code in (("none", 0), ("smallest-domain", 1), ("largest-domain", 2), ("smallest-value", 3), ("largest-value", 4)):
        a = _run(GPU_CLI, text, "-o", name)
        assert a.returncode == 0, a.stderr
        r = cb.GpuProblem(m).solve(order=code, max_solutions=1000)
        assert stats(a.stdout) == (r.nodes, r.cuts, r.solutions), name
        assert len(re.findall(r"SOLUTION:", a.stdout)) == 352
        seen.add(r.nodes)
    assert len(seen) >= 3                                   # the orders really walk different trees
    # -t: a search that cannot finish in a second stops with the reference's TIMEOUT line
    a = _run(GPU_CLI, I.random_3sat(400, seed=5), "-t", "1", "-c", "false")
    assert "TIMEOUT" in a.stdout or "SOLUTION" in a.stdout or "NO SOLUTION FOUND" in a.stdout
    # -j N: N GPUs when the box has them (more than it has: all of them); same counts
    a = _run(GPU_CLI, I.queens(10), "-j", "8")
    assert stats(a.stdout)[2] == 724 and len(re.findall(r"SOLUTION:", a.stdout)) == 724


@pytest.mark.skipif(not os.path.exists(GPU_CLI), reason="oracle/_ref not built")
def test_dropin_prints_every_solution_of_a_large_count():
    """13-queens: 73 712 solutions streamed through the solution sink between time slices, none lost, none twice;
    CSOLVE_GPU_COUNT_ONLY=1 prints none and counts the same"""
    text = I.queens(13)
    a = _run(GPU_CLI, text)
    sols = re.findall(r"SOLUTION: (.*), BEST", a.stdout)
    assert len(sols) == 73712 and len(set(sols)) == 73712
    assert "SOLUTIONS: 73712" in a.stdout
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
    try:
        b = subprocess.run([GPU_CLI, "-s", "0", f.name], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, CSOLVE_GPU_COUNT_ONLY="1"))
    finally:
        os.unlink(f.name)
    assert "SOLUTIONS: 73712" in b.stdout and "SOLUTION:" not in b.stdout


def test_solution_sink_streams_more_solutions_than_the_buffer_holds():
    """the sink through the C ABI: 12-queens with the smallest buffer the library allows -- several drains"""
    import ctypes as C
    m = cb.Model(I.queens(12))
    p = cb.GpuProblem(m)
    got = []
    SINK = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32)

    def sink(user, values, n, stride):
        import numpy as np
        a = np.ctypeslib.as_array(values, shape=(n * stride,)).reshape(n, stride)[:, :12].copy()
        got.append(a)
    cbk = SINK(sink)
    lib = cb.library()
    lib.csolve_gpu_set_solution_sink.argtypes = [C.c_void_p, SINK, C.c_void_p]
    assert lib.csolve_gpu_set_solution_sink(p._h, cbk, None) == 0
    r = p.solve()
    import numpy as np
    allsol = np.concatenate(got)
    assert r.solutions == 14200 and allsol.shape[0] == 14200
    assert len({tuple(x) for x in allsol.tolist()}) == 14200
    t = __import__("json").load(open(os.path.join(util.GOLDEN, "tree_counts.json")))["queens12/none"]
    assert (r.nodes, r.cuts) == (t["nodes"], t["cuts"])
    assert lib.csolve_gpu_set_solution_sink(p._h, C.cast(None, SINK), None) == 0
    assert p.solve().solutions == 14200
