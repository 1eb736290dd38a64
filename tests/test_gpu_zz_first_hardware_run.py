"""Two device changes made after the round's GPU minutes were spent; this file sorts last among the GPU tests.

Both were developed against the CPU emulation of the kernel source (tests/harness/simt_emu.h,
tests/test_emulated_kernels.py, scripts/emu_fuzz.py) and have not met hardware before the driver's run; every kernel
other than the two named here is bit-identical in SASS to the build the rest of this suite was last run on. Hence
xfail(strict=False): XPASS = confirmed on the device, XFAIL = not yet.

1. csolve_solve_options.backjump (conflict_backtrack, src/csolve.c:350-364): the back-jumping instance of k_search
   (csrc/kernels_bj.cu). Off by default everywhere; run in a child process with a time limit.
2. k_search_sat with time slices: a warp that k_rebalance handed one frame searched what lay below it in its own stack
   from earlier slices (ALL-mode counts too high). Found by the emulation, which now gets the tree's counters for every
   slice length; on the device a sliced ALL-mode search of a pure SAT model had no test.
3. k_search_sat and the solution sink: the kernel did not end its slice when the solution buffer was nearly full, so a
   pure-SAT ALL model with more than 2^20 solutions ended in CSOLVE_ERR_CAPACITY through the drop-in (loud, never a
   wrong answer). Same origin, same state."""
import json
import os
import subprocess
import sys

import pytest

import csolve_b200 as cb
import util
from csolve_b200 import instances as I

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first hardware run of the back-jumping kernel instance (see the module docstring)")
def test_backjump_keeps_status_optimum_and_all_counts():
    proc = subprocess.run([sys.executable, os.path.join(HERE, "backjump_check.py")], capture_output=True, text=True, timeout=420)
    lines = [json.loads(l) for l in proc.stdout.splitlines() if l.startswith("{")]
    sys.stdout.write(proc.stdout[-4000:])
    sys.stderr.write(proc.stderr[-4000:])
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert lines and lines[-1].get("done") is True
    assert lines[-1]["failed"] == []
    assert lines[-1]["backjumps"] > 0          # the searches did drop more than one level somewhere


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first hardware run of the k_search_sat fix (see the module docstring)")
def test_sliced_all_mode_search_of_a_sat_model_counts_the_tree():
    """random 3-SAT n=70, m=224, ALL: 855 772 solutions, 7 929 685 nodes, 46 108 cuts (oracle, 6 s) -- about ten 1 ms slices
    of k_search_sat on one B200, with k_rebalance between them"""
    m = cb.Model(I.random_3sat(70, 3.2, 2, "ALL"))
    o, _ = util.Oracle(m).solve_tree(0)
    want = (o.solutions, o.calls, o.cuts)
    assert want == (855772, 7929685, 46108)
    p = cb.GpuProblem(m)
    whole = p.solve()
    assert (whole.solutions, whole.nodes, whole.cuts) == want
    for kw in ({"slice_ms": 1}, {"slice_ms": 1, "split_target": 1}, {"slice_ms": 2, "split_target": 64}, {"time_limit_ms": 60000}):
        r = p.solve(**kw)
        assert r.timed_out == 0 and (r.solutions, r.nodes, r.cuts) == want, (kw, r)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="first hardware run of the k_search_sat sink fix (see the module docstring)")
def test_sat_model_streams_two_million_solutions_through_the_sink():
    """random 3-SAT n=70, m=217, ALL: 1 963 113 solutions (more than the 2^20 the buffer holds), 17 006 882 nodes, 74 177
    cuts (oracle, 17 s): every solution once, each satisfying the CNF"""
    import ctypes as C
    import numpy as np
    cnf = I.random_3sat_cnf(70, 3.1, 2)
    m = cb.Model(I.cnf_to_csolve(70, cnf, "ALL"))
    o, _ = util.Oracle(m).solve_tree(0)
    assert (o.solutions, o.calls, o.cuts) == (1963113, 17006882, 74177)
    p = cb.GpuProblem(m)
    got = []
    SINK = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32)

    def sink(user, values, n, stride):
        got.append(np.ctypeslib.as_array(values, shape=(n * stride,)).reshape(n, stride)[:, :70].astype(np.uint8))
    cbk = SINK(sink)
    lib = cb.library()
    lib.csolve_gpu_set_solution_sink.argtypes = [C.c_void_p, SINK, C.c_void_p]
    assert lib.csolve_gpu_set_solution_sink(p._h, cbk, None) == 0
    try:
        r = p.solve()
    finally:
        assert lib.csolve_gpu_set_solution_sink(p._h, C.cast(None, SINK), None) == 0
    assert (r.solutions, r.nodes, r.cuts) == (o.solutions, o.calls, o.cuts)
    sols = np.concatenate(got)
    assert sols.shape[0] == o.solutions and len(got) >= 2
    assert sols.max() <= 1
    packed = np.ascontiguousarray(np.packbits(sols, axis=1))                       # 9 bytes per assignment
    assert np.unique(packed.view(np.dtype((np.void, packed.shape[1])))).shape[0] == o.solutions
    col = {name: i for i, name in enumerate(m.var_names)}
    ok = np.ones(sols.shape[0], bool)
    for cl in cnf:
        sat = np.zeros(sols.shape[0], bool)
        for l in cl:
            sat |= sols[:, col["x%d" % abs(l)]] == (1 if l > 0 else 0)
        ok &= sat
    assert ok.all()
