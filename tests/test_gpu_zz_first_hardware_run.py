"""Two device changes made after the round's GPU minutes were spent; this file sorts last among the GPU tests.

Both were developed against the CPU emulation of the kernel source (tests/harness/simt_emu.h,
tests/test_emulated_kernels.py, scripts/emu_fuzz.py) and have not met hardware before the driver's run; every kernel
other than the two named here is bit-identical in SASS to the build the rest of this suite was last run on. Hence
xfail(strict=False): XPASS = confirmed on the device, XFAIL = not yet.

1. csolve_solve_options.backjump (conflict_backtrack, src/csolve.c:350-364): the back-jumping instance of k_search
   (csrc/kernels_bj.cu). Off by default everywhere; run in a child process with a time limit.
2. k_search_sat with time slices: a warp that k_rebalance handed one frame searched what lay below it in its own stack
   from earlier slices (ALL-mode counts too high). Found by the emulation, which now gets the tree's counters for every
   slice length; on the device a sliced ALL-mode search of a pure SAT model had no test."""
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
    proc = subprocess.run([sys.executable, os.path.join(HERE, "backjump_check.py")], capture_output=True, text=True, timeout=900)
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
