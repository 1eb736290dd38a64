"""csolve_solve_options.backjump (conflict_backtrack, src/csolve.c:350-364) on the device.

This file sorts last among the GPU tests and runs its searches in a child process with a time limit. The back-jumping
instance of k_search (csrc/kernels_bj.cu) was written after the round's GPU minutes were spent: every other kernel of
the library is bit-identical in SASS to the build the rest of this suite was run on, this one has not met hardware
before the driver's run. Hence xfail(strict=False): XPASS = the option works on the device, XFAIL = it does not yet,
and nothing else depends on it (the option is off by default everywhere)."""
import json
import os
import subprocess
import sys

import pytest

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
