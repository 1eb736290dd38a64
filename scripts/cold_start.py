"""Cold-process wall time of the drop-in binary (oracle/_ref/csolve_gpu = the reference's main.c + front end with
solve() from integration/csolve_gpu_shim.c): one solve per process, the way the reference is used. CUDA context
creation, module load and the first allocations are inside the measurement. Prints a markdown table.
    python scripts/cold_start.py [n_runs]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from csolve_b200 import instances as I

GPU = os.path.join(ROOT, "oracle", "_ref", "csolve_gpu")
REF = os.path.join(ROOT, "oracle", "_ref", "csolve_ref")
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def timed(binary, text, env=None, flags=()):
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write(text)
    best = None
    try:
        for _ in range(runs):
            t0 = time.perf_counter()
            with open(os.devnull, "w") as null:
                subprocess.run([binary, "-s", "0", *flags, f.name], stdout=null, stderr=subprocess.STDOUT, env=env, check=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    finally:
        os.unlink(f.name)
    return best


count_only = dict(os.environ, CSOLVE_GPU_COUNT_ONLY="1")
rows = []
for n in (8, 12, 14, 16):
    text = I.queens(n)
    g_count = timed(GPU, text, env=count_only)
    g_print = timed(GPU, text) if n <= 14 else None
    ref = timed(REF, text) if n <= 12 else None
    rows.append("| queens%d ALL | %.3f | %s | %s |" % (n, g_count, "%.3f" % g_print if g_print else "-", "%.3f" % ref if ref else "-"))
for name, text, flags in (("wcet MAX", I.wcet(), ("-f", "false")), ("3-SAT n=200 seed 1 ANY", I.random_3sat(200, seed=1), ("-c", "false"))):
    rows.append("| %s | %.3f | - | - |" % (name, timed(GPU, text, flags=flags)))
print("| instance | csolve_gpu, count only (s) | csolve_gpu, every solution printed to /dev/null (s) | csolve_ref (s) |")
print("|---|---|---|---|")
print("\n".join(rows))
print("\nbest of %d cold processes each; the reference CPU needs 6 629.7 s for queens16 (BASELINE.md)" % runs)
