"""Reference CLI (oracle/_ref/csolve_ref) on the config-5 instances, seeds 1..11, on THIS host: default flags
(-c true -f true -r 100) and -c false, time-boxed. Prints a markdown table (the GPU columns come from sat_probe.py)."""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from csolve_b200 import instances as I
REF = os.path.join(ROOT, "oracle", "_ref", "csolve_ref")
limit = int(sys.argv[1]) if len(sys.argv) > 1 else 120
print("| seed | default flags: s, CALLS, status | -c false: s, CALLS, status |")
print("|---|---|---|")
for seed in range(1, 12):
    cells = []
    for flags in ((), ("-c", "false")):
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(I.random_3sat(200, seed=seed))
        t0 = time.perf_counter()
        out = subprocess.run([REF, "-s", "0", "-t", str(limit), *flags, f.name], capture_output=True, text=True).stdout
        dt = time.perf_counter() - t0
        os.unlink(f.name)
        m = re.search(r"CALLS: (\d+)", out)
        st = "TIMEOUT" if "TIMEOUT" in out else ("UNSAT" if "NO SOLUTION FOUND" in out else "SAT")
        cells.append("%.2f s, %s, %s" % (dt, m.group(1) if m else "?", st))
    print("| %d | %s | %s |" % (seed, cells[0], cells[1]), flush=True)
