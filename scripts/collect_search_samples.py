#!/usr/bin/env python3
"""On a B200: search the BASELINE instances with node sampling on (csolve_solve_options.sample_mod), check every
record against the oracle right away and save the records under gpurun_out/samples/ so that
tests/golden/make_search_samples.py can replay them through the compiled reference in the build container.

    python scripts/collect_search_samples.py [names...]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import search_samples as S  # noqa: E402
import util  # noqa: E402


def main():
    names = sys.argv[1:] or list(S.SAMPLED)
    out_dir = os.path.join(ROOT, "gpurun_out", "samples")
    os.makedirs(out_dir, exist_ok=True)
    summary = {}
    for name in names:
        t0 = time.time()
        m, r, s = S.run_sampled(name)
        t1 = time.time()
        fl = s["flags"]
        nonfailed = int(((fl & S.FAILED) == 0).sum())
        n, nf, bad = S.check_against_oracle(m, s, limit=30000)
        row = dict(nodes=int(r.nodes), cuts=int(r.cuts), solutions=int(r.solutions), best=int(r.best), seen=int(s["seen"]),
                   kept=int(len(fl)), nonfailed=nonfailed, counted=int(((fl & S.COUNTED) != 0).sum()),
                   leaves=int(((fl & S.LEAF) != 0).sum()), oracle_checked=n, oracle_mismatches=len(bad),
                   search_s=round(t1 - t0, 2), check_s=round(time.time() - t1, 2), kernel_ms=round(r.kernel_ms, 3))
        summary[name] = row
        print(name, json.dumps(row), flush=True)
        for b in bad[:3]:
            print("   MISMATCH", b[:2], flush=True)
        S.save_samples(os.path.join(out_dir, "raw_%s.npz" % name), s, nodes=np.int64(r.nodes), cuts=np.int64(r.cuts),
                       solutions=np.int64(r.solutions))
    json.dump(summary, open(os.path.join(out_dir, "summary.json"), "w"), indent=1, sort_keys=True)
    if any(v["oracle_mismatches"] for v in summary.values()):
        raise SystemExit("oracle mismatches")


if __name__ == "__main__":
    main()
