#!/usr/bin/env python3
"""Replay search-sampled nodes through the COMPILED REFERENCE and commit its answers.

Input : gpurun_out/samples/raw_<name>.npz -- nodes recorded by the search kernels on a B200
        (scripts/collect_search_samples.py; parent domains, variable, value, incumbent, the kernel's own result).
Output: tests/golden/search_<name>.npz -- the same nodes with failed / child recomputed by the unmodified reference
        engine: bind + objective_update_val + propagate_clauses exactly as solve() performs them
        (oracle/ref/replay.c: ref_replay, src/csolve.c:448-457) and, for records flagged as accepted leaves,
        is_true(eval(root)) (ref_eval_root, src/csolve.c:226). The kernel's own answers are NOT copied into the
        fixture; the script reports how many differ (expected: none).

Run in the build container only (needs oracle/_ref, built from /root/reference by oracle/ref/Makefile):
    python tests/golden/make_search_samples.py [names...]
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import search_samples as S  # noqa: E402
import util  # noqa: E402

I32P = C.POINTER(C.c_int32)
MAX_NONFAILED, MAX_FAILED = 12000, 6000


def main():
    ref = util.reference_lib()
    assert ref is not None, "build oracle/_ref first (make -C oracle/ref)"
    names = sys.argv[1:] or list(S.SAMPLED)
    for name in names:
        raw = os.path.join(ROOT, "gpurun_out", "samples", "raw_%s.npz" % name)
        if not os.path.exists(raw):
            print(name, "no raw samples")
            continue
        s = S.load_samples(raw)
        text = S.SAMPLED[name]["text"]()
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(text)
            path = f.name
        V = ref.ref_load(path.encode(), 0, 1)
        os.unlink(path)
        assert V * 2 == s["parent"].shape[1], (name, V)
        ov = ref.ref_obj_var()
        # bounded, deterministic subset: the first MAX_NONFAILED records the kernel did not fail, the first MAX_FAILED it did
        kf = (s["flags"] & S.FAILED) != 0
        keep = np.concatenate([np.flatnonzero(~kf)[:MAX_NONFAILED], np.flatnonzero(kf)[:MAX_FAILED]])
        keep.sort()
        out = dict(flags=[], var=[], val=[], best=[], parent=[], child=[])
        diff = 0
        for i in keep:
            dom = np.ascontiguousarray(s["parent"][i], np.int32)
            o = np.zeros(2 * V, np.int32)
            f = ref.ref_replay(dom.ctypes.data_as(I32P), int(s["var"][i]), int(s["val"][i]), int(s["best"][i]), o.ctypes.data_as(I32P), None)
            if not f and ov >= 0 and o[2 * ov] > o[2 * ov + 1]:
                f = 1          # the reference leaves an empty <obj> undetected at this node; nothing below it is accepted
            flags = int(s["flags"][i]) & S.COUNTED
            if f:
                flags |= S.FAILED
                o[:] = 0
            elif np.array_equal(o[0::2], o[1::2]) and (s["flags"][i] & S.LEAF):
                if ref.ref_eval_root(o.ctypes.data_as(I32P)):
                    flags |= S.LEAF
            same = flags == int(s["flags"][i]) and (f or np.array_equal(o, s["child"][i]))
            diff += 0 if same else 1
            out["flags"].append(flags); out["var"].append(int(s["var"][i])); out["val"].append(int(s["val"][i]))
            out["best"].append(int(s["best"][i])); out["parent"].append(dom); out["child"].append(o)
        res = {k: np.array(v, np.int32) for k, v in out.items()}
        extra = {k: s[k] for k in ("nodes", "cuts", "solutions") if k in s}
        S.save_samples(S.fixture_path(name), res, **extra)
        nf = int(((res["flags"] & S.FAILED) == 0).sum())
        print("%s: %d records (%d not failed, %d counted in bulk, %d leaves); kernel vs reference differences: %d; %.1f KB"
              % (name, len(keep), nf, int(((res["flags"] & S.COUNTED) != 0).sum()), int(((res["flags"] & S.LEAF) != 0).sum()),
                 diff, os.path.getsize(S.fixture_path(name)) / 1024), flush=True)


if __name__ == "__main__":
    main()
