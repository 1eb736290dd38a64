"""Bounded runs of the two CPU fuzzers (scripts/diff_fuzz.py, scripts/fuzz_front.py; longer runs by hand):
 * the product's contractors as the kernels run them per lane (compiled watch records: NOT(EQ) forms, literal clauses,
   linear clauses with the lanes emulated, small linear relations, memoised interpreter) against the oracle on random walks over generated
   models -- fail flags and post-fixpoint domains must be equal; the one documented deviation (an <obj> interval emptied
   by the incumbent fails the node at once) is skipped;
 * mutated inputs through the built-in front end: an error code or a model, never a crash."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_contractors_equal_oracle_on_generated_models():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "diff_fuzz.py"), "500000", "1500"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    m = re.search(r"models (\d+) \(with linear clause (\d+), with small linear relations (\d+)\), node transitions (\d+), mismatches (\d+)", out.stdout)
    assert m, out.stdout[-2000:]
    models, lin, rel, nodes, mism = map(int, m.groups())
    assert mism == 0, out.stdout[-4000:]
    assert models >= 250 and nodes >= 10000 and lin >= 8 and rel >= 40


def test_front_end_survives_mutated_inputs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_front.py"), "77", "1500"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, "front end crashed: rc=%d %s" % (out.returncode, out.stderr[-2000:])
    m = re.search(r"parsed (\d+) rejected (\d+)", out.stdout)
    assert m and int(m.group(1)) + int(m.group(2)) == 1500 and int(m.group(1)) > 0
