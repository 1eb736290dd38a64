"""Nodes sampled out of the REAL search kernels at BASELINE sizes (tests/search_samples.py): the kernels' own records
of (parent domains, decision, incumbent) -> (fail flag, post-fixpoint domains) against
  * the oracle, record by record, and
  * tests/golden/search_<name>.npz -- the answers of the compiled reference (ref_replay) for records collected on a
    B200: deterministic trees must reproduce the fixture's record set exactly; for every instance the fixture's
    parents are also pushed through csolve_gpu_propagate_batch."""
import os

import numpy as np
import pytest

import csolve_b200 as cb
import search_samples as S
import util

pytestmark = pytest.mark.gpu

FIXTURES = [n for n in S.SAMPLED if os.path.exists(S.fixture_path(n))]


@pytest.mark.parametrize("name", list(S.SAMPLED))
def test_sampled_search_nodes_match_the_oracle(name):
    m, r, s = S.run_sampled(name)
    assert s["seen"] == len(s["var"]) > 0              # nothing dropped: the buffer was large enough
    n, nonfailed, bad = S.check_against_oracle(m, s, limit=12000)
    assert not bad, bad[:3]
    assert n > 0 and nonfailed > 0


def test_the_fixtures_are_there():
    assert len(FIXTURES) >= 8, FIXTURES


@pytest.mark.parametrize("name", [n for n in FIXTURES if S.SAMPLED[n]["determ"]])
def test_deterministic_trees_reproduce_the_reference_fixture(name):
    """same model, same sampling rule -> the same set of nodes; the fixture holds the reference's answers for them"""
    fx = S.load_samples(S.fixture_path(name))
    m, r, s = S.run_sampled(name)
    if "nodes" in fx:
        assert (int(r.nodes), int(r.cuts), int(r.solutions)) == (int(fx["nodes"]), int(fx["cuts"]), int(fx["solutions"]))
    have = {row.tobytes() for row in S.canonical(s)}
    want = S.canonical(fx)
    missing = [row for row in want if row.tobytes() not in have]
    assert not missing, (len(missing), len(want))


@pytest.mark.parametrize("name", FIXTURES)
def test_fixture_parents_through_the_parity_hook(name):
    fx = S.load_samples(S.fixture_path(name))
    cfg = S.SAMPLED[name]
    old = {k: os.environ.get(k) for k in cfg.get("env", {})}
    os.environ.update(cfg.get("env", {}))
    try:
        p = cb.GpuProblem(cb.Model(cfg["text"]()))
        out, failed = p.propagate_batch(fx["parent"], fx["var"], fx["val"], fx["best"])
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    ref_failed = (fx["flags"] & S.FAILED) != 0
    assert np.array_equal(failed.astype(bool), ref_failed)
    assert np.array_equal(out[~ref_failed], fx["child"][~ref_failed])
