"""The search kernels' SOURCE (csolve_b200/csrc/kernels.cu) on the CPU: tests/harness/simt_emu.h runs every CUDA thread as
a fiber with real warp collectives, tests/harness/emu_search.cpp drives the depth-first phase the way capi.cu does (root
frame, donation ring, time slices with k_rebalance between them). Checked against the oracle: ALL counters are the
tree's, optima and statuses are the oracle's, assignments satisfy the model. Test infrastructure only: the product has no
CPU path (tests/test_abi.py), and what runs here is the kernels' logic, not their memory ordering (simt_emu.h)."""
import platform
import random

import numpy as np
import pytest

import csolve_b200 as cb
import util
from csolve_b200 import instances as I

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's fiber switch is x86-64 assembly")


def tree(m, order=0):
    o, _ = util.Oracle(m).solve_tree(order)
    return (o.solutions, o.calls, o.cuts)


def counters(r):
    return (r.solutions, r.nodes, r.cuts)


def satisfies(cnf, names, assignment):
    val = {k: v for k, v in zip(names, assignment) if k.startswith("x")}
    return set(val.values()) <= {0, 1} and all(any((val["x%d" % abs(l)] == 1) == (l > 0) for l in cl) for cl in cnf)


@pytest.mark.parametrize("n", [4, 6, 8, 9])
def test_headline_kernel_counts_the_oracle_tree(n):
    """k_search_lov<false, true> (N-queens, the headline configuration): solutions, nodes and cuts of the whole tree, in
    every variable order, with work shared between 16 warps, with and without time slices"""
    m = cb.Model(I.queens(n))
    for order in range(5):
        want = tree(m, order)
        for blocks, slice_clock in ((1, 0), (2, 0), (2, 20000)):
            r, _ = util.emu_search(m, order=order, n_blocks=blocks, general=False, slice_clock=slice_clock)
            assert counters(r) == want, (n, order, blocks, slice_clock)
            if slice_clock and n >= 8:
                assert r.slices > 1
    assert want[0] == {4: 2, 6: 4, 8: 92, 9: 352}[n]


def test_general_kernel_on_the_same_trees():
    """k_search<false> on N-queens (what batched roots run on) and on a 3-SAT model in ALL mode"""
    for text in (I.queens(6), I.queens(8), I.random_3sat(30, 3.6, 21, "ALL")):
        m = cb.Model(text)
        for order in (0, 1, 4):
            want = tree(m, order)
            for blocks, slice_clock in ((1, 0), (2, 3000)):
                r, _ = util.emu_search(m, order=order, n_blocks=blocks, general=True, slice_clock=slice_clock)
                assert counters(r) == want, (text[:20], order, blocks, slice_clock)


def test_expansion_and_the_partition_between_ranks():
    """k_search<true> / k_search_lov<true> expand the root breadth-first; in ALL mode every rank does, keeps the frames
    whose path hash maps to it, and the ranks' counters add up to the tree (the multi-GPU mode of the headline bench)"""
    for text in (I.queens(8), I.queens(9), I.random_3sat(30, 3.6, 21, "ALL")):
        m = cb.Model(text)
        want = tree(m)
        for general in (False, True):
            for split, world in ((8, 1), (64, 1), (64, 2), (200, 3), (100000, 1)):
                tot = [0, 0, 0]
                for rank in range(world):
                    r, _ = util.emu_search(m, n_blocks=2, general=general, split_target=split, part_rank=rank, part_count=world, slice_clock=20000)
                    tot[0] += r.solutions; tot[1] += r.nodes; tot[2] += r.cuts
                    assert r.expand_levels >= 1 and r.frontier >= min(split, 8)
                assert tuple(tot) == want, (text[:12], general, split, world)


def test_sudoku_kernel():
    """k_search_lovk: one puzzle per search here (the batched root phase is device-side host logic)"""
    rng = random.Random(5)
    for _ in range(3):
        g = I.sudoku_puzzle(rng, 26)
        m = cb.Model(I.sudoku(g, "ALL"))
        want = tree(m)
        for blocks, slice_clock in ((1, 0), (2, 2000)):
            r, _ = util.emu_search(m, n_blocks=blocks, general=False, slice_clock=slice_clock)
            assert counters(r) == want and r.solutions == 1


def test_batched_sudoku_roots():
    """csolve_gpu_solve_batch: the root phase on the device (k_root_frames_lovk / k_root_frames: every root propagated to
    fixpoint, one tagged frame per consistent root), then one search over all roots with per-root solution counters --
    against the oracle's per-instance trees (tests/golden/tree_counts.json, the fixture of the device test)"""
    import json
    import os
    t = json.load(open(os.path.join(util.GOLDEN, "tree_counts.json")))["sudoku_batch200_seed20261018/smallest-domain"]
    grids = I.sudoku_batch(200, seed=20261018)[:24]
    want = t["per_root"][:24]
    m = cb.Model(I.sudoku("." * 81))
    roots = I.sudoku_roots(m.var_names, grids)
    for general in (False, True):
        for blocks, slice_clock in ((1, 0), (2, 3000)):
            r, counts, failed = util.emu_search_batch(m, roots, order=1, n_blocks=blocks, general=general, slice_clock=slice_clock)
            assert not failed.any() and counts.tolist() == [w[0] for w in want]
            assert counters(r) == tuple(sum(w[k] for w in want) for k in range(3))
    r1, c1, _ = util.emu_search_batch(m, roots[5:6], order=1)
    assert counters(r1) == tuple(want[5])
    # a root whose clues contradict each other fails in the root phase and contributes nothing
    bad = np.array(roots[:3], np.int32).copy()
    a, b = m.var_names.index(I._cell(0, 0)), m.var_names.index(I._cell(0, 1))
    bad[1, 2 * a:2 * a + 2] = 5
    bad[1, 2 * b:2 * b + 2] = 5
    r, counts, failed = util.emu_search_batch(m, bad, order=1)
    assert failed.tolist() == [0, 1, 0] and counts.tolist() == [want[0][0], 0, want[2][0]]


def test_bit_state_sat_kernel_with_time_slices():
    """k_search_sat: the tree's counters in ALL mode -- also when the search is cut into time slices and k_rebalance hands
    parked frames to idle warps (this test found that a warp which received one frame went on to search what was left
    below it in its own stack from earlier slices: counts too high; fixed in the kernel)"""
    cnf = I.random_3sat_cnf(30, 3.6, 21)
    m = cb.Model(I.cnf_to_csolve(30, cnf, "ALL"))
    want = tree(m)
    assert want == (1152, 9413, 156)
    for blocks in (1, 2):
        for slice_clock in (0, 30000, 10000, 3000):
            r, _ = util.emu_search(m, n_blocks=blocks, general=False, slice_clock=slice_clock)
            assert counters(r) == want, (blocks, slice_clock, counters(r))
    for n, ratio, seed in ((40, 4.26, 1), (50, 4.6, 4), (60, 4.26, 3)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(I.cnf_to_csolve(n, cnf))
        sat = tree(m)[0] > 0
        for blocks, slice_clock, pf in ((1, 0, False), (2, 5000, False), (2, 0, True)):
            r, sols = util.emu_search(m, n_blocks=blocks, general=False, slice_clock=slice_clock, prefer_failing=pf)
            assert r.has_solution == (1 if sat else 0)
            assert not sat or satisfies(cnf, m.var_names, sols[0])


def test_every_solution_through_a_bounded_buffer():
    """the solution sink of the drop-in (csolve_gpu_set_solution_sink): the kernels end a slice when the solution buffer
    is nearly full, the host drains it -- every solution arrives exactly once, on all four search kernels (the bit-state
    SAT kernel did not look at the buffer at all: a loud overflow error on the device; found here, fixed in the kernel)"""
    for text in (I.queens(8), I.queens(9), I.random_3sat(30, 3.6, 21, "ALL"), I.sudoku(I.sudoku_puzzle(random.Random(7), 24), "ALL")):
        m = cb.Model(text)
        orc = util.Oracle(m)
        o, _ = orc.solve_tree(0)
        for general in (False, True):
            for blocks, slice_clock, headroom in ((1, 0, 40), (2, 5000, 80), (3, 20000, 120)):      # (tight: the schedule is fine-grained)
                r, sols = util.emu_search(m, n_blocks=blocks, general=general, slice_clock=slice_clock, sink_headroom=headroom,
                                          sink_rows=int(o.solutions) + 8)
                uniq = {tuple(s) for s in sols}
                assert len(sols) == len(uniq) == o.solutions == r.solutions, (text[:12], general, blocks, slice_clock)
                for s1 in list(uniq)[:50]:
                    assert orc.leaf_true(np.array([x for v in s1 for x in (v, v)], np.int32))


def test_search_sampled_nodes_replay_through_the_oracle():
    """the parity instrumentation (csolve_solve_options.sample_mod, the SAMPLE instances of the kernels): every recorded
    node -- executed, or counted in bulk by a kernel shortcut -- has the oracle's fail flag and post-fixpoint domains;
    all four search kernels, expansion included. (Here the general kernel's record copy turned out to lack a warp sync
    before the push overwrites the parent domains: harmless on a converged warp, wrong words under the emulator's lane
    order; the sync is in the kernel now.)"""
    import search_samples as S
    cases = (("q8", I.queens(8), {}), ("q9", I.queens(9), dict(order=1)), ("sat30", I.random_3sat(30, 3.6, 21, "ALL"), {}),
             ("sat60", I.random_3sat(60, 4.26, 3), {}), ("schedule", I.schedule(), dict(max_solutions=16)),
             ("sudoku", I.sudoku(I.sudoku_puzzle(random.Random(5), 26)), dict(order=1)))
    counted = 0
    for name, text, kw in cases:
        m = cb.Model(text)
        for general in (False, True):
            for split in (1, 64):
                r, smp = util.emu_sampled_search(m, 3, 2, n_blocks=2, general=general, split_target=split, slice_clock=20000, **kw)
                n, nonfailed, bad = S.check_against_oracle(m, smp)
                assert n > 10 and not bad, (name, general, split, bad[:1])
                assert smp["seen"] == n
                counted += int(((smp["flags"] & S.COUNTED) != 0).sum())
                if m.objective == cb.OBJ_ALL:
                    assert counters(r) == tree(m, kw.get("order", 0))
    assert counted > 1000          # the bulk shortcuts of the specialised kernels were exercised


def test_results_do_not_depend_on_the_lane_schedule():
    """race detection: between two collectives the emulator runs a warp's lanes one after the other; whatever the order
    (ascending, descending, stride 13) counters, optima, recorded nodes and learned-search statuses stay the same -- a lane
    that read what another lane wrote without a warp sync in between would show here"""
    import search_samples as S
    try:
        for step in (31, 13):
            util.emu_set_lane_step(step)
            for text, kw in ((I.queens(8), {}), (I.random_3sat(30, 3.6, 21, "ALL"), {}), (I.sudoku(I.sudoku_puzzle(random.Random(5), 26)), dict(order=1))):
                m = cb.Model(text)
                want = tree(m, kw.get("order", 0))
                for general in (False, True):
                    r, smp = util.emu_sampled_search(m, 3, 2, n_blocks=2, general=general, split_target=16, slice_clock=20000, **kw)
                    assert counters(r) == want and not S.check_against_oracle(m, smp)[2], (step, text[:10], general)
            cnf = I.random_3sat_cnf(50, 4.6, 4)
            m = cb.Model(I.cnf_to_csolve(50, cnf))
            for bj in (False, True):
                r, sols = util.emu_search(m, learn=True, backjump=bj, n_blocks=2, slice_clock=5000)
                assert r.has_solution == 1 and satisfies(cnf, m.var_names, sols[0])
            r, _ = util.emu_search(cb.Model(I.schedule()), n_blocks=2, max_solutions=16, slice_clock=2000)
            assert r.best == 11
    finally:
        util.emu_set_lane_step(1)


def test_several_gpus_on_one_tree():
    """csolve_gpu_comm for ANY / MIN / MAX models, 2 and 3 emulated ranks side by side in one launch: one root frontier on
    rank 0 claimed by everybody (system-scope atomics on front_ctl), incumbents and "found" through the CommBlocks
    (comm_push_best / comm_push_stop / comm_poll), a rank running dry asks its peers from its waiting loop and their busy
    warps serve its donation ring; k_rebalance between the slices. Optimum and witness, status and model are the
    oracle's; an unsatisfiable instance in static order is the same tree whatever the number of ranks."""
    for n, ratio, seed in ((20, 3.2, 12), (24, 3.5, 13), (26, 3.6, 14)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(I.cnf_to_csolve(n, cnf, "MIN " + " + ".join("x%d" % i for i in range(1, n + 1))))
        o, _ = util.Oracle(m).solve_tree(0)
        for world in (1, 2, 3):
            r, w = util.emu_search_comm(m, world, split_target=16, slice_clock=5000)
            assert r.best == o.best and w is not None and satisfies(cnf, m.var_names, w)
            assert sum(v for k, v in zip(m.var_names, w) if k.startswith("x")) == r.best
    for world in (2, 3):
        r, _ = util.emu_search_comm(cb.Model(I.schedule()), world, split_target=16, slice_clock=20000)
        assert r.best == 11
    nodes = set()
    cnf = I.random_3sat_cnf(60, 4.7, 11)
    m = cb.Model(I.cnf_to_csolve(60, cnf))
    assert tree(m)[0] == 0
    for world in (1, 2, 3):
        for general in (False, True):
            r, w = util.emu_search_comm(m, world, split_target=32, slice_clock=10000, general=general)
            assert r.has_solution == 0 and w is None
            nodes.add((r.nodes, r.cuts))
    assert len(nodes) == 1 and nodes.pop() == tree(m)[1:]
    jumps = 0
    for n, ratio, seed in ((50, 4.6, 4), (60, 4.26, 3)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(I.cnf_to_csolve(n, cnf))
        for world in (2, 3):
            for general, pf in ((False, False), (True, True)):
                r, w = util.emu_search_comm(m, world, split_target=32, slice_clock=10000, general=general, prefer_failing=pf)
                assert r.has_solution == 1 and satisfies(cnf, m.var_names, w)
            # -c under -j N: every rank learns into a pool of its own, with and without back-jumping
            for bj in (False, True):
                r, w = util.emu_search_comm(m, world, split_target=8, slice_clock=10000, learn=True, backjump=bj)
                assert r.has_solution == 1 and satisfies(cnf, m.var_names, w) and r.conflicts > 0
                jumps += r.backjumps
    assert jumps > 0


def test_reference_node_transitions_through_the_emulated_kernels():
    """csolve_gpu_propagate_batch (k_propagate_batch / _lov / _lovk): the 25 k node transitions replayed through the COMPILED
    REFERENCE (tests/golden/replay_*.npz, the fixtures of the device test) -- bit-exact fail flags and post-fixpoint
    domains, on the kernel the product picks and on the general one"""
    import glob
    import os
    from make_instances import instance_table
    inst = instance_table()
    n = 0
    for path in sorted(glob.glob(os.path.join(util.GOLDEN, "replay_*.npz"))):
        name = os.path.basename(path)[7:-4]
        if name == "random":
            continue
        z = np.load(path)
        g = {k: z[k] for k in z.files}
        m = cb.Model(inst[name])
        for general in (False, True):
            out, failed = util.emu_propagate_batch(m, g["dom_in"], g["var"], g["val"], g["best"], general=general)
            assert np.array_equal(failed.astype(bool), g["failed"].astype(bool)), (name, general)
            ok = ~g["failed"].astype(bool)
            assert np.array_equal(out[ok], g["dom_out"][ok]), (name, general)
        n += len(g["var"])
    assert n > 20000


def test_frames_shipped_between_ranks_at_the_slice_boundaries():
    """csolve_gpu_set_rebalance + csolve_gpu_export_frames / _import_frames: ranks search their path-hash shares slice by
    slice, a rank that ran dry gets frames k_export_frames split off the busiest rank's parked stacks, k_import_frames
    puts them into its ring as served tickets -- nothing lost, nothing searched twice: the counters are the tree's"""
    moved = 0
    for text in (I.queens(8), I.queens(9), I.random_3sat(30, 3.6, 21, "ALL")):
        m = cb.Model(text)
        want = tree(m)
        for general in (False, True):
            for world, split, slice_clock in ((2, 8, 3000), (3, 64, 5000), (4, 16, 2000)):
                r, n = util.emu_search_exchange(m, world, split_target=split, slice_clock=slice_clock, general=general)
                assert counters(r) == want, (text[:12], general, world)
                moved += n
    assert moved > 50


def test_luby_restarts():
    """-r on ANY models (src/csolve.c:76-83, 264-276): the warps report their failed nodes, the slice ends at the Luby
    threshold, the host drops every frame and expands the root again in the order of the priorities learned so far
    (the harness mirrors capi.cu's loop) -- same status, valid models, on the bit-state kernel and the general one"""
    restarts = 0
    for n, ratio, seed in ((40, 4.26, 1), (50, 4.6, 4), (60, 4.26, 3), (80, 4.26, 5)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(I.cnf_to_csolve(n, cnf))
        sat = tree(m)[0] > 0
        for general in (False, True):
            for blocks, rf, split in ((1, 1, 1), (2, 2, 16), (2, 5, 64)):
                r, sols = util.emu_search(m, n_blocks=blocks, general=general, prefer_failing=True, restart_frequency=rf, split_target=split)
                assert r.has_solution == (1 if sat else 0), (n, general, blocks, rf)
                assert not sat or satisfies(cnf, m.var_names, sols[0])
                restarts += r.restarts
    assert restarts > 50


def test_branch_and_bound_models():
    """MIN / MAX: schedule (optimum 11), a generated weighted model; linear clauses contracted by the whole warp"""
    r, sols = util.emu_search(cb.Model(I.schedule()), n_blocks=1, max_solutions=16)
    assert r.best == 11 and r.has_solution
    r, _ = util.emu_search(cb.Model(I.schedule()), n_blocks=2, max_solutions=16, slice_clock=2000)
    assert r.best == 11
    for n, ratio, seed in ((16, 3.0, 11), (20, 3.2, 12), (24, 3.5, 13)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(I.cnf_to_csolve(n, cnf, "MIN " + " + ".join("x%d" % i for i in range(1, n + 1))))
        o, _ = util.Oracle(m).solve_tree(0)
        for blocks, slice_clock in ((1, 0), (2, 3000)):
            r, sols = util.emu_search(m, n_blocks=blocks, max_solutions=16, slice_clock=slice_clock)
            assert r.best == o.best
            w = dict(zip(m.var_names, sols[0]))
            assert sum(v for k, v in w.items() if k.startswith("x")) == r.best and satisfies(cnf, m.var_names, sols[0])


def test_conflict_learning_and_backjump():
    """-c on the device with and without back-jumping (csolve_solve_options.backjump, the kernels_bj.cu instance): same
    status, valid models, same optima; the search does jump; ALL models never do and keep their solution count"""
    jumps = 0
    for n, ratio, seed in ((20, 4.26, 1), (30, 4.26, 2), (40, 4.26, 1), (40, 4.26, 2), (50, 4.6, 4), (60, 4.26, 3), (70, 4.26, 9)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(I.cnf_to_csolve(n, cnf))
        sat = tree(m)[0] > 0
        for bj in (False, True):
            for blocks, slice_clock, pf in ((1, 0, False), (2, 0, True), (3, 5000, False)):
                r, sols = util.emu_search(m, learn=True, backjump=bj, prefer_failing=pf, n_blocks=blocks, slice_clock=slice_clock)
                assert r.has_solution == (1 if sat else 0), (n, seed, bj, blocks)
                assert not sat or satisfies(cnf, m.var_names, sols[0])
                assert r.conflicts > 0 and (bj or r.backjumps == 0)
                jumps += r.backjumps
    assert jumps > 100
    for n, ratio, seed in ((16, 3.0, 11), (20, 3.2, 12), (24, 3.5, 13)):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        m = cb.Model(I.cnf_to_csolve(n, cnf, "MIN " + " + ".join("x%d" % i for i in range(1, n + 1))))
        o, _ = util.Oracle(m).solve_tree(0)
        for blocks, slice_clock in ((1, 0), (2, 2000)):
            r, sols = util.emu_search(m, learn=True, backjump=True, n_blocks=blocks, max_solutions=16, slice_clock=slice_clock)
            assert r.best == o.best and satisfies(cnf, m.var_names, sols[0])
    m = cb.Model(I.random_3sat(30, 3.6, 21, "ALL"))
    r, _ = util.emu_search(m, learn=True, backjump=True, n_blocks=2)
    assert r.solutions == 1152 and r.backjumps == 0 and r.conflicts > 0


def test_learned_nogoods_are_implied_by_the_model():
    """conflict_create on the device (learn_nogood, src/conflict.c:327-362): every nogood the emulated search learns -- with
    or without back-jumping, 0/1 literals only (src/conflict.c:173-179) -- leaves the model without a solution when its
    literals are added as constraints"""
    checked = 0
    for n, ratio, seed, obj in ((20, 4.0, 3, "ANY"), (24, 3.8, 5, "ALL"), (30, 4.26, 2, "ANY"), (22, 3.5, 8, "MIN")):
        cnf = I.random_3sat_cnf(n, ratio, seed)
        head = obj if obj != "MIN" else "MIN " + " + ".join("x%d" % i for i in range(1, n + 1))
        text = I.cnf_to_csolve(n, cnf, head)
        m = cb.Model(text)
        names = m.var_names
        for bj in (False, True):
            r, _ = util.emu_search(m, learn=True, backjump=bj, n_blocks=2, max_solutions=16)
            ngs = util.emu_nogoods(bj)
            assert len(ngs) == r.conflicts > 0
            for ng in ngs[:: max(1, len(ngs) // 12)]:
                assert all(val in (0, 1) for _, val in ng) and len({v for v, _ in ng}) == len(ng)
                assert all(names[v].startswith("x") for v, _ in ng)
                extra = "".join("%s = %d;\n" % (names[v], val) for v, val in ng)
                try:
                    o, _ = util.Oracle(cb.Model(I.cnf_to_csolve(n, cnf, "ALL") + extra)).solve_tree(0)
                    assert o.solutions == 0, (n, seed, bj, ng)
                except cb.CsolveError as e:
                    assert e.code == -3                      # already infeasible at root
                checked += 1
    assert checked >= 40


def test_generated_models_through_the_emulated_kernels():
    """the differential fuzzer of scripts/emu_fuzz.py, a bounded run: SAT-shaped, N-queens / sudoku and generated models
    (every operator of the grammar) through the kernel the product picks or the general one, random warps / slices"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("emu_fuzz", os.path.join(util.ROOT, "scripts", "emu_fuzz.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    n = 0
    for seed in range(150):
        rng = random.Random(seed)
        fn = (fz.sat_case, fz.generic_case, fz.queens_case)[seed % 3]
        ok, what, r = fn(rng, seed % 2 == 0)
        assert ok, (seed, what)
        n += r is not None
    assert n >= 100
