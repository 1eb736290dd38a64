"""world_size-2 gloo test of the multi-rank host logic (csolve_b200/distributed.py): partition
arguments handed to each rank and the final reductions. The per-rank search is replaced by the
oracle run on a value-partition of the first variable, so the whole thing runs on CPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Res:
    pass


def _worker(rank, world, port, objective_name, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import csolve_b200 as cb
    from csolve_b200 import distributed as D
    from csolve_b200 import instances as I
    import util

    seen = {}

    class OracleBackedProblem:
        """stands in for GpuProblem: rank r handles the values v of the first variable with v % world == r"""

        def __init__(self, text_fn):
            self.text_fn = text_fn

        def solve(self, part_rank=0, part_count=1, **kw):
            seen["part"] = (part_rank, part_count)
            tot = _Res()
            tot.solutions = tot.nodes = tot.cuts = tot.props = tot.clause_visits = tot.kernel_launches = 0
            tot.best = 2**31 - 1
            tot.has_solution = tot.timed_out = 0
            tot.kernel_ms = 10.0 * (part_rank + 1)
            tot.expand_ms = 1.0
            for v in range(1, 7):
                if v % part_count != part_rank:
                    continue
                m = cb.Model(self.text_fn(v))
                r, _ = util.Oracle(m).solve_reference() if m.obj_var >= 0 else util.Oracle(m).solve_tree(0)
                tot.solutions += r.solutions; tot.nodes += r.calls; tot.cuts += r.cuts; tot.props += r.props
                if r.has_solution:
                    tot.has_solution = 1
                    tot.best = min(tot.best, r.best)
            return tot

    if objective_name == "ALL":
        prob = OracleBackedProblem(lambda v: I.queens(6) + "X1 = %d;\n" % v)
        out, mine = D.solve_partitioned(prob, cb.OBJ_ALL)
    else:
        # minimise X1 + X2 over 6-queens placements with X1 fixed per part
        prob = OracleBackedProblem(lambda v: I.queens(6, "MIN X1 + X2") + "X1 = %d;\n" % v)
        out, mine = D.solve_partitioned(prob, cb.OBJ_MIN)
    q.put((rank, seen["part"], out, int(mine.solutions)))
    dist.barrier()
    dist.destroy_process_group()


def _run(objective_name):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, objective_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(outs, key=lambda o: o[0])


def test_partitioned_all_solutions_sum():
    outs = _run("ALL")
    assert [o[1] for o in outs] == [(0, 2), (1, 2)]
    a, b = outs[0][2], outs[1][2]
    assert a == b                                  # every rank holds the reduced totals
    assert a["solutions"] == 4                     # 6-queens has 4 solutions
    assert outs[0][3] + outs[1][3] == 4 and outs[0][3] != 4
    assert a["kernel_ms"] == 20.0 and a["kernel_ms_min"] == 10.0   # MAX over ranks


def test_partitioned_incumbent_min():
    outs = _run("MIN")
    a = outs[0][2]
    assert a == outs[1][2]
    assert a["has_solution"] == 1
    # 6-queens solutions: (2,4,6,1,3,5) (3,6,2,5,1,4) (4,1,5,2,6,3) (5,3,1,6,4,2): min X1+X2 = 5
    assert a["best"] == 5


# ---- frontier rebalancing between ranks (csolve_b200/distributed.py: plan_transfers / make_rebalance) -----------
def test_plan_transfers():
    sys.path.insert(0, ROOT)
    from csolve_b200.distributed import plan_transfers
    assert plan_transfers([10, 12], 64) == [[0, 0], [0, 0]]                 # nobody ran dry
    assert plan_transfers([0, 0], 64) == [[0, 0], [0, 0]]                   # nobody has anything
    g = plan_transfers([0, 1000], 64)                                       # capped by max_frames
    assert g == [[0, 0], [64, 0]]
    g = plan_transfers([0, 40, 0, 8], 1024)                                 # fair = 12: rank 1 spares min(20, 28), rank 3 nothing
    assert g[1] == [10, 0, 10, 0] and g[3] == [0, 0, 0, 0] and g[0] == [0] * 4
    for busy in ([0, 3], [5, 0, 0, 0, 0, 0, 0, 0], [0, 4736, 4736, 0]):
        g = plan_transfers(busy, 1024)
        for d, row in enumerate(g):
            assert sum(row) <= busy[d] // 2 and (busy[d] > 0 or sum(row) == 0)
            assert all(x == 0 for r, x in enumerate(row) if busy[r] > 0)     # only ranks that ran dry receive


def _rebalance_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    import csolve_b200 as cb
    from csolve_b200 import distributed as D

    FW = 12

    class SlicedFake:
        """stands in for GpuProblem: a bag of unit-work frames, 10 of them searched per slice; the per-slice
        protocol (exchange, then rebalance while somebody still works) is the one of csolve_gpu_solve()"""

        def __init__(self, n):
            self.frames = [np.full(FW, 1000 * rank + i, np.int32) for i in range(n)]
            self.done_ids = []
            self.exchange = self.rebalance = None

        def set_exchange(self, fn):
            self.exchange = fn

        def set_rebalance(self, fn):
            self.rebalance = fn

        def export_frames(self, n):
            n = min(n, len(self.frames) // 2)
            out, self.frames = self.frames[:n], self.frames[n:]
            return np.stack(out) if out else np.zeros((0, FW), np.int32)

        def import_frames(self, fr):
            self.frames += [np.array(f) for f in fr]
            return len(fr)

        def solve(self, part_rank=0, part_count=1, **kw):
            slices = 0
            while True:
                for f in self.frames[:10]:
                    self.done_ids.append(int(f[0]))
                self.frames = self.frames[10:]
                slices += 1
                local_done = 0 if self.frames else 1
                _, _, all_done = self.exchange(0, 0, local_done)
                if all_done:
                    break
                self.rebalance(self, 0 if self.frames else 8, len(self.frames), FW)
            r = _Res()
            r.solutions = len(self.done_ids); r.nodes = r.cuts = r.props = r.clause_visits = 0
            r.kernel_launches = slices
            r.best = 0; r.has_solution = 1; r.timed_out = 0; r.kernel_ms = float(slices); r.expand_ms = 0.0
            return r

    prob = SlicedFake(200 if rank == 0 else 0)
    out, mine = D.solve_partitioned(prob, cb.OBJ_ALL, rebalance=True)
    q.put((rank, out, sorted(prob.done_ids)))
    dist.barrier()
    dist.destroy_process_group()


def test_rebalance_moves_frames_to_the_rank_that_ran_dry():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rebalance_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in procs], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids0, ids1 = outs[0][2], outs[1][2]
    assert sorted(ids0 + ids1) == list(range(200))          # every frame searched exactly once
    assert len(ids1) > 40                                   # the idle rank took a real share
    assert outs[0][1]["solutions"] == 200 and outs[0][1] == outs[1][1]
    assert outs[0][1]["kernel_ms"] < 20                     # 200 frames at 10 per slice on one rank would be 20 slices


# ---- solve_comm: the optimum's witness when the incumbent a rank reports is a peer's -------------------------------
def _witness_worker(rank, world, port, lost, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import csolve_b200 as cb
    from csolve_b200 import distributed as D

    class _Model:
        obj_var = 2

    class CommFake:
        """stands in for GpuProblem.solve(comm=...): both ranks end with the incumbent 7 (pushed over peer memory), only
        rank 1 found it and holds the assignment -- unless `lost`, where its ring was overwritten"""
        model = _Model()

        def solve(self, comm=None, **kw):
            r = _Res()
            r.solutions = 3 if rank == 0 else 2
            r.nodes, r.cuts, r.props, r.clause_visits, r.kernel_launches = 100 + rank, 50, 10, 20, 4
            r.best, r.has_solution, r.timed_out = 7, 1, 0
            r.kernel_ms, r.expand_ms = 5.0 + rank, 0.25
            chain = [[1, 2, 12], [2, 1, 9]] if rank == 0 else [[4, 4, 8], [3, 0, 7]]
            r.assignments = chain[:1] if (lost and rank == 1) else chain
            return r

    try:
        out, mine = D.solve_comm(CommFake(), object(), cb.OBJ_MIN)
        q.put((rank, out, None))
    except RuntimeError as e:
        q.put((rank, None, str(e)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lost", [False, True])
def test_solve_comm_checks_the_witness_of_the_optimum(lost):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_witness_worker, args=(r, 2, port, lost, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in procs], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if lost:
        assert all(o[1] is None and "witness" in o[2] for o in outs)      # loud on every rank, never a silent optimum without proof
    else:
        a = outs[0][1]
        assert a == outs[1][1]
        assert (a["best"], a["has_solution"], a["has_witness"]) == (7, 1, 1)
        assert (a["solutions"], a["nodes"], a["kernel_launches"]) == (5, 201, 8)
        assert abs(a["kernel_ms"] - 6.0) < 1e-6 and abs(a["kernel_ms_min"] - 5.0) < 1e-6 and abs(a["expand_ms"] - 0.25) < 1e-6
