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
