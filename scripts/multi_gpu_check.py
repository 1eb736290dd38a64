"""Run under torchrun on N GPUs: every objective kind through csolve_b200.distributed (partition, per-slice
incumbent / first-solution exchange over NCCL, final reduction). Prints one line per instance on rank 0."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csolve_b200 as cb
from csolve_b200 import distributed as D
from csolve_b200 import instances as I

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
cases = [("queens10 ALL", I.queens(10), dict(split_target=300), ("solutions", 724)),
         ("queens15 ALL", I.queens(15), {}, ("solutions", 2279184)),
         ("sudoku ALL", I.sudoku(I.SUDOKU_EXAMPLE), dict(order="smallest-domain"), ("solutions", 1)),
         ("schedule MIN", I.schedule(), {}, ("best", 11)),
         ("wcet MAX", I.wcet(), dict(slice_ms=2), ("best", 1560)),
         ("sat200 s1 ANY", I.random_3sat(200, seed=1), dict(prefer_failing=True), ("has_solution", 0)),
         ("sat200 s2 ANY", I.random_3sat(200, seed=2), dict(prefer_failing=True), ("has_solution", 1))]
# the same again with frontier rebalancing between the ranks (frames shipped from busy ranks to ranks that ran dry);
# the lopsided cases give rank 0 almost nothing to begin with (split_target=1: one root frame, owned by one rank)
cases += [(n + " +rebal", t, dict(kw, rebalance=True), w) for n, t, kw, w in cases[:]]
cases += [("queens13 lopsided +rebal", I.queens(13), dict(split_target=1, slice_ms=1, rebalance=True), ("solutions", 73712)),
          ("queens13 lopsided", I.queens(13), dict(split_target=1, slice_ms=1, exchange=True), ("solutions", 73712)),
          ("sat200 s1 lopsided +rebal", I.random_3sat(200, seed=1), dict(prefer_failing=True, split_target=1, slice_ms=2, rebalance=True), ("has_solution", 0)),
          ("sat200 s1 lopsided", I.random_3sat(200, seed=1), dict(prefer_failing=True, split_target=1, slice_ms=2), ("has_solution", 0))]
ok = True
counts = {}
# ---- the comm path (csolve_gpu_comm): shared frontier + incumbents over peer memory, no Python in the loop ----------
comm = None
try:
    comm = D.make_comm(local)
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print("comm unavailable: %s" % e, flush=True)
    ok = False
if comm is not None:
    for name, text, kw, (key, want) in [c for c in cases if "rebal" not in c[0] and "lopsided" not in c[0]] * 2:
        kw = {k: v for k, v in kw.items() if k not in ("exchange", "rebalance", "slice_ms")}
        m = cb.Model(text)
        p = cb.GpuProblem(m, device=local)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, mine = D.solve_comm(p, comm, m.objective, device=dev, **kw)
        dt = time.perf_counter() - t0
        good = out[key] == want
        ok &= good
        nodes_all = [None] * dist.get_world_size()
        dist.all_gather_object(nodes_all, int(mine.nodes))
        if rank == 0:
            print("comm %-14s world=%d %s=%s (want %s) %s nodes=%d per-rank=%s dev=%.2f ms wall=%.1f ms" % (
                name, dist.get_world_size(), key, out[key], want, "OK" if good else "WRONG", out["nodes"], nodes_all,
                out["kernel_ms"] + out["expand_ms"], dt * 1e3), flush=True)
        if "ALL" in name:
            ref = counts.setdefault("comm " + name, (out["solutions"], out["nodes"], out["cuts"]))
            ok &= ref == (out["solutions"], out["nodes"], out["cuts"])
for name, text, kw, (key, want) in ([] if "--comm-only" in sys.argv else cases):
    m = cb.Model(text)
    p = cb.GpuProblem(m, device=local)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, mine = D.solve_partitioned(p, m.objective, device=dev, **kw)
    dt = time.perf_counter() - t0
    good = out[key] == want
    ok &= good
    if rank == 0:
        print("%-14s world=%d %s=%s (want %s) %s nodes=%d rank0_nodes=%d wall=%.1f ms" % (
            name, dist.get_world_size(), key, out[key], want, "OK" if good else "WRONG", out["nodes"], mine.nodes, dt * 1e3), flush=True)
    if "ALL" in name or "lopsided" in name and "queens" in name:
        # all-solutions counters are traversal-independent: rebalancing must not change them
        ref = counts.setdefault(name.replace(" +rebal", ""), (out["solutions"], out["nodes"], out["cuts"]))
        if ref != (out["solutions"], out["nodes"], out["cuts"]):
            ok = False
            if rank == 0:
                print("   counters differ from the run without rebalancing: %s vs %s" % ((out["solutions"], out["nodes"], out["cuts"]), ref))
dist.barrier()
if comm is not None:
    comm.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
