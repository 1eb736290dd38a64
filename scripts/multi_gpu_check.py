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
ok = True
for name, text, kw, (key, want) in cases:
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
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
