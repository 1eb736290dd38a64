"""One process per GPU: the search tree is partitioned across ranks, results are reduced.

The reference's only parallel strategy is search-space splitting (worker_spawn, src/csolve.c:105-152)
with a shared incumbent / solution counter (struct shared_t, src/csolve.h:259-266). Here every rank
expands the root frontier identically, keeps the sub-trees whose path hash maps to it and searches them
with no data-path collective; one reduction at the end combines
    solutions, nodes, cuts, props, clause visits   -> SUM
    incumbent (MIN / MAX objective)                -> MIN / MAX
    time                                           -> MAX over ranks
torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist

from .host import OBJ_ANY, OBJ_MAX, OBJ_MIN


def reduce_results(res, objective, device=None, group=None, witness=None):
    """res: SolveResult-like of this rank -> dict with the whole-job totals (same on every rank).
    ONE collective: every rank contributes one row of twelve integers (all-gather), the reduction itself is done
    locally -- sums, flags, the incumbent and the times need three different operators, and each extra collective is
    tens of microseconds on a search that takes a few milliseconds.
    witness: None, or whether this rank holds the assignment that attains ITS `res.best`; the result then says
    whether any rank holds the witness of the whole job's optimum (`has_witness`)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(solutions=int(res.solutions), nodes=int(res.nodes), cuts=int(res.cuts), props=int(res.props),
                    clause_visits=int(res.clause_visits), best=int(res.best), has_solution=int(res.has_solution),
                    timed_out=int(res.timed_out), kernel_ms=float(res.kernel_ms), expand_ms=float(res.expand_ms),
                    kernel_launches=int(res.kernel_launches), kernel_ms_min=float(res.kernel_ms),
                    has_witness=int(bool(witness)) if witness is not None else None)
    dev = device if device is not None else torch.device("cpu")
    world = dist.get_world_size(group)
    row = torch.tensor([res.solutions, res.nodes, res.cuts, res.props, res.clause_visits, res.kernel_launches,
                        res.has_solution, res.timed_out, res.best, int(round(res.kernel_ms * 1e6)),
                        int(round(res.expand_ms * 1e6)), 1 if witness else 0], dtype=torch.int64, device=dev)
    rows = torch.empty(world * row.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(rows, row, group=group)
    rows = rows.view(world, row.numel()).cpu().tolist()
    s = [sum(r[k] for r in rows) for k in range(6)]
    has_solution = max(r[6] for r in rows)
    timed_out = max(r[7] for r in rows)
    # a rank without a solution must not win the incumbent reduction
    cands = [r[8] for r in rows if r[6]]
    if objective == OBJ_MIN and cands:
        best = min(cands)
    elif objective == OBJ_MAX and cands:
        best = max(cands)
    else:
        best = int(res.best)
    solutions = s[0]
    if objective == OBJ_ANY:
        solutions = min(solutions, 1)   # found_any(): one solution is reported (src/csolve.c:207-209)
    has_witness = None
    if witness is not None:
        has_witness = int(any(r[11] and r[6] and r[8] == best for r in rows))
    return dict(solutions=solutions, nodes=s[1], cuts=s[2], props=s[3], clause_visits=s[4], kernel_launches=s[5],
                best=best, has_solution=int(has_solution), timed_out=int(timed_out),
                kernel_ms=max(r[9] for r in rows) * 1e-6, expand_ms=max(r[10] for r in rows) * 1e-6,
                kernel_ms_min=min(r[9] for r in rows) * 1e-6, has_witness=has_witness)


def make_comm(device_index, group=None, frontier_bytes=0):
    """One process per GPU: this rank's Comm, connected to the other ranks' (the 64-byte CUDA IPC handles of the
    ranks' segments travel through ONE all-gather; after that the ranks talk over NVLink peer memory only).
    Returns None for a single rank."""
    from .host import Comm, COMM_HANDLE_BYTES
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    comm = Comm(device_index, rank, world, frontier_bytes)
    dev = torch.device("cuda", device_index) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(comm.handle()), dtype=torch.uint8, device=dev)
    allh = [torch.zeros(COMM_HANDLE_BYTES, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    comm.connect([bytes(t.cpu().tolist()) for t in allh])
    dist.barrier(group=group)
    return comm


def solve_comm(problem, comm, objective, device=None, group=None, **solve_kw):
    """Collective search through a Comm (shared root frontier, incumbents over peer memory), then ONE reduction of
    the ranks' counters. -> (whole-job dict, this rank's SolveResult)"""
    res = problem.solve(comm=comm, **solve_kw)
    witness = None
    obj_var = getattr(getattr(problem, "model", None), "obj_var", -1)
    if objective in (OBJ_MIN, OBJ_MAX) and obj_var is not None and obj_var >= 0:
        # the incumbent a rank reports may be a peer's (pushed over NVLink); whoever found it stored the assignment
        witness = bool(res.has_solution and res.assignments and res.assignments[-1][obj_var] == res.best)
    out = reduce_results(res, objective, device=device, group=group, witness=witness)
    if witness is not None and out["has_solution"] and not out["has_witness"]:
        raise RuntimeError("the witness of the optimum was overwritten in a rank's solution ring; raise max_solutions")
    return out, res


def make_exchange(objective, device=None, group=None):
    """The per-slice exchange between ranks: ONE all-reduce of three integers.
    MIN models reduce with MIN over [best, -found, done]; everything else with MAX over [best, found, -done]
    (so that `done` is always an AND over the ranks and `found` an OR)."""
    dev = device if device is not None else torch.device("cpu")
    is_min = objective == OBJ_MIN

    def exchange(best, found, local_done):
        if is_min:
            t = torch.tensor([best, -found, local_done], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            b, f, d = t.tolist()
            return b, -f, d
        t = torch.tensor([best, found, -local_done], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        b, f, d = t.tolist()
        return b, f, -d
    return exchange


def plan_transfers(busy, max_frames):
    """Who ships how many frames to whom. busy[r] = warps of rank r that still own work (0: the rank ran dry).
    Ranks that ran dry receive an even share of what the busy ranks can spare: a donor keeps at least half of its busy
    warps' worth of frames and never exports more than max_frames. Deterministic: every rank computes the same plan.
    Returns give[d][r] (frames from donor d to receiver r)."""
    world = len(busy)
    give = [[0] * world for _ in range(world)]
    receivers = [r for r in range(world) if busy[r] == 0]
    donors = [d for d in range(world) if busy[d] > 0]
    if not receivers or not donors:
        return give
    fair = sum(busy) // world                      # frames per rank if one busy warp is worth one frame
    for d in donors:
        spare = min(max_frames, busy[d] // 2, max(busy[d] - fair, 0))
        share = spare // len(receivers)
        if share == 0:
            continue
        for r in receivers:
            give[d][r] = share
    return give


def make_rebalance(device=None, group=None, max_frames=1024):
    """The per-slice frontier rebalancing between ranks: one all-gather of the busy counts; only when
    some rank ran dry while another still works, one all-gather of the exported frames (padded to max_frames)."""
    dev = device if device is not None else torch.device("cpu")

    def rebalance(problem, n_idle, n_busy, frame_words):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        mine = torch.tensor([n_busy], dtype=torch.int64, device=dev)
        allb = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine, group=group)
        busy = [int(t.item()) for t in allb]
        give = plan_transfers(busy, max_frames)
        if not any(any(row) for row in give):
            return 0
        want = sum(give[rank])
        frames = problem.export_frames(want) if want > 0 else None
        n_out = 0 if frames is None else int(frames.shape[0])
        buf = torch.zeros((max_frames * frame_words + 1,), dtype=torch.int32, device=dev)
        buf[0] = n_out
        if n_out:
            buf[1:1 + n_out * frame_words] = torch.from_numpy(frames.reshape(-1)).to(dev)
        allf = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(allf, buf, group=group)
        got = 0
        for d in range(world):
            if give[d][rank] == 0:
                continue
            n_d = int(allf[d][0].item())
            # donor d dealt its exported frames to its receivers in rank order, give[d][r] each (fewer if it ran short)
            start = 0
            for r in range(rank):
                start += give[d][r]
            cnt = max(0, min(give[d][rank], n_d - start))
            if cnt:
                fr = allf[d][1 + start * frame_words:1 + (start + cnt) * frame_words].cpu().numpy()
                got += problem.import_frames(fr.reshape(cnt, frame_words))
        return got
    return rebalance


def solve_partitioned(problem, objective, device=None, group=None, exchange=None, rebalance=False, **solve_kw):
    """Search this rank's share of the tree and reduce. `problem` is a GpuProblem (or anything with .solve).
    exchange=True installs the per-slice incumbent / first-solution exchange (MIN, MAX and ANY models);
    rebalance=True additionally ships frames from busy ranks to ranks that ran dry (it needs the exchange: that is
    where the ranks agree that everybody is done)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if exchange is None:
        exchange = world > 1 and (rebalance or objective in (OBJ_MIN, OBJ_MAX, OBJ_ANY))
    exchange = exchange or rebalance
    if exchange and world > 1 and hasattr(problem, "set_exchange"):
        problem.set_exchange(make_exchange(objective, device=device, group=group))
    if rebalance and world > 1 and hasattr(problem, "set_rebalance"):
        problem.set_rebalance(make_rebalance(device=device, group=group))
    try:
        res = problem.solve(part_rank=rank, part_count=world, **solve_kw)
    finally:
        if exchange and world > 1 and hasattr(problem, "set_exchange"):
            problem.set_exchange(None)
        if rebalance and world > 1 and hasattr(problem, "set_rebalance"):
            problem.set_rebalance(None)
    return reduce_results(res, objective, device=device, group=group), res
