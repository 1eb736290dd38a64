#!/usr/bin/env python3
"""Benchmark of the csolve search hot path (BASELINE.json: search nodes/sec and time-to-solution,
16-queens all-solutions, 1/2/4/8 B200 vs the reference CPU solver).

    python bench.py --gpus N --steps K --warmup W            (ours; under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path)

One step = one complete all-solutions search of the workload (default: 16-queens, the instance
scripts/gen_queens.sh writes, header ALL). For N > 1 every rank searches its share of the tree
(path-hash partition of the root frontier, no data-path collective); `value` = nodes searched by
all ranks / max-over-ranks device time. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "search_nodes_per_sec"
UNIT = "nodes/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queens", type=int, default=16, help="board size of the workload (BASELINE config 3: 14..16)")
    ap.add_argument("--order", default="none")
    ap.add_argument("--cpu-queens", type=int, default=0,
                    help="board size of the bounded CPU sample (default: 13-queens, ~10 s, for cpu_baseline; 12-queens, "
                         "~2 s per step, for the steps of --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref, built from /root/reference where it was
# available) or, failing that, the oracle port. Only used as the reported baseline / reference arm.
def cpu_reference_run(n_queens):
    """one single-threaded all-solutions run; returns (nodes, seconds, kind, solutions)"""
    from csolve_b200 import instances as I
    text = I.queens(n_queens)
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "csolve_ref")
    if os.path.exists(ref_cli):
        import re
        import tempfile
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(text)
        t0 = time.perf_counter()
        # the reference's defaults (-c true -f true -w true -o none -r 100 -j 1), stats printing off,
        # solutions discarded by the pipe reader (it prints every solution)
        out = subprocess.run([ref_cli, "-s", "0", f.name], capture_output=True, text=True).stdout
        dt = time.perf_counter() - t0
        os.unlink(f.name)
        m = re.search(r"CALLS: (\d+).*SOLUTIONS: (\d+)", out)
        return int(m.group(1)), dt, "reference", int(m.group(2))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import csolve_b200 as cb
    import util
    m = cb.Model(text)
    o = util.Oracle(m)
    t0 = time.perf_counter()
    r, _ = o.solve_reference()
    dt = time.perf_counter() - t0
    return int(r.calls), dt, "port", int(r.solutions)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_queens or 12
    for _ in range(max(args.warmup, 0) and 1):   # one untimed run is enough to page the binary in
        cpu_reference_run(min(n, 10))
    nodes = 0
    secs = 0.0
    kind = "reference"
    for _ in range(args.steps):
        c, dt, kind, sols = cpu_reference_run(n)
        nodes += c
        secs += dt
    value = nodes / secs
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "impl": "reference",
        "config": {"workload": "queens%d-all" % args.queens, "sample": "queens%d-all" % n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": "%d-queens all-solutions per step (%d nodes), single thread, default flags; "
                                   "the reference's -j fork mode does not scale (BASELINE.md)" % (n, nodes // max(args.steps, 1))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, workload):
    """real DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r1_traffic.json; dram__bytes_read.sum + dram__bytes_write.sum), None when there is no capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)
        if t.get("kernel") == kernel and t.get("workload") == workload:
            return float(t["dram_bytes_per_launch"])
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import csolve_b200 as cb
    from csolve_b200 import distributed as D
    from csolve_b200 import instances as I

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the search path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line: while the job runs, file descriptor 1 points at stderr, so whatever a
    # library prints there (NCCL prints its version banner at any NCCL_DEBUG level >= VERSION) cannot get in front of it
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    text = I.queens(args.queens)
    order = cb.host.ORDER_NAMES[args.order]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def one_step():
        """the call a user makes: parse + root phase, upload, search, read the result back"""
        t0 = time.perf_counter()
        model = cb.Model(text)                          # host front end
        prob = cb.GpuProblem(model, device=local)       # H2D: compiled model
        res = prob.solve(order=order, part_rank=rank, part_count=world)   # D2H: counters, incumbent, status
        wall = time.perf_counter() - t0
        f = model.flat
        # model arrays + root frame + the 640-byte control block once per expansion level and once for the search
        levels = max(int(res.kernel_launches) - 3, 0)
        h2d = (f.n_clauses * 16 + (f.n_vars + 1) * 4 + f.n_watch * 4 + f.n_nodes * 13 + f.n_vars * 16
               + (8 + f.n_vars * 2 + 4) * 4 + 640 * (levels + 1))
        d2h = 640 * (levels + 2) + 96                   # control block per expansion level / slice / at the end + counters
        prob.close(); model.close()
        return res, wall, h2d, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
        flush.fill_(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    tot_nodes = tot_sols = tot_launch = 0
    dev_ms = wall_s = search_ms = 0.0
    my_nodes = 0
    h2d = d2h = 0
    objective = cb.OBJ_ALL
    for _ in range(args.steps):
        flush.fill_(int(time.time()) & 1)               # flush L2 between timed iterations
        barrier()
        res, wall, h2d, d2h = one_step()
        red = D.reduce_results(res, objective, device=dev)
        w = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        tot_nodes += red["nodes"]; tot_sols = red["solutions"]; tot_launch += red["kernel_launches"]
        dev_ms += red["kernel_ms"] + red["expand_ms"]   # max over ranks, device clock (CUDA events on the library's stream)
        search_ms += res.kernel_ms
        print("[bench rank %d] nodes=%d search_ms=%.2f expand_ms=%.2f launches=%d wall_ms=%.2f" % (
            rank, res.nodes, res.kernel_ms, res.expand_ms, res.kernel_launches, wall * 1e3), file=sys.stderr, flush=True)
        my_nodes += res.nodes
        wall_s += float(w.item())
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    expected = {4: 2, 5: 10, 6: 4, 7: 40, 8: 92, 9: 352, 10: 724, 11: 2680, 12: 14200, 13: 73712, 14: 365596,
                15: 2279184, 16: 14772512, 17: 95815104}.get(args.queens)
    if expected is not None and tot_sols != expected:
        raise SystemExit("bench.py: wrong solution count %d (expected %d)" % (tot_sols, expected))

    if rank == 0:
        V = args.queens
        bytes_per_node = 2 * (8 * V + 16)                       # parent domains + header in, child out
        peak, peak_src = measured_peak_gbs()
        achieved = (my_nodes * bytes_per_node) / (search_ms / 1000.0) / 1e9 if search_ms > 0 else 0.0
        kernel = "k_search_lov<false,true>" if args.queens <= 32 else "k_search<false>"
        line = {
            "metric": METRIC, "value": tot_nodes / (dev_ms / 1000.0), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "queens%d-all" % V, "order": args.order, "solutions": tot_sols,
                       "nodes_per_step": tot_nodes // args.steps, "parallelism": "tree-partition x%d" % world,
                       "l2": "flushed between iterations (256 MiB write)",
                       "time_to_solution_s": dev_ms / 1000.0 / args.steps},
            "clocks": clocks,
            "e2e": {"value": tot_nodes / wall_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "time_to_solution_s": wall_s / args.steps,
                    "path": "Model(text) -> GpuProblem -> solve() through libcsolve_b200.so, host buffers"},
            "gpu_launches": tot_launch,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(kernel, "queens%d-all" % V), "kernel": kernel,
                         "bytes_per_node": bytes_per_node, "peak_source": peak_src,
                         "note": "rank 0; algorithmic bytes = nodes x 2 x (8V+16); the DFS stacks live in shared memory during a slice, so real DRAM traffic is far below this nominal figure (DESIGN.md)"},
        }
        if not args.no_cpu_baseline and world == 1:
            nq = args.cpu_queens or 13
            c, dt, kind, sols = cpu_reference_run(nq)
            line["cpu_baseline"] = {"value": c / dt, "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": "%d-queens all-solutions, %d nodes in %.2f s, single thread, default flags"
                                              % (nq, c, dt)}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)
