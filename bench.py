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
    ap.add_argument("--workload", default="queens", choices=["queens", "wcet", "sat200"],
                    help="queens: N-queens ALL (BASELINE headline, config 3); wcet: examples/wcet.txt MAX (config 4, "
                         "incumbent shared between the GPUs); sat200: random 3-SAT n=200 seed 1 ANY (config 5, UNSAT)")
    ap.add_argument("--no-comm", action="store_true",
                    help="N > 1: static path-hash partition of a replicated frontier instead of the shared frontier of a csolve_gpu_comm")
    ap.add_argument("--order", default="none")
    ap.add_argument("--cpu-queens", type=int, default=0,
                    help="board size of the bounded CPU sample of the queens workload (default: 13-queens, ~10 s per run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split-target", type=int, default=0, help="frames of the expanded root frontier (0: the library's default)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation (oracle/_ref, built from /root/reference where it was
# available) or, failing that, the oracle port. Only used as the reported baseline / reference arm.
def host_info():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"host_cores": os.cpu_count(), "cpu_model": model}


def workload_text(args, n_queens=None):
    from csolve_b200 import instances as I
    if args.workload == "wcet":
        return I.wcet()
    if args.workload == "sat200":
        return I.random_3sat(200, seed=1)
    return I.queens(n_queens or args.queens)


def cpu_reference_run(n_queens, text=None, flags=()):
    """one single-threaded run of the reference CLI; returns (nodes, seconds, kind, solutions)"""
    from csolve_b200 import instances as I
    if text is None:
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
        out = subprocess.run([ref_cli, "-s", "0", *flags, f.name], capture_output=True, text=True).stdout
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


# reference CALLS of 16-queens ALL (measured once in the build container: 6 629.7 s there)
REF_CALLS_QUEENS16 = 1048203447


def reference_sample(args):
    """(text, flags, description) of the bounded CPU sample of the workload: the reference needs ~1 h for 16-queens,
    83 s for wcet and 37 s for 3-SAT n=200 seed 1, so a step of the reference arm is a smaller board / a time-boxed run
    of the same instance (the CLI's own -t; CALLS are printed when it stops)."""
    if args.workload == "wcet":
        return workload_text(args), ("-t", "10"), "wcet MAX, first 10 s of the search (-t 10), single thread, default flags"
    if args.workload == "sat200":
        return workload_text(args), ("-t", "10", "-c", "false"), "3-SAT n=200 seed 1 ANY, first 10 s (-t 10), single thread, -c false"
    n = args.cpu_queens or 13
    return workload_text(args, n), (), "%d-queens all-solutions, single thread, default flags" % n


def workload_name(args):
    return {"queens": "queens%d-all" % args.queens, "wcet": "wcet-max", "sat200": "sat200-seed1-any"}[args.workload]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    text, flags, desc = reference_sample(args)
    if args.warmup > 0:                      # one untimed run is enough to page the binary in
        cpu_reference_run(8)
    nodes = 0
    secs = 0.0
    kind = "reference"
    for _ in range(args.steps):
        c, dt, kind, sols = cpu_reference_run(0, text, flags)
        nodes += c
        secs += dt
    value = nodes / secs
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args), "sample": desc},
        "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                              "sample": "%s; %d nodes per step; the reference's -j fork mode does not scale (BASELINE.md)"
                                        % (desc, nodes // max(args.steps, 1))}, **host_info()),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.workload == "queens" and args.queens == 16:
        # like-for-like time to solution: the reference's CALLS for 16-queens at the node rate this host just showed
        # (its rate FALLS with the board size -- BASELINE.md: 201 k/s at N=12, 158 k/s at N=16 -- so this flatters the CPU)
        line["config"]["time_to_solution_s_calibrated"] = REF_CALLS_QUEENS16 / value
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
    (profiles/r2_traffic.json, r1_traffic.json; dram__bytes_read.sum + dram__bytes_write.sum), None when there is no capture"""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            for e in (t if isinstance(t, list) else [t]):
                if e.get("kernel") == kernel and e.get("workload") == workload:
                    return float(e["dram_bytes_per_launch"])
        except Exception:
            pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import csolve_b200 as cb
    from csolve_b200 import distributed as D

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
    comm = None
    mode = "single GPU"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        mode = "tree-partition x%d (path hash of a replicated frontier)" % world
        if not args.no_comm:
            # the ranks' segments are mapped into each other once (CUDA IPC handles through one all-gather); a box
            # that cannot do that falls back to the static partition -- on EVERY rank, so they decide together
            ok = 1
            try:
                comm = D.make_comm(local)
            except Exception as e:                      # noqa: BLE001
                print("[bench rank %d] comm unavailable (%s): static partition" % (rank, e), file=sys.stderr, flush=True)
                ok = 0
            t = torch.tensor([ok], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0:
                comm = None
            else:
                if args.workload == "queens":
                    mode = ("csolve_gpu_comm x%d: ALL model, every rank expands the root and searches the frames of its "
                            "path-hash share (no data-path exchange; results summed)" % world)
                else:
                    mode = ("csolve_gpu_comm x%d: shared root frontier, incumbents / first solution and donated frames "
                            "over NVLink peer memory" % world)

    text = workload_text(args)
    order = cb.host.ORDER_NAMES[args.order]
    objective = {"queens": cb.OBJ_ALL, "wcet": cb.OBJ_MAX, "sat200": cb.OBJ_ANY}[args.workload]
    solve_kw = dict(order=order)
    if args.split_target > 0:
        solve_kw["split_target"] = args.split_target
    if args.workload == "sat200":
        solve_kw["prefer_failing"] = True          # the reference's defaults: -f true ...
        if world == 1:
            solve_kw["restart_frequency"] = 100    # ... -r 100 (Luby restarts; single-GPU searches only)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def one_step():
        """the call a user makes: parse + root phase, upload, search, read the result back -- and, with several
        ranks, the reduction that puts the whole job's result on every rank; the clock stops after it"""
        t0 = time.perf_counter()
        model = cb.Model(text)                          # host front end
        prob = cb.GpuProblem(model, device=local)       # H2D: compiled model
        if comm is not None:
            res = prob.solve(comm=comm, **solve_kw)
        else:
            res = prob.solve(part_rank=rank, part_count=world, **solve_kw)   # D2H: counters, incumbent, status
        red = D.reduce_results(res, objective, device=dev)
        wall = time.perf_counter() - t0
        f = model.flat
        # model arrays + root frame + the 640-byte control block once per expansion level and once for the search
        levels = max(int(res.kernel_launches) - 3, 0)
        h2d = (f.n_clauses * 16 + (f.n_vars + 1) * 4 + f.n_watch * 4 + f.n_nodes * 13 + f.n_vars * 16
               + (8 + f.n_vars * 2 + 4) * 4 + 640 * (levels + 1))
        d2h = 640 * (levels + 2) + 96                   # control block per expansion level / slice / at the end + counters
        prob.close(); model.close()
        return res, red, wall, h2d, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler starts with the warm-up steps (the same work as the timed ones): nvidia-smi needs a few hundred
    # milliseconds to deliver its first row, and 10 timed steps on 8 GPUs are over in less than that
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
        flush.fill_(1)
    barrier()
    tot_nodes = tot_sols = tot_launch = 0
    dev_ms = wall_s = search_ms = 0.0
    my_nodes = 0
    h2d = d2h = 0
    best = None
    for _ in range(args.steps):
        flush.fill_(int(time.time()) & 1)               # flush L2 between timed iterations
        barrier()
        res, red, wall, h2d, d2h = one_step()
        w = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        tot_nodes += red["nodes"]; tot_sols = red["solutions"]; tot_launch += red["kernel_launches"]
        best = red["best"] if red["has_solution"] else None
        dev_ms += red["kernel_ms"] + red["expand_ms"]   # max over ranks, device clock (CUDA events on the library's stream)
        search_ms += res.kernel_ms
        print("[bench rank %d] nodes=%d search_ms=%.2f expand_ms=%.2f launches=%d wall_ms=%.2f" % (
            rank, res.nodes, res.kernel_ms, res.expand_ms, res.kernel_launches, wall * 1e3), file=sys.stderr, flush=True)
        my_nodes += res.nodes
        wall_s += float(w.item())
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # results identical to the reference's: solution count (OEIS A000170), optimum, SAT status
    if args.workload == "queens":
        expected = {4: 2, 5: 10, 6: 4, 7: 40, 8: 92, 9: 352, 10: 724, 11: 2680, 12: 14200, 13: 73712, 14: 365596,
                    15: 2279184, 16: 14772512, 17: 95815104}.get(args.queens)
        if expected is not None and tot_sols != expected:
            raise SystemExit("bench.py: wrong solution count %d (expected %d)" % (tot_sols, expected))
    elif args.workload == "wcet" and best != 1560:
        raise SystemExit("bench.py: wrong optimum %r (expected 1560)" % (best,))
    elif args.workload == "sat200" and tot_sols != 0:
        raise SystemExit("bench.py: 3-SAT n=200 seed 1 is unsatisfiable, got a solution")

    if rank == 0:
        V = {"queens": args.queens, "wcet": 12, "sat200": 200}[args.workload]
        bytes_per_node = 2 * (8 * V + 16)                       # parent domains + header in, child out
        peak, peak_src = measured_peak_gbs()
        achieved = (my_nodes * bytes_per_node) / (search_ms / 1000.0) / 1e9 if search_ms > 0 else 0.0
        kernel = {"queens": "k_search_lov<false,true>" if args.queens <= 32 else "k_search<false>",
                  "wcet": "k_search<false,false,LIN=true>", "sat200": "k_search_sat<false>"}[args.workload]
        line = {
            "metric": METRIC, "value": tot_nodes / (dev_ms / 1000.0), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args), "order": args.order, "solutions": tot_sols, "best": best,
                       "nodes_per_step": tot_nodes // args.steps, "parallelism": mode,
                       "l2": "flushed between iterations (256 MiB write)",
                       "time_to_solution_s": dev_ms / 1000.0 / args.steps},
            "clocks": clocks,
            "e2e": {"value": tot_nodes / wall_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "time_to_solution_s": wall_s / args.steps,
                    "path": "Model(text) -> GpuProblem -> solve() through libcsolve_b200.so, host buffers"
                            + ("; the clock stops after the all-reduce of the ranks' results" if world > 1 else "")},
            "gpu_launches": tot_launch,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(kernel, workload_name(args)), "kernel": kernel,
                         "bytes_per_node": bytes_per_node, "peak_source": peak_src,
                         "note": "rank 0; algorithmic bytes = nodes x 2 x (8V+16); the DFS stacks live in shared memory during a slice, so real DRAM traffic is far below this nominal figure (DESIGN.md)"},
        }
        if not args.no_cpu_baseline and world == 1:
            rtext, rflags, rdesc = reference_sample(args)
            c, dt, kind, sols = cpu_reference_run(0, rtext, rflags)
            line["cpu_baseline"] = dict({"value": c / dt, "unit": UNIT, "cores": 1, "kind": kind,
                                         "sample": "%s: %d nodes in %.2f s" % (rdesc, c, dt)}, **host_info())
            if args.workload == "queens" and args.queens == 16:
                # like-for-like: the reference's 16-queens CALLS at the node rate this host just showed (which flatters
                # the CPU: its rate falls with the board size) against the measured end-to-end time
                cal = REF_CALLS_QUEENS16 / (c / dt)
                line["cpu_baseline"]["time_to_solution_s_calibrated"] = cal
                line["e2e"]["time_to_solution_speedup_calibrated"] = cal / (wall_s / args.steps)
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        if comm is not None:
            comm.close()
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
