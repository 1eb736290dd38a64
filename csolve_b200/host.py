"""ctypes mirror of include/csolve_b200.h.

Names follow the reference's interface for this path: a *model* is what the
front end (src/parser.y:55-85) hands to ``solve()`` (src/csolve.h:395); a
``GpuProblem`` is that model resident on the device; ``GpuProblem.solve()`` is the
replacement for ``solve()``; ``GpuProblem.propagate_batch()`` replays node
transitions (bind + propagate_clauses, src/csolve.c:448-457) in bulk.

There is no CPU fallback: if the CUDA library is missing or no device is
present the calls raise ``CsolveError``.
"""
import ctypes as C
import os

import numpy as np

OBJ_ANY, OBJ_ALL, OBJ_MIN, OBJ_MAX = 0, 1, 2, 3
ORDER_NONE, ORDER_SMALLEST_DOMAIN, ORDER_LARGEST_DOMAIN, ORDER_SMALLEST_VALUE, ORDER_LARGEST_VALUE = range(5)
ORDER_NAMES = {
    "none": ORDER_NONE,
    "smallest-domain": ORDER_SMALLEST_DOMAIN,
    "largest-domain": ORDER_LARGEST_DOMAIN,
    "smallest-value": ORDER_SMALLEST_VALUE,
    "largest-value": ORDER_LARGEST_VALUE,
}  # the -o values of src/main.c:187-209

ERR_NAMES = {
    -1: "invalid argument", -2: "syntax error", -3: "infeasible problem", -4: "unbounded variable",
    -5: "unsupported construct", -6: "CUDA error", -7: "no CUDA device", -8: "capacity exceeded",
}


class CsolveError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s (%d): %s" % (ERR_NAMES.get(code, "error"), code, message))
        self.code = code
        self.message = message


class FlatModel(C.Structure):
    """struct csolve_flat_model"""
    _fields_ = [
        ("n_vars", C.c_int32), ("n_nodes", C.c_int32), ("n_clauses", C.c_int32), ("n_watch", C.c_int32),
        ("objective", C.c_int32), ("obj_var", C.c_int32),
        ("node_op", C.POINTER(C.c_uint8)), ("node_l", C.POINTER(C.c_int32)), ("node_r", C.POINTER(C.c_int32)),
        ("clause_first", C.POINTER(C.c_int32)), ("watch_ptr", C.POINTER(C.c_int32)),
        ("watch_idx", C.POINTER(C.c_int32)), ("var_lo", C.POINTER(C.c_int32)), ("var_hi", C.POINTER(C.c_int32)),
        ("var_prio", C.POINTER(C.c_int64)), ("var_name", C.POINTER(C.c_char_p)),
    ]

    def to_dict(self):
        def arr(p, n):
            return [p[i] for i in range(n)]
        return dict(
            n_vars=self.n_vars, n_nodes=self.n_nodes, n_clauses=self.n_clauses, n_watch=self.n_watch,
            objective=self.objective, obj_var=self.obj_var,
            node_op=arr(self.node_op, self.n_nodes), node_l=arr(self.node_l, self.n_nodes),
            node_r=arr(self.node_r, self.n_nodes), clause_first=arr(self.clause_first, self.n_clauses + 1),
            watch_ptr=arr(self.watch_ptr, self.n_vars + 1), watch_idx=arr(self.watch_idx, self.n_watch),
            var_lo=arr(self.var_lo, self.n_vars), var_hi=arr(self.var_hi, self.n_vars),
            var_prio=arr(self.var_prio, self.n_vars),
            var_name=[self.var_name[i].decode() for i in range(self.n_vars)] if self.var_name else None,
        )


class _FrontOptions(C.Structure):
    _fields_ = [("compute_weights", C.c_int32), ("objective_override", C.c_int32)]


class _GpuConfig(C.Structure):
    _fields_ = [("device", C.c_int32)]


class _SolveOptions(C.Structure):
    _fields_ = [("order", C.c_int32), ("part_rank", C.c_int32), ("part_count", C.c_int32),
                ("split_target", C.c_int32), ("max_solutions", C.c_int32), ("time_limit_ms", C.c_int32),
                ("slice_ms", C.c_int32), ("create_conflicts", C.c_int32), ("backjump", C.c_int32),
                ("prefer_failing", C.c_int32), ("sample_mod", C.c_uint32), ("sample_cap", C.c_int32),
                ("restart_frequency", C.c_int32), ("sample_failed_keep", C.c_uint32), ("reserved2", C.c_int32 * 4)]


SAMPLE_FAILED, SAMPLE_COUNTED, SAMPLE_LEAF = 1, 2, 4   # CSOLVE_SAMPLE_* flags of a sampled search node


class _GpuResult(C.Structure):
    _fields_ = [("solutions", C.c_uint64), ("nodes", C.c_uint64), ("cuts", C.c_uint64), ("props", C.c_uint64),
                ("clause_visits", C.c_uint64), ("best", C.c_int32), ("has_solution", C.c_int32),
                ("timed_out", C.c_int32), ("n_stored", C.c_int32), ("kernel_ms", C.c_double),
                ("expand_ms", C.c_double), ("kernel_launches", C.c_uint64), ("conflicts", C.c_uint64),
                ("conflicts_abandoned", C.c_uint64), ("restarts", C.c_uint64), ("backjumps", C.c_uint64)]


# int (*csolve_exchange_fn)(void *user, int32_t *best, int32_t *found, int32_t local_done)
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32)


# int (*csolve_rebalance_fn)(void *user, csolve_gpu_problem *p, int32_t n_idle, int32_t n_busy, int32_t frame_words)
REBALANCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32)


def library_path():
    # CSOLVE_B200_LIB: development override to compare builds; the default is the in-tree library
    return os.environ.get("CSOLVE_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcsolve_b200.so")


_lib = None


def library():
    """Load libcsolve_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise CsolveError(-6, "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback for the search path)" % path)
    lib = C.CDLL(path)
    I32P, U8P = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    lib.csolve_model_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_FrontOptions), C.POINTER(C.c_void_p)]
    lib.csolve_model_flat.argtypes = [C.c_void_p]
    lib.csolve_model_flat.restype = C.POINTER(FlatModel)
    lib.csolve_model_free.argtypes = [C.c_void_p]
    lib.csolve_model_free.restype = None
    lib.csolve_gpu_init.argtypes = [C.POINTER(_GpuConfig)]
    lib.csolve_gpu_shutdown.restype = None
    lib.csolve_gpu_load.argtypes = [C.POINTER(FlatModel), C.POINTER(C.c_void_p)]
    lib.csolve_gpu_load_device.argtypes = [C.POINTER(FlatModel), C.c_int32, C.POINTER(C.c_void_p)]
    lib.csolve_gpu_get_samples.argtypes = [C.c_void_p, I32P, C.c_int32, I32P, I32P]
    lib.csolve_gpu_unload.argtypes = [C.c_void_p]
    lib.csolve_gpu_unload.restype = None
    lib.csolve_gpu_propagate_batch.argtypes = [C.c_void_p, C.c_int32, I32P, I32P, I32P, I32P, I32P, U8P]
    lib.csolve_gpu_solve.argtypes = [C.c_void_p, C.POINTER(_SolveOptions), C.POINTER(_GpuResult)]
    lib.csolve_gpu_get_solution.argtypes = [C.c_void_p, C.c_int32, I32P]
    lib.csolve_gpu_get_nogoods.argtypes = [C.c_void_p, I32P, C.c_int32, I32P, C.c_int32, I32P]
    lib.csolve_gpu_get_solution_key.argtypes = [C.c_void_p, C.c_int32, I32P]
    lib.csolve_gpu_solve_batch.argtypes = [C.c_void_p, C.POINTER(_SolveOptions), C.c_int32, I32P,
                                           C.POINTER(C.c_uint32), U8P, C.POINTER(_GpuResult)]
    lib.csolve_gpu_set_exchange.argtypes = [C.c_void_p, EXCHANGE_FN, C.c_void_p]
    lib.csolve_gpu_set_rebalance.argtypes = [C.c_void_p, REBALANCE_FN, C.c_void_p]
    lib.csolve_gpu_export_frames.argtypes = [C.c_void_p, C.c_int32, I32P, I32P]
    lib.csolve_gpu_import_frames.argtypes = [C.c_void_p, I32P, C.c_int32]
    lib.csolve_gpu_comm_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.csolve_gpu_comm_handle.argtypes = [C.c_void_p, C.c_void_p]
    lib.csolve_gpu_comm_connect.argtypes = [C.c_void_p, C.c_void_p]
    lib.csolve_gpu_comm_connect_local.argtypes = [C.POINTER(C.c_void_p), C.c_int32]
    lib.csolve_gpu_comm_destroy.argtypes = [C.c_void_p]
    lib.csolve_gpu_comm_destroy.restype = None
    lib.csolve_gpu_solve_comm.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_SolveOptions), C.POINTER(_GpuResult)]
    lib.csolve_gpu_device_count.argtypes = [I32P]
    lib.csolve_gpu_group_create.argtypes = [C.c_int32, I32P, C.c_size_t, C.POINTER(C.c_void_p)]
    lib.csolve_gpu_group_load.argtypes = [C.c_void_p, C.POINTER(FlatModel)]
    lib.csolve_gpu_group_solve.argtypes = [C.c_void_p, C.POINTER(_SolveOptions), C.POINTER(_GpuResult), C.POINTER(_GpuResult)]
    lib.csolve_gpu_group_get_solution.argtypes = [C.c_void_p, C.c_int32, I32P, I32P]
    lib.csolve_gpu_group_destroy.argtypes = [C.c_void_p]
    lib.csolve_gpu_group_destroy.restype = None
    lib.csolve_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise CsolveError(rc, library().csolve_last_error().decode(errors="replace"))


class Model:
    """A parsed, root-normalised and flattened csolve instance (host side)."""

    def __init__(self, text, compute_weights=True, objective=None):
        if isinstance(text, str):
            text = text.encode()
        lib = library()
        opt = _FrontOptions(1 if compute_weights else 0, -1 if objective is None else int(objective))
        h = C.c_void_p()
        _check(lib.csolve_model_parse(text, len(text), C.byref(opt), C.byref(h)))
        self._h = h
        self.flat = lib.csolve_model_flat(h).contents
        self.flat._owner = self   # the arrays belong to the C model: keep it alive as long as the view

    @classmethod
    def from_file(cls, path, **kw):
        with open(path, "rb") as f:
            return cls(f.read(), **kw)

    n_vars = property(lambda self: self.flat.n_vars)
    objective = property(lambda self: self.flat.objective)
    obj_var = property(lambda self: self.flat.obj_var)

    @property
    def var_names(self):
        return [self.flat.var_name[i].decode() for i in range(self.flat.n_vars)]

    @property
    def root_domains(self):
        v = self.flat.n_vars
        out = np.empty(2 * v, np.int32)
        out[0::2] = [self.flat.var_lo[i] for i in range(v)]
        out[1::2] = [self.flat.var_hi[i] for i in range(v)]
        return out

    def close(self):
        if getattr(self, "_h", None):
            library().csolve_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SolveResult:
    def __init__(self, r, solutions):
        for name, _ in _GpuResult._fields_:
            setattr(self, name, getattr(r, name))
        self.assignments = solutions

    def __repr__(self):
        return ("SolveResult(solutions=%d, nodes=%d, cuts=%d, props=%d, best=%d, has_solution=%d, kernel_ms=%.3f)"
                % (self.solutions, self.nodes, self.cuts, self.props, self.best, self.has_solution, self.kernel_ms))


def solve_options(order=ORDER_NONE, part_rank=0, part_count=1, split_target=0, max_solutions=0, time_limit_ms=0,
                  slice_ms=0, prefer_failing=False, create_conflicts=False, sample_mod=0, sample_cap=0,
                  restart_frequency=0, sample_failed_keep=1, backjump=False):
    if isinstance(order, str):
        order = ORDER_NAMES[order]
    return _SolveOptions(order, part_rank, part_count, split_target, max_solutions, time_limit_ms, slice_ms,
                         1 if create_conflicts else 0, 1 if backjump else 0, 1 if prefer_failing else 0, int(sample_mod), int(sample_cap),
                         int(restart_frequency), int(sample_failed_keep))


COMM_HANDLE_BYTES = 64


class Comm:
    """One rank of a csolve_gpu_comm (include/csolve_b200.h): the ranks search one tree together over NVLink peer
    memory. One process per GPU: create, all-gather `handle()`, `connect(handles)` (distributed.make_comm does that
    over torch.distributed)."""

    def __init__(self, device, rank, world, frontier_bytes=0):
        h = C.c_void_p()
        _check(library().csolve_gpu_comm_create(int(device), int(rank), int(world), int(frontier_bytes), C.byref(h)))
        self._h = h
        self.rank, self.world = int(rank), int(world)

    def handle(self):
        buf = (C.c_ubyte * COMM_HANDLE_BYTES)()
        _check(library().csolve_gpu_comm_handle(self._h, buf))
        return bytes(buf)

    def connect(self, handles):
        """handles: the ranks' handle() bytes in rank order"""
        blob = b"".join(handles)
        assert len(blob) == self.world * COMM_HANDLE_BYTES
        _check(library().csolve_gpu_comm_connect(self._h, blob))

    def close(self):
        if getattr(self, "_h", None):
            library().csolve_gpu_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_count():
    n = C.c_int32()
    _check(library().csolve_gpu_device_count(C.byref(n)))
    return n.value


class GpuGroup:
    """Several GPUs of THIS process on one search tree (csolve_gpu_group_*: one host thread per device inside the
    library -- what the drop-in solve() uses for `-j N`)."""

    def __init__(self, n_devices, devices=None, frontier_bytes=0):
        h = C.c_void_p()
        dev = None
        if devices is not None:
            dev = (C.c_int32 * n_devices)(*devices)
        _check(library().csolve_gpu_group_create(int(n_devices), dev, int(frontier_bytes), C.byref(h)))
        self._h = h
        self.n_devices = int(n_devices)
        self.n_vars = 0

    def load(self, model):
        flat = model.flat if isinstance(model, Model) else model
        self.n_vars = flat.n_vars
        _check(library().csolve_gpu_group_load(self._h, C.byref(flat)))

    def solve(self, **kw):
        """-> (SolveResult of the whole job, [SolveResult per device])"""
        opt = solve_options(**kw)
        res = _GpuResult()
        per = (_GpuResult * self.n_devices)()
        _check(library().csolve_gpu_group_solve(self._h, C.byref(opt), C.byref(res), per))
        sols = []
        buf = (C.c_int32 * self.n_vars)()
        for i in range(res.n_stored):
            _check(library().csolve_gpu_group_get_solution(self._h, i, buf, None))
            sols.append(list(buf))
        return SolveResult(res, sols), [SolveResult(per[i], []) for i in range(self.n_devices)]

    def close(self):
        if getattr(self, "_h", None):
            library().csolve_gpu_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GpuProblem:
    """A model resident on one GPU."""

    def __init__(self, model, device=0):
        lib = library()
        self.model = model
        flat = model.flat if isinstance(model, Model) else model
        self.n_vars = flat.n_vars
        h = C.c_void_p()
        _check(lib.csolve_gpu_load_device(C.byref(flat), int(device), C.byref(h)))
        self._h = h

    def propagate_batch(self, dom_in, var, val, best=None):
        """dom_in: [B, 2*V] int32 (lo,hi pairs); var, val: [B]; returns (dom_out [B, 2*V], failed [B] uint8)."""
        I32P, U8P = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        dom_in = np.ascontiguousarray(dom_in, np.int32).reshape(-1, 2 * self.n_vars)
        n = dom_in.shape[0]
        var = np.ascontiguousarray(var, np.int32).reshape(n)
        val = np.ascontiguousarray(val, np.int32).reshape(n)
        bestp = None
        if best is not None:
            best = np.ascontiguousarray(np.broadcast_to(np.asarray(best, np.int32), (n,)))
            bestp = best.ctypes.data_as(I32P)
        out = np.empty_like(dom_in)
        failed = np.empty(n, np.uint8)
        _check(library().csolve_gpu_propagate_batch(
            self._h, n, dom_in.ctypes.data_as(I32P), var.ctypes.data_as(I32P), val.ctypes.data_as(I32P),
            bestp, out.ctypes.data_as(I32P), failed.ctypes.data_as(U8P)))
        return out, failed

    def solve(self, order=ORDER_NONE, part_rank=0, part_count=1, split_target=0, max_solutions=0,
              time_limit_ms=0, slice_ms=0, prefer_failing=False, create_conflicts=False, sample_mod=0, sample_cap=0,
              restart_frequency=0, sample_failed_keep=1, comm=None, backjump=False):
        """comm: a connected Comm -- the call is then COLLECTIVE over the comm's ranks (csolve_gpu_solve_comm): they
        search one tree together and each gets its own share of the counters back."""
        opt = solve_options(order, part_rank, part_count, split_target, max_solutions, time_limit_ms, slice_ms,
                            prefer_failing, create_conflicts, sample_mod, sample_cap, restart_frequency, sample_failed_keep,
                            backjump)
        res = _GpuResult()
        if comm is not None:
            _check(library().csolve_gpu_solve_comm(self._h, comm._h, C.byref(opt), C.byref(res)))
        else:
            _check(library().csolve_gpu_solve(self._h, C.byref(opt), C.byref(res)))
        sols = []
        buf = (C.c_int32 * self.n_vars)()
        for i in range(res.n_stored):
            _check(library().csolve_gpu_get_solution(self._h, i, buf))
            sols.append(list(buf))
        return SolveResult(res, sols)

    def samples(self, cap=1 << 20):
        """search nodes recorded by the last solve(sample_mod=k): dict of arrays flags [n], var [n], val [n], best [n],
        parent [n, 2V], child [n, 2V] (+ 'seen': hits including the ones dropped when the buffer was full)"""
        W = 4 + 4 * self.n_vars
        n, seen = C.c_int32(), C.c_int32()
        I32P = C.POINTER(C.c_int32)
        _check(library().csolve_gpu_get_samples(self._h, None, 0, C.byref(n), C.byref(seen)))
        cap = min(cap, max(seen.value, 1))
        buf = np.zeros((cap, W), np.int32)
        _check(library().csolve_gpu_get_samples(self._h, buf.ctypes.data_as(I32P), cap, C.byref(n), C.byref(seen)))
        buf = buf[:n.value]
        V = self.n_vars
        return dict(flags=buf[:, 0].copy(), var=buf[:, 1].copy(), val=buf[:, 2].copy(), best=buf[:, 3].copy(),
                    parent=buf[:, 4:4 + 2 * V].copy(), child=buf[:, 4 + 2 * V:].copy(), seen=seen.value)

    def nogoods(self, cap_lits=1 << 22, cap_ng=1 << 18):
        """nogoods learned by the last solve(create_conflicts=True): list of [(var, value), ...]"""
        lits = np.zeros(cap_lits, np.int32)
        starts = np.zeros(cap_ng + 1, np.int32)
        n = C.c_int32()
        I32P = C.POINTER(C.c_int32)
        _check(library().csolve_gpu_get_nogoods(self._h, lits.ctypes.data_as(I32P), cap_lits, starts.ctypes.data_as(I32P),
                                                cap_ng, C.byref(n)))
        return [[(int(c) >> 1, int(c) & 1) for c in lits[starts[k]:starts[k + 1]]] for k in range(n.value)]

    def set_exchange(self, fn):
        """fn(best, found, local_done) -> (best, found, all_done): called once per time slice (see
        csolve_gpu_set_exchange in include/csolve_b200.h). Pass None to remove it."""
        if fn is None:
            self._exchange_cb = None
            _check(library().csolve_gpu_set_exchange(self._h, C.cast(None, EXCHANGE_FN), None))
            return

        def tramp(user, best_p, found_p, local_done):
            b, f, done = fn(int(best_p[0]), int(found_p[0]), int(local_done))
            best_p[0] = int(b)
            found_p[0] = int(f)
            return 1 if done else 0
        self._exchange_cb = EXCHANGE_FN(tramp)          # keep the trampoline alive
        _check(library().csolve_gpu_set_exchange(self._h, self._exchange_cb, None))

    def set_rebalance(self, fn):
        """fn(problem, n_idle, n_busy, frame_words) -> number of frames imported: called once per time slice after the
        exchange callback while some rank still has work (csolve_gpu_set_rebalance in include/csolve_b200.h). Inside
        it export_frames() / import_frames() move frames between the ranks. Pass None to remove it."""
        if fn is None:
            self._rebalance_cb = None
            _check(library().csolve_gpu_set_rebalance(self._h, C.cast(None, REBALANCE_FN), None))
            return

        def tramp(user, handle, n_idle, n_busy, frame_words):
            self._frame_words = int(frame_words)
            try:
                return int(fn(self, int(n_idle), int(n_busy), int(frame_words)))
            except Exception:            # an exception cannot cross the C frames: abort the search instead
                import traceback
                traceback.print_exc()
                return -1
        self._rebalance_cb = REBALANCE_FN(tramp)        # keep the trampoline alive
        _check(library().csolve_gpu_set_rebalance(self._h, self._rebalance_cb, None))

    def export_frames(self, max_frames):
        """inside the rebalance callback: split up to max_frames frames off this rank's busy warps -> [n, frame_words] int32"""
        buf = np.zeros((max(int(max_frames), 1), self._frame_words), np.int32)
        n = C.c_int32()
        _check(library().csolve_gpu_export_frames(self._h, int(max_frames), buf.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)))
        return buf[:n.value]

    def import_frames(self, frames):
        """inside the rebalance callback: add frames ([n, frame_words] int32) to this rank's pool"""
        frames = np.ascontiguousarray(frames, np.int32).reshape(-1, self._frame_words)
        if frames.shape[0]:
            _check(library().csolve_gpu_import_frames(self._h, frames.ctypes.data_as(C.POINTER(C.c_int32)), frames.shape[0]))
        return frames.shape[0]

    def solve_batch(self, root_domains, order=ORDER_NONE, part_rank=0, part_count=1, split_target=0,
                    max_solutions=0, time_limit_ms=0, slice_ms=0, sample_mod=0, sample_cap=0, sample_failed_keep=1):
        """Search many roots that share this model's network. root_domains: [R, 2*V] int32.
        Returns (SolveResult, per-root solution counts [R], per-root infeasible-at-root flags [R]);
        SolveResult.assignments are (root id, values) pairs."""
        I32P, U8P = C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        if isinstance(order, str):
            order = ORDER_NAMES[order]
        roots = np.ascontiguousarray(root_domains, np.int32).reshape(-1, 2 * self.n_vars)
        n = roots.shape[0]
        counts = np.zeros(n, np.uint32)
        failed = np.zeros(n, np.uint8)
        opt = _SolveOptions(order, part_rank, part_count, split_target, max_solutions, time_limit_ms, slice_ms, 0, 0, 0,
                            int(sample_mod), int(sample_cap), 0, int(sample_failed_keep))
        res = _GpuResult()
        _check(library().csolve_gpu_solve_batch(self._h, C.byref(opt), n, roots.ctypes.data_as(I32P),
                                                counts.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                failed.ctypes.data_as(U8P), C.byref(res)))
        sols = []
        buf = (C.c_int32 * self.n_vars)()
        key = C.c_int32()
        for i in range(res.n_stored):
            _check(library().csolve_gpu_get_solution(self._h, i, buf))
            _check(library().csolve_gpu_get_solution_key(self._h, i, C.byref(key)))
            sols.append((key.value, list(buf)))
        return SolveResult(res, sols), counts, failed

    def close(self):
        if getattr(self, "_h", None):
            library().csolve_gpu_unload(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
