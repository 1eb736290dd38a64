"""Shared test helpers: build/load the oracle, the host contractor harness and (when present)
the compiled reference; small flat-model builder for unit vectors."""
import ctypes as C
import hashlib
import json
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "libcsolve_oracle.so")
HARNESS_SO = os.path.join(ROOT, "tests", "harness", "libhost_contract.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcsolve_ref.so")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "csolve_ref")

I32P = C.POINTER(C.c_int32)
DMIN, DMAX = -2**31, 2**31 - 1


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle():
    src = os.path.join(ROOT, "oracle", "csolve_oracle.c")
    if not _newer(ORACLE_SO, [src, os.path.join(ROOT, "include", "csolve_b200.h")]):
        subprocess.check_call(["gcc", "-std=c99", "-O3", "-Wall", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                               src, "-o", ORACLE_SO])
    return ORACLE_SO


def build_harness():
    csrc = os.path.join(ROOT, "csolve_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "harness", "host_contract.cpp"), os.path.join(csrc, "compile.cpp"),
            os.path.join(csrc, "contract.cuh"), os.path.join(csrc, "device_model.h"), os.path.join(csrc, "compile.hpp")]
    if not _newer(HARNESS_SO, srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                               "-I", csrc, srcs[0], srcs[1], "-o", HARNESS_SO])
    return HARNESS_SO


EMU_DIR = os.path.join(ROOT, "build", "emu")


def emu_kernel_source(text):
    """kernels.cu rewritten for the emulator: shared-memory declarations become per-block storage of the emulator, the
    launch wrappers (<<< >>>, cudaFuncSetAttribute) are cut off; everything else is the product's text."""
    text, n1 = re.subn(r"extern __shared__ __align__\(16\) int smem\[\];", "int *smem = emu_dynamic_smem();", text)
    assert n1 >= 1

    def static_shared(mt):
        indent, ty, decls = mt.group(1), mt.group(2), mt.group(3)
        out = []
        for k, d in enumerate(x.strip() for x in decls.split(",")):
            arr = re.match(r"(\w+)\[(.+)\]$", d)
            if arr:
                out.append("%s (&%s)[%s] = *reinterpret_cast<%s (*)[%s]>(emu_static_shared(__LINE__ * 16 + %d, sizeof(%s) * (%s)));"
                           % (ty, arr.group(1), arr.group(2), ty, arr.group(2), k, ty, arr.group(2)))
            else:
                out.append("%s &%s = *reinterpret_cast<%s *>(emu_static_shared(__LINE__ * 16 + %d, sizeof(%s)));" % (ty, d, ty, k, ty))
        return indent + " ".join(out)

    text, n2 = re.subn(r"(?m)^(\s*)__shared__ ((?:unsigned long long|unsigned int|unsigned|int)) ([^;]+);", static_shared, text)
    assert n2 >= 1 and "__shared__" not in re.sub(r"//.*", "", text), "a __shared__ declaration the emulator does not know"
    text = text.replace("__noinline__", "EMU_NOINLINE")
    # control words every lane of a warp reads for itself to steer a branch with collectives in it: a converged warp
    # issues such a load once on the device; under the emulator the lanes run one after the other between two
    # collectives (lane 0 may already have changed the word), so the read is made explicitly uniform
    text = text.replace("*reinterpret_cast<volatile int *>(&s_blk_hungry)", "EMU_UNIFORM(*reinterpret_cast<volatile int *>(&s_blk_hungry))")
    n0 = text.count("if (*reinterpret_cast<volatile int *>(&ctl->signal) != SIG_RUN) break;")
    text = text.replace("if (*reinterpret_cast<volatile int *>(&ctl->signal) != SIG_RUN) break;",
                        "if (EMU_UNIFORM(*reinterpret_cast<volatile int *>(&ctl->signal)) != SIG_RUN) break;")
    n1 = text.count("*reinterpret_cast<volatile int *>(&ctl->n_stored) > a.max_solutions")
    text = text.replace("*reinterpret_cast<volatile int *>(&ctl->n_stored) > a.max_solutions", "EMU_UNIFORM(*reinterpret_cast<volatile int *>(&ctl->n_stored)) > a.max_solutions")
    assert n0 == 4 and n1 == 4, (n0, n1)       # the poll sections of the four search kernels
    text = re.sub(r'asm volatile\("mov\.u64 %0, %%globaltimer;" : "=l"\((\w+)\)\);', r"\1 = (unsigned long long)clock64();", text)
    assert "asm" not in re.sub(r"//.*", "", text)
    cut = text.index("static cudaError_t ensure_smem(")
    return text[:cut] + "\n}  // namespace csolve_dev\n#endif  // CSOLVE_BJ_UNIT (emulator: the launch wrappers are cut off)\n"


def build_emu(backjump=False):
    """The product's general search kernel compiled for the CPU under the SIMT emulator (tests/harness/simt_emu.h):
    kernels.cu itself, with its dynamic / static __shared__ declarations rewritten to the emulator's per-block storage."""
    csrc = os.path.join(ROOT, "csolve_b200", "csrc")
    os.makedirs(EMU_DIR, exist_ok=True)
    so = os.path.join(EMU_DIR, "libemu_search%s.so" % ("_bj" if backjump else ""))
    srcs = [os.path.join(ROOT, "tests", "harness", "emu_search.cpp"), os.path.join(ROOT, "tests", "harness", "simt_emu.h"),
            os.path.join(csrc, "kernels.cu"), os.path.join(csrc, "kernels.cuh"), os.path.join(csrc, "contract.cuh"),
            os.path.join(csrc, "device_model.h"), os.path.join(csrc, "compile.cpp"), os.path.join(csrc, "compile.hpp"),
            os.path.abspath(__file__)]
    if _newer(so, srcs):
        return so
    text = emu_kernel_source(open(os.path.join(csrc, "kernels.cu")).read())
    inc = os.path.join(EMU_DIR, "kernels_emu.inc")
    if not os.path.exists(inc) or open(inc).read() != text:
        open(inc, "w").write(text)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-fno-omit-frame-pointer", "-w",
                           "-I", os.path.join(ROOT, "include"), "-I", csrc, "-I", EMU_DIR, "-I", cuda_inc,
                           "-I", os.path.join(ROOT, "tests", "harness")] + (["-DCSOLVE_BJ=1"] if backjump else []) +
                          [srcs[0], os.path.join(csrc, "compile.cpp"), "-o", so])
    return so


def ensure_built():
    import csolve_b200
    if not os.path.exists(csolve_b200.library_path()):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "csolve_b200", "csrc")])
    build_oracle()
    build_harness()


class OrcResult(C.Structure):
    _fields_ = [("solutions", C.c_uint64), ("calls", C.c_uint64), ("cuts", C.c_uint64), ("props", C.c_uint64),
                ("best", C.c_int32), ("has_solution", C.c_int32), ("hit_limit", C.c_int32), ("pad", C.c_int32)]


_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        import csolve_b200 as cb
        lib = C.CDLL(build_oracle())
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.POINTER(cb.FlatModel)]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_node.argtypes = [C.c_void_p, I32P, C.c_int, C.c_int32, C.c_int32, I32P]
        lib.orc_leaf_true.argtypes = [C.c_void_p, I32P]
        lib.orc_prop_root.argtypes = [C.c_void_p, I32P, C.c_int32, C.c_int32, I32P]
        lib.orc_eval_root.argtypes = [C.c_void_p, I32P, I32P]
        lib.orc_solve_reference.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.POINTER(OrcResult), I32P]
        lib.orc_solve_tree.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.POINTER(OrcResult), I32P]
        lib.orc_solve_tree_part.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(OrcResult), I32P]
        for f in ("orc_neg", "orc_add", "orc_mul", "orc_min", "orc_max"):
            getattr(lib, f).restype = C.c_int32
        _orc = lib
    return _orc


def p32(a):
    return a.ctypes.data_as(I32P)


class Oracle:
    """The CPU restatement (oracle/csolve_oracle.c) over a flat model."""

    def __init__(self, model):
        # accepts a csolve_b200.Model (kept alive: the flat arrays belong to it) or a bare FlatModel
        self.lib = oracle_lib()
        self._owner = model
        flat = model.flat if hasattr(model, "flat") else model
        self.flat = flat
        self.V = flat.n_vars
        self.h = self.lib.orc_create(C.byref(flat))

    def node(self, dom_in, var, val, best=0):
        dom_in = np.ascontiguousarray(dom_in, np.int32)
        out = np.empty(2 * self.V, np.int32)
        f = self.lib.orc_node(self.h, p32(dom_in), int(var), int(val), int(best), p32(out))
        return out, int(f)

    def leaf_true(self, dom):
        dom = np.ascontiguousarray(dom, np.int32)
        return bool(self.lib.orc_leaf_true(self.h, p32(dom)))

    def prop_root(self, dom_in, vlo, vhi):
        dom_in = np.ascontiguousarray(dom_in, np.int32)
        out = np.empty(2 * self.V, np.int32)
        r = self.lib.orc_prop_root(self.h, p32(dom_in), vlo, vhi, p32(out))
        return out, r

    def eval_root(self, dom_in):
        dom_in = np.ascontiguousarray(dom_in, np.int32)
        out = np.empty(2, np.int32)
        self.lib.orc_eval_root(self.h, p32(dom_in), p32(out))
        return [int(out[0]), int(out[1])]

    def solve_reference(self, order=0, prefer_failing=1, max_calls=0):
        r = OrcResult()
        sol = np.zeros(max(self.V, 1), np.int32)
        self.lib.orc_solve_reference(self.h, order, prefer_failing, max_calls, C.byref(r), p32(sol))
        return r, sol[:self.V].copy()

    def solve_tree(self, order=0, max_calls=0):
        r = OrcResult()
        sol = np.zeros(max(self.V, 1), np.int32)
        self.lib.orc_solve_tree(self.h, order, max_calls, C.byref(r), p32(sol))
        return r, sol[:self.V].copy()

    def solve_tree_part(self, order, part, n_parts, split_level=0, max_calls=0):
        """share `part` of the tree: the nodes of level split_level, dealt round-robin in DFS order, with their subtrees"""
        r = OrcResult()
        sol = np.zeros(max(self.V, 1), np.int32)
        self.lib.orc_solve_tree_part(self.h, order, max_calls, part, n_parts, split_level, C.byref(r), p32(sol))
        return r

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass


_hc = None


def harness_lib():
    global _hc
    if _hc is None:
        import csolve_b200 as cb
        lib = C.CDLL(build_harness())
        lib.hc_load.argtypes = [C.POINTER(cb.FlatModel), C.c_int]
        lib.hc_error.restype = C.c_char_p
        lib.hc_node.argtypes = [I32P, C.c_int, C.c_int32, C.c_int32, I32P]
        lib.hc_leaf_true.argtypes = [I32P]
        lib.hc_node_lov.argtypes = [I32P, C.c_int, C.c_int32, I32P]
        lib.hc_prop_nogood.argtypes = [I32P, C.c_int, I32P, C.c_int, I32P]
        lib.hc_prop_root.argtypes = [I32P, C.c_int32, C.c_int32, I32P]
        lib.hc_eval_root.argtypes = [I32P, I32P]
        for f in ("hc_sneg", "hc_sadd", "hc_smul"):
            getattr(lib, f).restype = C.c_int32
        _hc = lib
    return _hc


def reference_lib():
    """The compiled reference (oracle/_ref), or None when it has not been built (no /root/reference)."""
    if not os.path.exists(REF_SO):
        return None
    lib = C.CDLL(REF_SO)
    import csolve_b200 as cb
    lib.ref_flat.restype = C.POINTER(cb.FlatModel)
    lib.ref_replay.argtypes = [I32P, C.c_int, C.c_int32, C.c_int32, I32P, C.c_void_p]
    lib.ref_get_domains.argtypes = [I32P]
    lib.ref_eval_root.argtypes = [I32P]
    return lib


# ---- hand-built flat models for unit vectors ------------------------------------------------------
OPS = {"EQ": 2, "LT": 3, "NEG": 4, "ADD": 5, "MUL": 6, "NOT": 7, "AND": 8, "OR": 9}


class HandModel:
    """A flat model with one clause OP(v0[, v1]) over free variables (no watch lists)."""

    def __init__(self, op, n_operands, domains):
        import csolve_b200 as cb
        V = len(domains)
        ops, ls, rs = [], [], []
        for i in range(n_operands):
            ops.append(0); ls.append(i); rs.append(-1)
        if n_operands == 1:
            ops.append(OPS[op]); ls.append(0); rs.append(-1)
        else:
            ops.append(OPS[op]); ls.append(0); rs.append(1)
        n = len(ops)
        self._keep = dict(
            op=(C.c_uint8 * n)(*ops), l=(C.c_int32 * n)(*ls), r=(C.c_int32 * n)(*rs),
            cf=(C.c_int32 * 2)(0, n), wp=(C.c_int32 * (V + 1))(*([0] * (V + 1))), wi=(C.c_int32 * 1)(0),
            # root domains are irrelevant for the unit hooks (they take explicit domains); keep them finite
            lo=(C.c_int32 * V)(*([-(1 << 30)] * V)), hi=(C.c_int32 * V)(*([1 << 30] * V)),
            prio=(C.c_int64 * V)(*([0] * V)))
        k = self._keep
        f = cb.FlatModel()
        f.n_vars, f.n_nodes, f.n_clauses, f.n_watch, f.objective, f.obj_var = V, n, 1, 0, 1, -1
        f.node_op = C.cast(k["op"], C.POINTER(C.c_uint8)); f.node_l = C.cast(k["l"], I32P); f.node_r = C.cast(k["r"], I32P)
        f.clause_first = C.cast(k["cf"], I32P); f.watch_ptr = C.cast(k["wp"], I32P); f.watch_idx = C.cast(k["wi"], I32P)
        f.var_lo = C.cast(k["lo"], I32P); f.var_hi = C.cast(k["hi"], I32P)
        f.var_prio = C.cast(k["prio"], C.POINTER(C.c_int64)); f.var_name = None
        self.flat = f


def flat_digest(flat):
    """canonical digest of a flat model (names included)"""
    d = flat.to_dict()
    return hashlib.sha256(json.dumps(d, sort_keys=True).encode()).hexdigest()


def load_vectors():
    with open(os.path.join(GOLDEN, "ref_unit_vectors.json")) as f:
        return json.load(f)


class EmuResult(C.Structure):
    _fields_ = [("solutions", C.c_uint64), ("nodes", C.c_uint64), ("cuts", C.c_uint64), ("props", C.c_uint64),
                ("best", C.c_int32), ("has_solution", C.c_int32), ("n_stored", C.c_int32), ("conflicts", C.c_int32),
                ("conflicts_abandoned", C.c_int32), ("backjumps", C.c_int32), ("claims", C.c_int32), ("slices", C.c_int32),
                ("expand_levels", C.c_int32), ("frontier", C.c_int32), ("restarts", C.c_int32), ("pad", C.c_int32),
                ("switches", C.c_uint64), ("collectives", C.c_uint64), ("site_mismatches", C.c_uint64)]


_emu = {}


def emu_lib(backjump=False):
    """the general search kernel under the SIMT emulator (build_emu); backjump: the back-jumping build"""
    if backjump not in _emu:
        import csolve_b200 as cb
        lib = C.CDLL(build_emu(backjump))
        lib.emu_search.argtypes = [C.POINTER(cb.FlatModel), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(EmuResult), I32P]
        lib.emu_error.restype = C.c_char_p
        assert lib.emu_backjump_build() == (1 if backjump else 0)
        lib.emu_set_lane_step(int(os.environ.get("EMU_LANE_STEP", "1")))     # the emulator's lane schedule (simt_emu.h)
        lib.emu_set_warp_quantum(int(os.environ.get("EMU_WARP_QUANTUM", "64")))   # ... and its warp schedule
        lib.emu_set_workspace_fill.argtypes = [C.c_int32]
        lib.emu_set_workspace_fill(int(os.environ.get("EMU_WORKSPACE_FILL", "0"), 0))  # what stacks / pools hold before a search
        _emu[backjump] = lib
    return _emu[backjump]


def emu_search(model, order=0, learn=False, backjump=False, prefer_failing=False, n_blocks=1, max_solutions=0, general=True,
               slice_clock=0, sink_headroom=0, sink_rows=0, split_target=1, part_rank=0, part_count=1, restart_frequency=0):
    """-> (EmuResult, [assignments]) of one whole search of `model` (csolve_b200.Model) on the emulated kernels.
    general=False: the kernel the product picks for the model (lane-owns-variable, K-per-lane, bit-state, general);
    slice_clock > 0: time slices of that many emulator clock units with k_rebalance between them;
    sink_headroom > 0 (ALL models): a bounded solution buffer drained between slices, every solution returned (at most
    sink_rows); split_target > 1: breadth-first expansion of the root first (k_search<true>); part_rank / part_count: the
    share of the expanded frontier whose path hash maps to this rank (the ALL-mode partition between GPUs)."""
    lib = emu_lib(backjump)
    res = EmuResult()
    cap = sink_rows if sink_headroom > 0 else (max_solutions if max_solutions > 0 else 16)
    buf = np.zeros((cap, model.n_vars + 1), np.int32)
    rc = lib.emu_search(C.byref(model.flat), order, 1 if (learn or backjump) else 0, 1 if prefer_failing else 0, n_blocks,
                        max_solutions, 1 if general else 0, int(slice_clock), int(sink_headroom), int(sink_rows), int(split_target), int(part_rank),
                        int(part_count), int(restart_frequency), C.byref(res),
                        buf.ctypes.data_as(I32P))
    if rc != 0:
        raise RuntimeError("emu_search: %d %s" % (rc, lib.emu_error().decode()))
    n = min(res.n_stored, cap)
    return res, [buf[k, :model.n_vars].tolist() for k in range(n)]


def emu_nogoods(backjump=False):
    """nogoods learned by the last emu_search(learn=True) of that build: [[(var, value), ...], ...]"""
    lib = emu_lib(backjump)
    lib.emu_nogoods.argtypes = [I32P, C.c_int]
    n = lib.emu_nogoods(None, 0)
    buf = np.zeros(max(n, 1), np.int32)
    lib.emu_nogoods(buf.ctypes.data_as(I32P), n)
    out, i = [], 0
    while i < n:
        k = int(buf[i])
        out.append([(int(l) >> 1, int(l) & 1) for l in buf[i + 1:i + 1 + k]])
        i += 1 + k
    return out


def emu_search_batch(model, root_domains, order=0, n_blocks=1, general=False, slice_clock=0, split_target=1):
    """csolve_gpu_solve_batch on the emulated kernels: -> (EmuResult, per-root solution counts, per-root failed flags)"""
    lib = emu_lib(False)
    U32P, U8P = C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
    lib.emu_search_batch.argtypes = [C.POINTER(type(model.flat)), C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int, I32P,
                                     U32P, U8P, C.POINTER(EmuResult)]
    roots = np.ascontiguousarray(root_domains, np.int32).reshape(-1, 2 * model.n_vars)
    n = roots.shape[0]
    counts = np.zeros(n, np.uint32)
    failed = np.zeros(n, np.uint8)
    res = EmuResult()
    rc = lib.emu_search_batch(C.byref(model.flat), order, n_blocks, 1 if general else 0, int(slice_clock), int(split_target), n,
                              roots.ctypes.data_as(I32P), counts.ctypes.data_as(U32P), failed.ctypes.data_as(U8P), C.byref(res))
    if rc != 0:
        raise RuntimeError("emu_search_batch: %d %s" % (rc, lib.emu_error().decode()))
    return res, counts, failed


def emu_sampled_search(model, sample_mod, failed_keep=1, **kw):
    """emu_search on the SAMPLE instances of the kernels: -> (EmuResult, records as GpuProblem.samples() returns them)"""
    lib = emu_lib(False)
    lib.emu_set_sampling.argtypes = [C.c_uint, C.c_uint]
    lib.emu_samples.argtypes = [I32P, C.c_int, C.POINTER(C.c_int)]
    lib.emu_set_sampling(int(sample_mod), int(failed_keep))
    try:
        r, _ = emu_search(model, **kw)
    finally:
        lib.emu_set_sampling(0, 1)
    seen = C.c_int()
    n_words = lib.emu_samples(None, 0, C.byref(seen))
    V = model.n_vars
    W = 4 + 4 * V
    buf = np.zeros(max(n_words, 1), np.int32)
    lib.emu_samples(buf.ctypes.data_as(I32P), n_words, C.byref(seen))
    buf = buf[:n_words].reshape(-1, W)
    return r, dict(flags=buf[:, 0].copy(), var=buf[:, 1].copy(), val=buf[:, 2].copy(), best=buf[:, 3].copy(),
                   parent=buf[:, 4:4 + 2 * V].copy(), child=buf[:, 4 + 2 * V:].copy(), seen=seen.value)


def emu_set_lane_step(step):
    """lane schedule of the emulator (both builds): odd stride, 1 = ascending lanes, 31 = descending (simt_emu.h)"""
    for bj in (False, True):
        emu_lib(bj).emu_set_lane_step(int(step))


def emu_search_comm(model, world, order=0, prefer_failing=False, n_blocks=1, split_target=64, slice_clock=0, general=True,
                    learn=False, backjump=False):
    """ANY / MIN / MAX model searched by `world` emulated GPUs of a csolve_gpu_comm (shared root frontier on rank 0, claims
    with system-scope atomics, incumbents / "found" through the CommBlocks, a rank running dry served by its peers' warps)
    -> (EmuResult of the whole job, witness or None)"""
    lib = emu_lib(backjump)
    lib.emu_search_comm.argtypes = [C.POINTER(type(model.flat)), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int, C.c_int,
                                    C.POINTER(EmuResult), I32P]
    res = EmuResult()
    buf = np.zeros(model.n_vars + 1, np.int32)
    rc = lib.emu_search_comm(C.byref(model.flat), order, 1 if prefer_failing else 0, n_blocks, world, int(split_target),
                             int(slice_clock), 1 if general else 0, 1 if (learn or backjump) else 0, C.byref(res), buf.ctypes.data_as(I32P))
    if rc != 0:
        raise RuntimeError("emu_search_comm: %d %s" % (rc, lib.emu_error().decode()))
    return res, (buf[:model.n_vars].tolist() if res.n_stored else None)


def emu_propagate_batch(model, dom_in, var, val, best=None, general=False, n_blocks=4):
    """csolve_gpu_propagate_batch on the emulated kernels (k_propagate_batch / _lov / _lovk): -> (dom_out [n, 2V], failed [n])"""
    lib = emu_lib(False)
    U8P = C.POINTER(C.c_uint8)
    lib.emu_propagate_batch.argtypes = [C.POINTER(type(model.flat)), C.c_int, C.c_int, C.c_int, I32P, I32P, I32P, I32P, I32P, U8P]
    dom_in = np.ascontiguousarray(dom_in, np.int32).reshape(-1, 2 * model.n_vars)
    n = dom_in.shape[0]
    var = np.ascontiguousarray(var, np.int32)
    val = np.ascontiguousarray(val, np.int32)
    bestp = None if best is None else np.ascontiguousarray(best, np.int32)
    out = np.zeros_like(dom_in)
    failed = np.zeros(n, np.uint8)
    rc = lib.emu_propagate_batch(C.byref(model.flat), 1 if general else 0, n_blocks, n, dom_in.ctypes.data_as(I32P), var.ctypes.data_as(I32P),
                                 val.ctypes.data_as(I32P), None if bestp is None else bestp.ctypes.data_as(I32P),
                                 out.ctypes.data_as(I32P), failed.ctypes.data_as(U8P))
    if rc != 0:
        raise RuntimeError("emu_propagate_batch: %d %s" % (rc, lib.emu_error().decode()))
    return out, failed


def emu_search_exchange(model, world, order=0, n_blocks=1, split_target=64, slice_clock=5000, general=True):
    """ALL model searched by `world` emulated ranks (path-hash shares) with frames shipped from the busiest rank to ranks
    that ran dry at the slice boundaries (k_export_frames / k_import_frames) -> (EmuResult of the whole job, frames moved)"""
    lib = emu_lib(False)
    lib.emu_search_exchange.argtypes = [C.POINTER(type(model.flat)), C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_int,
                                        C.POINTER(EmuResult), I32P]
    res = EmuResult()
    moved = C.c_int32(0)
    rc = lib.emu_search_exchange(C.byref(model.flat), order, n_blocks, world, int(split_target), int(slice_clock), 1 if general else 0,
                                 C.byref(res), C.byref(moved))
    if rc != 0:
        raise RuntimeError("emu_search_exchange: %d %s" % (rc, lib.emu_error().decode()))
    return res, moved.value
