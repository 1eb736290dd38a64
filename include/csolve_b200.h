/*
 * csolve_b200.h -- C ABI of the B200-native csolve search hot path.
 *
 * Everything a host (the reference's C front end, or any FFI) binds is
 * declared here: plain pointers and sizes, no C++/torch types.
 *
 * The seam this library replaces is
 *     void solve(size_t size, struct env_t *env, struct constr_t *constr)
 * (reference src/csolve.h:395, defined src/csolve.c:398, sole caller
 * src/parser.y:86) together with everything solve() drives:
 * propagate_clauses() (src/csolve.h:362, src/propagate.c:488-538), the
 * per-type eval_*()/propagate_*() contractors (src/csolve.h:342-349), the
 * objective_*() incumbent logic (src/csolve.h:380-392) and the branching
 * strategy (src/csolve.h:425-435).
 *
 * The host hands over the root-normalised constraint network as a flat
 * structure-of-arrays CSR (csolve_flat_model). integration/csolve_gpu_shim.c
 * builds it from the reference's env_t/constr_t structures; the built-in front
 * end (csolve_model_parse) builds it from csolve input text.
 */
#ifndef CSOLVE_B200_H
#define CSOLVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSOLVE_B200_ABI_VERSION 5

/* ---- error codes (all entry points return 0 on success) ------------------ */
#define CSOLVE_OK                 0
#define CSOLVE_ERR_INVALID       (-1)  /* bad argument / malformed model */
#define CSOLVE_ERR_SYNTAX        (-2)  /* front end: lexer/parser error (src/csolve.h:513-516) */
#define CSOLVE_ERR_INFEASIBLE    (-3)  /* front end: root phase failed ("INFEASIBLE PROBLEM", src/parser.y:71-73) */
#define CSOLVE_ERR_UNBOUNDED     (-4)  /* front end: "unbounded variable: %s" (src/parser_support.c:249-251) */
#define CSOLVE_ERR_UNSUPPORTED   (-5)  /* construct outside the device path (e.g. all_different nested in an operator) */
#define CSOLVE_ERR_CUDA          (-6)  /* CUDA runtime error; csolve_last_error() has the text */
#define CSOLVE_ERR_NO_DEVICE     (-7)  /* no CUDA device: there is NO CPU fallback */
#define CSOLVE_ERR_CAPACITY      (-8)  /* a device pool (frontier, clause pool) overflowed */

/* ---- node operators: the reference's enum operator_t (src/csolve.h:133-161),
 *      TERM split into variable / constant leaves -------------------------- */
#define CSOLVE_OP_VAR    0   /* TERM with env: l = variable index                       */
#define CSOLVE_OP_CONST  1   /* TERM without env: l = lo, r = hi                        */
#define CSOLVE_OP_EQ     2
#define CSOLVE_OP_LT     3
#define CSOLVE_OP_NEG    4
#define CSOLVE_OP_ADD    5
#define CSOLVE_OP_MUL    6
#define CSOLVE_OP_NOT    7
#define CSOLVE_OP_AND    8
#define CSOLVE_OP_OR     9

/* objective kinds, numbered like enum objective_t (src/csolve.h:241-247) */
#define CSOLVE_OBJ_ANY 0
#define CSOLVE_OBJ_ALL 1
#define CSOLVE_OBJ_MIN 2
#define CSOLVE_OBJ_MAX 3

/* variable orders, numbered like enum order_t (src/csolve.h:249-256) */
#define CSOLVE_ORDER_NONE            0  /* static: by initial priority, ties by index */
#define CSOLVE_ORDER_SMALLEST_DOMAIN 1
#define CSOLVE_ORDER_LARGEST_DOMAIN  2
#define CSOLVE_ORDER_SMALLEST_VALUE  3
#define CSOLVE_ORDER_LARGEST_VALUE   4

/*
 * Flat model: the root-normalised constraint network after the reference's
 * root phase (src/parser.y:55-92), flattened.
 *
 * clause k   = one wand_expr_t that is not itself a WAND (a top-level element
 *              of the root WAND or an element of a nested all_different WAND,
 *              src/parser_support.c:350-361), in traversal order. Elements that
 *              are exactly the constant TERM [1,1] are dropped.
 * nodes      = the clause's expression tree in post-order (children before the
 *              parent, left subtree before right subtree); the nodes of clause
 *              k are clause_first[k] .. clause_first[k+1]-1 and the root is the
 *              last one. node_l/node_r hold GLOBAL node indices for operators.
 * watch list = for variable v, the clauses that mention v, in the order
 *              clauses_init() appends them (src/parser_support.c:339-396);
 *              variables fixed at root have an empty list (parser_support.c:341).
 */
typedef struct csolve_flat_model {
  int32_t n_vars;
  int32_t n_nodes;
  int32_t n_clauses;
  int32_t n_watch;
  int32_t objective;           /* CSOLVE_OBJ_* */
  int32_t obj_var;             /* index of "<obj>" (src/parser.y:120,126) or -1 */
  const uint8_t *node_op;      /* [n_nodes] CSOLVE_OP_* */
  const int32_t *node_l;       /* [n_nodes] */
  const int32_t *node_r;       /* [n_nodes] (-1 for unary operators) */
  const int32_t *clause_first; /* [n_clauses + 1] */
  const int32_t *watch_ptr;    /* [n_vars + 1] */
  const int32_t *watch_idx;    /* [n_watch] clause indices */
  const int32_t *var_lo;       /* [n_vars] root domain (finite: src/parser_support.c:249) */
  const int32_t *var_hi;       /* [n_vars] */
  const int64_t *var_prio;     /* [n_vars] parse-time weights (env_t.prio, src/csolve.h:237) */
  const char *const *var_name; /* [n_vars] env_t.key; may be NULL */
} csolve_flat_model;

/* ---- built-in front end --------------------------------------------------
 * Parses csolve input text (grammar of src/parser.y, tokens of src/lexer.l),
 * runs the root phase (propagate / normalize / propagate, src/parser.y:55-69)
 * and flattens. CPU code; it is the host side of the boundary, not the hot path.
 */
typedef struct csolve_model csolve_model;

typedef struct csolve_front_options {
  int32_t compute_weights;     /* -w (src/main.c:127-130), default 1 */
  int32_t objective_override;  /* -1 keep the file's objective; CSOLVE_OBJ_ANY/ALL replace an ANY/ALL header */
} csolve_front_options;

int csolve_model_parse(const char *text, size_t len, const csolve_front_options *opt,
                       csolve_model **out);
const csolve_flat_model *csolve_model_flat(const csolve_model *m);
void csolve_model_free(csolve_model *m);

/* ---- device ---------------------------------------------------------------
 * Threading: a csolve_gpu_problem belongs to the device it was loaded on and must be used by one host thread at a
 * time; different problems (on the same or on different devices) may be used from different threads concurrently.
 * csolve_gpu_init() only chooses the device later csolve_gpu_load() calls use; csolve_gpu_load_device() names it.
 * Domains handed to csolve_gpu_solve_batch (root_dom) and csolve_gpu_propagate_batch (dom_in, val) must be
 * sub-intervals of the model's own root domains -- the model was compiled against those (constant folding of
 * variables fixed at root, no-saturation proofs of the specialised clause forms, the 32-value window of the
 * value-set kernels); anything else is rejected with CSOLVE_ERR_INVALID. */
typedef struct csolve_gpu_problem csolve_gpu_problem;

typedef struct csolve_gpu_config {
  int32_t device;              /* CUDA device ordinal */
} csolve_gpu_config;

typedef struct csolve_solve_options {
  int32_t order;               /* CSOLVE_ORDER_* (-o, src/main.c:96-100) */
  int32_t part_rank;           /* this process searches root-frontier items i with i % part_count == part_rank */
  int32_t part_count;          /* number of partitions (GPUs); 1 = whole tree */
  int32_t split_target;        /* expand the root until at least this many open sub-trees exist (0 = default) */
  int32_t max_solutions;       /* capacity of the solution buffer (assignments kept for printing); 0 = none */
  int32_t time_limit_ms;       /* -t; 0 = off */
  int32_t slice_ms;            /* length of one persistent-kernel time slice; 0 = default: the whole search in one
                                * slice for ANY / ALL models when there is neither a time limit nor an exchange
                                * callback, 20 ms with either, 2 ms for MIN / MAX (a slice end redistributes the open
                                * frames over all warps, which finds good incumbents sooner) */
  int32_t create_conflicts;    /* -c (src/main.c:57-61): learn decision nogoods from failed nodes (src/conflict.c) into a
                                * device clause pool that is propagated on later nodes. Only 0/1-valued facts can be
                                * recorded, so it matters for SAT-like models; models that run on the specialised
                                * NOT(EQ) kernels never produce a nogood (neither does the reference) and ignore it.
                                * Results are identical with and without. 0 = off */
  int32_t backjump;            /* with create_conflicts, ANY / MIN / MAX models: after a conflict the search unwinds to
                                * 1 + the second-highest level of the nogood's variables, propagates the nogood there and
                                * decides that level again (conflict_backtrack, src/csolve.c:350-364; conflict_update,
                                * src/conflict.c:311-324) instead of popping one level. Per search warp, within the frames
                                * the warp owns. Status, optimum and the validity of the returned assignment do not depend
                                * on it; node counters do. Ignored for ALL models (re-deciding a level would count
                                * solutions twice -- the reference's -c true over-counts them). 0 = chronological */
  int32_t prefer_failing;      /* -f (src/main.c:63-67): break ordering ties by a failure-driven priority that is
                                * shared by all warps and updated during the search (src/csolve.c:459-462,
                                * src/propagate.c:33-54). The tree then depends on timing: ALL counts and optima are
                                * unchanged, node counters are not reproducible. 0 = static parse-time priorities */
  uint32_t sample_mod;         /* parity instrumentation, 0 = off. Every search node -- executed or counted in bulk by a
                                * kernel shortcut -- whose identity hash (parent domains, variable, value) is 0 modulo
                                * sample_mod is recorded: parent domains, decision, incumbent, fail flag, post-fixpoint
                                * domains. The search then runs on instrumented instances of the same kernels (same
                                * source, a template flag). Read the records with csolve_gpu_get_samples(); replayed
                                * through the reference's bind + propagate_clauses (src/csolve.c:448-457) they must
                                * come out identical. Not available together with create_conflicts. */
  int32_t sample_cap;          /* records kept (0 = 65536); further hits are counted and dropped */
  int32_t restart_frequency;   /* -r (src/main.c:109-113): ANY models restart after restart_frequency * luby(k) failed
                                * nodes (src/csolve.c:76-83,264-276; Knuth's Luby sequence). On the device one "failed
                                * node" of the schedule is one per search warp: the whole search goes back to the root
                                * when the warps together have failed restart_frequency * luby(k) * n_warps times, and
                                * the root is expanded again in the order of the failure-driven priorities learned so
                                * far (prefer_failing must be on: without it a restart would rebuild the same tree;
                                * priorities -- and learned nogoods -- survive a restart as in the reference). Counters
                                * include the repeated work, like the reference's CALLS. Single-GPU searches only.
                                * 0 = never restart */
  uint32_t sample_failed_keep; /* of the sampled nodes that failed, keep 1 in sample_failed_keep (0, 1 = all): most nodes fail */
  int32_t reserved2[4];
} csolve_solve_options;

typedef struct csolve_gpu_result {
  uint64_t solutions;          /* ALL: solutions counted; ANY: 0/1; MIN/MAX: improving incumbents seen */
  uint64_t nodes;              /* search nodes = the reference's CALLS (src/csolve.c:65-68) */
  uint64_t cuts;               /* failed nodes = CUTS (src/csolve.c:256) */
  uint64_t props;              /* domain narrowings = PROPS (src/propagate.c:78) */
  uint64_t clause_visits;      /* clause contractions executed */
  int32_t  best;               /* MIN/MAX: best objective value (objective_best(), src/objective.c:133) */
  int32_t  has_solution;       /* 0 => "NO SOLUTION FOUND" (src/csolve.c:184-186) */
  int32_t  timed_out;          /* "TIMEOUT" (src/csolve.c:181-183) */
  int32_t  n_stored;           /* assignments stored in the solution buffer */
  double   kernel_ms;          /* device time of the search kernels (CUDA events) */
  double   expand_ms;          /* device time of root-frontier expansion */
  uint64_t kernel_launches;    /* kernels launched by this call */
  uint64_t conflicts;          /* nogoods learned = CONFL (src/conflict.c:361) */
  uint64_t conflicts_abandoned;/* analyses given up (non-0/1 value involved, too long, pool full) */
  uint64_t restarts;           /* RESTARTS (src/csolve.c:272) */
  uint64_t backjumps;          /* conflicts after which the search dropped more than one level (options.backjump) */
} csolve_gpu_result;

int  csolve_gpu_init(const csolve_gpu_config *cfg);
void csolve_gpu_shutdown(void);
int  csolve_gpu_load(const csolve_flat_model *m, csolve_gpu_problem **out);        /* on the device of the last csolve_gpu_init() */
int  csolve_gpu_load_device(const csolve_flat_model *m, int32_t device, csolve_gpu_problem **out);
void csolve_gpu_unload(csolve_gpu_problem *p);

/* Parity hook: B independent node transitions (src/csolve.c:448-457):
 * for node b: domains dom_in[b] (n_vars pairs lo,hi), decision (var[b] := val[b]),
 * incumbent best[b] (ignored for ANY/ALL) -> post-fixpoint domains dom_out[b] and failed[b].
 * All pointers are HOST pointers; copies are part of the call. */
int csolve_gpu_propagate_batch(csolve_gpu_problem *p, int32_t n_nodes,
                               const int32_t *dom_in, const int32_t *var, const int32_t *val,
                               const int32_t *best, int32_t *dom_out, uint8_t *failed);

/* Replacement for solve(): whole search on the device. */
int csolve_gpu_solve(csolve_gpu_problem *p, const csolve_solve_options *opt, csolve_gpu_result *res);

/* One process per GPU: exchange between the ranks, called on the host once per time slice of
 * csolve_gpu_solve() until it returns non-zero ("every rank is done"). It replaces the reference's shared page
 * (struct shared_t: objective_best / solutions, src/csolve.h:259-266):
 *   *best   in: this rank's incumbent        out: the best incumbent of all ranks (MIN/MAX models)
 *   *found  in: 1 if this rank has a solution (ANY models)   out: 1 if any rank has one -> this rank stops
 *   local_done: this rank has no work left (it keeps calling so that the collectives stay matched)
 * The callback typically wraps one NCCL all-reduce (csolve_b200/distributed.py). */
typedef int (*csolve_exchange_fn)(void *user, int32_t *best, int32_t *found, int32_t local_done);
int csolve_gpu_set_exchange(csolve_gpu_problem *p, csolve_exchange_fn fn, void *user);

/* Frontier rebalancing between the ranks (the reference's worker_spawn hands half of a worker's open values to a new
 * process whenever a slot is free, src/csolve.c:121-149; across GPUs the same bisection travels as frames).
 * The callback runs on the host once per time slice, right after the exchange callback and only while some rank still
 * has work, with every search warp parked:
 *   n_idle / n_busy  warps of this rank without / with work        frame_words  32-bit words of one frame
 * Inside it the two calls below move frames (HOST buffers, n * frame_words words):
 *   csolve_gpu_export_frames  splits up to max_frames frames off this rank's busy warps (shallowest untried work
 *                             first: whole frames below a warp's top frame, the upper half of a top frame)
 *   csolve_gpu_import_frames  adds frames to this rank's pool; its idle warps pick them up in the next slice
 * It returns the number of frames it imported (>= 0) or a negative value to abort the search. Typically it wraps one
 * all-gather of the ranks' (idle, busy) counts and one of the exported frames (csolve_b200/distributed.py). */
typedef int (*csolve_rebalance_fn)(void *user, csolve_gpu_problem *p, int32_t n_idle, int32_t n_busy, int32_t frame_words);
int csolve_gpu_set_rebalance(csolve_gpu_problem *p, csolve_rebalance_fn fn, void *user);
int csolve_gpu_export_frames(csolve_gpu_problem *p, int32_t max_frames, int32_t *frames, int32_t *n_out);
int csolve_gpu_import_frames(csolve_gpu_problem *p, const int32_t *frames, int32_t n_frames);

/* ---- several GPUs on ONE search tree ------------------------------------------------------------------
 * The reference's parallel mode lives inside solve(): `-j N` forks workers that split a variable's interval
 * (worker_spawn, src/csolve.c:105-152) and meet in a shared page (struct shared_t, src/csolve.h:259-266:
 * objective_best, solutions). Here the ranks of a csolve_gpu_comm (one GPU each, at most 8) do the same over
 * NVLink peer memory, with no host in the loop while the search runs:
 *   - ANY / MIN / MAX models: rank 0 expands the root breadth-first and leaves the frontier in its segment; EVERY rank
 *     claims frames of that one frontier with a system-scope atomicAdd on rank 0's counter (the sub-trees the value
 *     order prefers are searched first, by everybody); a rank that is running out of work has its donation ring served
 *     by the busy warps of its peers over NVLink, exactly like the tickets of its own waiting warps;
 *   - ALL models: every rank expands the root for itself and searches the frames whose path hash maps to it; nothing is
 *     exchanged while the search runs (measured on 8 x B200: every form of sharing cost more than the 2 % the hash
 *     partition is off by, profiles/r2_scaling.md);
 *   - an improving incumbent (MIN / MAX) and "a solution exists" (ANY) are stored straight into every peer's
 *     control block by the warp that found them (8 bytes, epoch-tagged, system-scope atomicMin / atomicMax);
 *   - each rank returns its own counters; their sum / the best incumbent is the result (csolve_gpu_group_solve does
 *     that for the in-process case, csolve_b200/distributed.py with one NCCL all-reduce for one process per GPU).
 * csolve_gpu_solve_comm is COLLECTIVE: every rank calls it, the same number of times, on the same model.
 *
 * One process per GPU (torchrun / MPI): create, exchange the 64-byte handles (any all-gather), connect:
 *     csolve_gpu_comm_create(device, rank, world, 0, &c); csolve_gpu_comm_handle(c, mine);
 *     <all-gather mine -> all>;                            csolve_gpu_comm_connect(c, all);
 * One process, several GPUs: csolve_gpu_group_* below (host threads, direct peer access) -- what the drop-in
 * solve() uses for `-j N`. frontier_bytes: capacity of rank 0's frontier buffer (0 = 256 MiB); a frontier that does
 * not fit falls back to "every rank expands, frames are partitioned by path hash" (what ALL models always do). */
#define CSOLVE_COMM_HANDLE_BYTES 64
typedef struct csolve_gpu_comm csolve_gpu_comm;
int  csolve_gpu_comm_create(int32_t device, int32_t rank, int32_t world, size_t frontier_bytes, csolve_gpu_comm **out);
int  csolve_gpu_comm_handle(csolve_gpu_comm *c, void *handle /* CSOLVE_COMM_HANDLE_BYTES */);
int  csolve_gpu_comm_connect(csolve_gpu_comm *c, const void *handles /* world x CSOLVE_COMM_HANDLE_BYTES, rank order */);
int  csolve_gpu_comm_connect_local(csolve_gpu_comm **comms, int32_t world);   /* all ranks in this process */
void csolve_gpu_comm_destroy(csolve_gpu_comm *c);
int  csolve_gpu_solve_comm(csolve_gpu_problem *p, csolve_gpu_comm *c, const csolve_solve_options *opt, csolve_gpu_result *res);

typedef struct csolve_gpu_group csolve_gpu_group;
typedef void (*csolve_solution_fn)(void *user, const int32_t *values, int32_t n, int32_t stride);   /* see csolve_gpu_set_solution_sink */
int  csolve_gpu_device_count(int32_t *n);
int  csolve_gpu_group_create(int32_t n_devices, const int32_t *devices /* NULL: 0..n-1 */, size_t frontier_bytes, csolve_gpu_group **out);
int  csolve_gpu_group_load(csolve_gpu_group *g, const csolve_flat_model *m);           /* the model on every device */
int  csolve_gpu_group_solve(csolve_gpu_group *g, const csolve_solve_options *opt, csolve_gpu_result *res /* whole job */,
                            csolve_gpu_result *per_device /* [n_devices] or NULL */);
int  csolve_gpu_group_get_solution(csolve_gpu_group *g, int32_t i, int32_t *values, int32_t *key /* or NULL */);
int  csolve_gpu_group_set_solution_sink(csolve_gpu_group *g, csolve_solution_fn fn, void *user);  /* calls are serialised */
void csolve_gpu_group_destroy(csolve_gpu_group *g);

/* Batched roots (BASELINE config 2: many instances that share one constraint network and differ only in
 * their root domains, e.g. 10 000 sudokus = the 27 all_different groups + per-instance clue domains).
 * root_dom: n_roots x (2 * n_vars) lo,hi pairs (HOST). The root phase (propagation of every variable's
 * clauses to fixpoint, what src/propagate.c:474-485 does once per process) runs on the device for all
 * roots at once, then all roots are searched together. Requires an ALL model.
 * root_solutions[r] = number of solutions of root r; root_failed[r] = 1 if root r is infeasible at root
 * ("INFEASIBLE PROBLEM"). Stored assignments carry the root id as their key. */
int csolve_gpu_solve_batch(csolve_gpu_problem *p, const csolve_solve_options *opt, int32_t n_roots,
                           const int32_t *root_dom, uint32_t *root_solutions, uint8_t *root_failed,
                           csolve_gpu_result *res);

/* Every solution, however many: the reference prints each accepted leaf as it finds it (src/csolve.c:228-236). With a
 * sink installed the solution buffer (max_solutions entries, at least 2^20 then) is drained between time slices: a
 * slice ends as soon as the buffer is nearly full, the host copies the assignments out, hands them to the sink
 * (n assignments of n_vars values each, `stride` int32 apart; the value after an assignment is its key) and the search
 * goes on -- nothing is dropped, nothing is kept for csolve_gpu_get_solution(). If a buffer overflows all the same
 * (a sink that cannot keep up is not the reason: the search waits for it) the call fails with CSOLVE_ERR_CAPACITY
 * rather than lose a solution silently. ALL models; MIN / MAX keep their chain of incumbents in the buffer as before. */
int csolve_gpu_set_solution_sink(csolve_gpu_problem *p, csolve_solution_fn fn, void *user);

/* Learned nogoods of the last csolve_gpu_solve() with create_conflicts: copies up to cap_lits literal codes
 * (var << 1 | value) into lits and the start offset of every nogood into starts[0..n] (starts[n] = total);
 * returns the number of nogoods n in *n_out. For inspection and tests. */
int csolve_gpu_get_nogoods(csolve_gpu_problem *p, int32_t *lits, int32_t cap_lits, int32_t *starts, int32_t cap_ng,
                           int32_t *n_out);

/* Records written by the last csolve_gpu_solve / csolve_gpu_solve_batch with sample_mod > 0. One record is
 * 4 + 4 * n_vars int32: flags (CSOLVE_SAMPLE_*), variable, value, incumbent, parent domains (n_vars lo,hi pairs: the
 * state before the assignment), post-fixpoint domains (n_vars pairs; meaningless when the node failed).
 * Copies up to cap_records records to `records` (HOST); *n_out = records copied, *n_seen = hits including dropped. */
#define CSOLVE_SAMPLE_FAILED  1   /* the node failed (PROP_ERROR) */
#define CSOLVE_SAMPLE_COUNTED 2   /* counted by a bulk shortcut of the kernel, not executed */
#define CSOLVE_SAMPLE_LEAF    4   /* an accepted leaf: every variable a value and every clause true */
int csolve_gpu_get_samples(csolve_gpu_problem *p, int32_t *records, int32_t cap_records, int32_t *n_out, int32_t *n_seen);

/* key of stored assignment i: MIN/MAX objective value, or the root id for batched roots */
int csolve_gpu_get_solution_key(csolve_gpu_problem *p, int32_t i, int32_t *key);

/* copy stored assignment i (n_vars values, variable order of the model) */
int csolve_gpu_get_solution(csolve_gpu_problem *p, int32_t i, int32_t *values);

const char *csolve_last_error(void);
int csolve_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CSOLVE_B200_H */
