/*
 * csolve_gpu_shim.c -- the reference-side binding of the drop-in.
 *
 * This is the file a csolve maintainer adds to the reference tree. It is
 * compiled against the reference's own csolve.h (not shipped here) and
 *   1. flattens the root-normalised env_t[] / constr_t network that
 *      src/parser.y:55-85 hands to solve() into a csolve_flat_model, and
 *   2. provides a replacement solve() (src/csolve.h:395) that runs the search
 *      on the GPU through the C ABI of include/csolve_b200.h and prints
 *      solutions / the final status in the reference's format
 *      (src/csolve.c:175-187,233-234, src/print.c:57-70).
 *
 * Build: see INTEGRATION.md (compile src/csolve.c with -Dsolve=solve_cpu so the
 * reference's own solve() stays available under another name).
 */
#include "csolve.h"
#include "csolve_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

struct flat_builder {
  size_t n_nodes, cap_nodes;
  uint8_t *op; int32_t *l, *r;
  size_t n_clauses, cap_clauses;
  int32_t *clause_first;
  const struct wand_expr_t **clause_ptr;  /* reference clause of flat clause k */
  struct env_t *env; size_t size;
  int unsupported;
};

static void fb_node(struct flat_builder *b, uint8_t op, int32_t l, int32_t r) {
  if (b->n_nodes == b->cap_nodes) {
    b->cap_nodes = b->cap_nodes ? 2 * b->cap_nodes : 1024;
    b->op = realloc(b->op, b->cap_nodes);
    b->l = realloc(b->l, b->cap_nodes * sizeof(int32_t));
    b->r = realloc(b->r, b->cap_nodes * sizeof(int32_t));
  }
  b->op[b->n_nodes] = op; b->l[b->n_nodes] = l; b->r[b->n_nodes] = r;
  b->n_nodes++;
}

/* emit the tree below c in post-order; returns the node index of c */
static int32_t fb_tree(struct flat_builder *b, const struct constr_t *c) {
  if (IS_TYPE(TERM, c)) {
    if (c->constr.term.env != NULL) {
      fb_node(b, CSOLVE_OP_VAR, (int32_t)(c->constr.term.env - b->env), -1);
    } else {
      fb_node(b, CSOLVE_OP_CONST, c->constr.term.val.lo, c->constr.term.val.hi);
    }
    return (int32_t)b->n_nodes - 1;
  }
  uint8_t op;
  switch (c->type->op) {
  case OP_EQ:  op = CSOLVE_OP_EQ; break;
  case OP_LT:  op = CSOLVE_OP_LT; break;
  case OP_NEG: op = CSOLVE_OP_NEG; break;
  case OP_ADD: op = CSOLVE_OP_ADD; break;
  case OP_MUL: op = CSOLVE_OP_MUL; break;
  case OP_NOT: op = CSOLVE_OP_NOT; break;
  case OP_AND: op = CSOLVE_OP_AND; break;
  case OP_OR:  op = CSOLVE_OP_OR; break;
  default:     /* WAND below an operator, CONFL at root: not on the device path */
    b->unsupported = 1;
    fb_node(b, CSOLVE_OP_CONST, 1, 1);
    return (int32_t)b->n_nodes - 1;
  }
  int32_t l = fb_tree(b, c->constr.expr.l);
  int32_t r = -1;
  if (op != CSOLVE_OP_NEG && op != CSOLVE_OP_NOT) {
    r = fb_tree(b, c->constr.expr.r);
  }
  fb_node(b, op, l, r);
  return (int32_t)b->n_nodes - 1;
}

static void fb_clause(struct flat_builder *b, const struct wand_expr_t *clause) {
  const struct constr_t *c = clause->constr;
  if (IS_TYPE(TERM, c) && c->constr.term.env == NULL &&
      c->constr.term.val.lo == 1 && c->constr.term.val.hi == 1) {
    return; /* folded to the constant 1: nothing to propagate or check */
  }
  if (b->n_clauses + 1 >= b->cap_clauses) {
    b->cap_clauses = b->cap_clauses ? 2 * b->cap_clauses : 256;
    b->clause_first = realloc(b->clause_first, (b->cap_clauses + 1) * sizeof(int32_t));
    b->clause_ptr = realloc(b->clause_ptr, b->cap_clauses * sizeof(*b->clause_ptr));
  }
  b->clause_first[b->n_clauses] = (int32_t)b->n_nodes;
  b->clause_ptr[b->n_clauses] = clause;
  fb_tree(b, c);
  b->n_clauses++;
  b->clause_first[b->n_clauses] = (int32_t)b->n_nodes;
}

/* clauses are the non-WAND elements, exactly as clauses_init_wand() assigns
 * them (src/parser_support.c:350-361) */
static void fb_wand(struct flat_builder *b, const struct constr_t *w) {
  for (size_t i = 0; i < w->constr.wand.length; i++) {
    const struct wand_expr_t *e = &w->constr.wand.elems[i];
    if (IS_TYPE(WAND, e->constr)) {
      fb_wand(b, e->constr);
    } else {
      fb_clause(b, e);
    }
  }
}

void csolve_flat_model_release(csolve_flat_model *m) {
  free((void *)m->node_op); free((void *)m->node_l); free((void *)m->node_r);
  free((void *)m->clause_first); free((void *)m->watch_ptr); free((void *)m->watch_idx);
  free((void *)m->var_lo); free((void *)m->var_hi); free((void *)m->var_prio);
  free((void *)m->var_name);
  memset(m, 0, sizeof(*m));
}

/* Flatten what solve() receives. env/constr are borrowed and only read. */
int csolve_flatten_reference(size_t size, struct env_t *env, struct constr_t *constr,
                             csolve_flat_model *out) {
  struct flat_builder b;
  memset(&b, 0, sizeof(b));
  b.env = env; b.size = size;
  if (!IS_TYPE(WAND, constr)) {
    return CSOLVE_ERR_INVALID;
  }
  b.clause_first = malloc(sizeof(int32_t));
  b.clause_first[0] = 0;
  fb_wand(&b, constr);

  memset(out, 0, sizeof(*out));
  out->n_vars = (int32_t)size;
  out->n_nodes = (int32_t)b.n_nodes;
  out->n_clauses = (int32_t)b.n_clauses;
  out->objective = (int32_t)objective();
  out->obj_var = -1;
  if (objective_val() != NULL && objective_val()->constr.term.env != NULL) {
    out->obj_var = (int32_t)(objective_val()->constr.term.env - env);
  }

  int32_t *lo = malloc(size * sizeof(int32_t)), *hi = malloc(size * sizeof(int32_t));
  int64_t *prio = malloc(size * sizeof(int64_t));
  const char **name = malloc(size * sizeof(char *));
  int32_t *wptr = malloc((size + 1) * sizeof(int32_t));
  size_t w = 0;
  for (size_t i = 0; i < size; i++) {
    lo[i] = env[i].val->constr.term.val.lo;
    hi[i] = env[i].val->constr.term.val.hi;
    prio[i] = env[i].prio;
    name[i] = env[i].key;
    wptr[i] = (int32_t)w;
    w += env[i].clauses.length;
  }
  wptr[size] = (int32_t)w;
  int32_t *widx = malloc((w ? w : 1) * sizeof(int32_t));
  for (size_t i = 0, k = 0; i < size; i++) {
    for (size_t j = 0; j < env[i].clauses.length; j++, k++) {
      const struct wand_expr_t *c = env[i].clauses.elems[j];
      int32_t id = -1;
      for (size_t q = 0; q < b.n_clauses; q++) {
        if (b.clause_ptr[q] == c) { id = (int32_t)q; break; }
      }
      if (id < 0) { b.unsupported = 1; id = 0; }
      widx[k] = id;
    }
  }
  out->n_watch = (int32_t)w;
  out->node_op = b.op; out->node_l = b.l; out->node_r = b.r;
  out->clause_first = b.clause_first;
  out->watch_ptr = wptr; out->watch_idx = widx;
  out->var_lo = lo; out->var_hi = hi; out->var_prio = prio; out->var_name = name;
  free(b.clause_ptr);
  if (b.unsupported) {
    csolve_flat_model_release(out);
    return CSOLVE_ERR_UNSUPPORTED;
  }
  return CSOLVE_OK;
}

#ifndef CSOLVE_SHIM_NO_SOLVE
/* ---- options --------------------------------------------------------------------------------------------
 * Everything src/main.c parsed is read back through the reference's own public getters where they exist:
 *   -c  strategy_create_conflicts()      -f  strategy_prefer_failing()      -r  strategy_restart_frequency()
 * -o has no getter (strategy.c keeps `_order` static); it is recovered from the behaviour of the public
 * variable heap (strategy_var_order_init / _push / _pop, src/csolve.h:430-436) on synthetic variables: with two
 * entries, pop returns the one pushed first unless strategy_var_cmp() strictly prefers the other.
 * -t and -j live in statics of csolve.c; integration/csolve_accessors.c (a translation unit that includes the
 * unmodified csolve.c and adds two getters) exposes them. Without it: CSOLVE_GPU_TIME_MAX / CSOLVE_GPU_DEVICES. */
uint32_t csolve_shim_time_max(void) __attribute__((weak));
uint32_t csolve_shim_workers_max(void) __attribute__((weak));

static struct env_t *probe_pop(struct env_t *pair, struct constr_t *terms, int lo0, int hi0, int lo1, int hi1) {
  memset(pair, 0, 2 * sizeof(*pair));
  memset(terms, 0, 2 * sizeof(*terms));
  terms[0].constr.term.val = INTERVAL(lo0, hi0);
  terms[1].constr.term.val = INTERVAL(lo1, hi1);
  pair[0].key = "a"; pair[0].val = &terms[0];
  pair[1].key = "b"; pair[1].val = &terms[1];
  strategy_var_order_init(2, pair);          /* pushes pair[0], then pair[1] */
  struct env_t *first = strategy_var_order_pop();
  strategy_var_order_free();
  return first;
}

/* which of two variables the configured order prefers: 0 / 1, or -1 for a tie (priorities are equal) */
static int probe_prefers(int lo0, int hi0, int lo1, int hi1) {
  struct env_t pair[2];
  struct constr_t terms[2];
  const int ab = probe_pop(pair, terms, lo0, hi0, lo1, hi1) == &pair[0] ? 0 : 1;
  const int ba = probe_pop(pair, terms, lo1, hi1, lo0, hi0) == &pair[0] ? 1 : 0;
  return ab == ba ? ab : -1;
}

static int probe_order(void) {
  /* A = [0,10] vs B = [2,5]: smallest-domain prefers B, everything else A; C = [0,3] vs D = [5,6]: largest-domain and
   * smallest-value prefer C; E = [0,9] vs F = [1,20]: smallest-value prefers E, largest-value F. -o none ties. */
  const int ab = probe_prefers(0, 10, 2, 5), cd = probe_prefers(0, 3, 5, 6), ef = probe_prefers(0, 9, 1, 20);
  if (ab < 0 && cd < 0 && ef < 0) return CSOLVE_ORDER_NONE;
  if (ab == 1) return CSOLVE_ORDER_SMALLEST_DOMAIN;
  if (cd == 1) return CSOLVE_ORDER_LARGEST_VALUE;
  return ef == 0 ? CSOLVE_ORDER_SMALLEST_VALUE : CSOLVE_ORDER_LARGEST_DOMAIN;
}

/* ---- printing (src/csolve.c:228-236, src/print.c:57-70) ------------------------------------------------------ */
struct sink_state { size_t size; struct env_t *env; int quiet; uint64_t printed; };

static void print_assignment(const struct sink_state *st, const int32_t *vals, int32_t best) {
  fprintf(stdout, "#1: SOLUTION: ");
  for (size_t v = 0; v < st->size; v++) {
    fprintf(stdout, "%s = %d, ", st->env[v].key, vals[v]);
  }
  fprintf(stdout, "BEST: %d\n", best);
}

static void solution_sink(void *user, const int32_t *values, int32_t n, int32_t stride) {
  struct sink_state *st = user;
  st->printed += (uint64_t)n;
  if (st->quiet) return;
  for (int32_t i = 0; i < n; i++) print_assignment(st, values + (size_t)i * stride, 0);
}

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

/* The drop-in: same signature and side effects as src/csolve.c:398. */
void solve(size_t size, struct env_t *env, struct constr_t *constr) {
  csolve_flat_model m;
  const int timing = getenv("CSOLVE_GPU_TIMING") != NULL;        /* where a cold process spends its time (stderr) */
  const double t_begin = now_s();
  int rc = csolve_flatten_reference(size, env, constr, &m);
  if (rc != CSOLVE_OK) {
    print_fatal("cannot flatten model for the GPU path: %d", rc);
  }
  /* the reference's heap of the real variables is not needed on this path; it is rebuilt below for the caller */
  strategy_var_order_free();
  csolve_solve_options opt;
  memset(&opt, 0, sizeof(opt));
  opt.order = probe_order();
  opt.part_count = 1;
  opt.create_conflicts = strategy_create_conflicts() ? 1 : 0;
  /* the reference back-jumps whenever it learns (src/csolve.c:350-364, 463-466); on the device that is an option of its
   * own (csolve_solve_options.backjump), opted into here with CSOLVE_GPU_BACKJUMP=1 until it has been measured */
  opt.backjump = opt.create_conflicts && getenv("CSOLVE_GPU_BACKJUMP") != NULL && atoi(getenv("CSOLVE_GPU_BACKJUMP")) != 0;
  opt.prefer_failing = strategy_prefer_failing() ? 1 : 0;
  opt.restart_frequency = (int32_t)(strategy_restart_frequency() > 0x7fffffffu ? 0x7fffffff : strategy_restart_frequency());
  uint32_t time_max = csolve_shim_time_max ? csolve_shim_time_max() : 0;
  if (getenv("CSOLVE_GPU_TIME_MAX")) time_max = (uint32_t)strtoul(getenv("CSOLVE_GPU_TIME_MAX"), NULL, 10);
  if (time_max > 0 && time_max < 2000000u) opt.time_limit_ms = (int32_t)(time_max * 1000u);   /* UINT32_MAX: no limit */
  int32_t n_dev = csolve_shim_workers_max ? (int32_t)csolve_shim_workers_max() : 1;            /* -j N: N GPUs */
  if (getenv("CSOLVE_GPU_DEVICES")) n_dev = atoi(getenv("CSOLVE_GPU_DEVICES"));
  int32_t have = 0;
  const double t_flat = now_s();
  if ((rc = csolve_gpu_device_count(&have)) != CSOLVE_OK) {
    print_fatal("%s", csolve_last_error());
  }
  const double t_count = now_s();
  if (n_dev < 1) n_dev = 1;
  if (n_dev > have) n_dev = have;
  if (n_dev > 8) n_dev = 8;

  /* ALL: every solution is printed as the reference does, streamed out of the device between time slices.
   * CSOLVE_GPU_COUNT_ONLY=1 keeps the count and the statistics but prints no assignment (then the kernels count the
   * last level instead of producing it). MIN / MAX: the chain of improving incumbents; ANY: the solution. */
  struct sink_state st = { size, env, 0, 0 };
  const int count_only = getenv("CSOLVE_GPU_COUNT_ONLY") != NULL && atoi(getenv("CSOLVE_GPU_COUNT_ONLY")) != 0;
  const int stream = m.objective == CSOLVE_OBJ_ALL && !count_only;

  csolve_gpu_result res;
  csolve_gpu_group *g = NULL;
  double t_create = 0, t_load = 0;
  if ((rc = csolve_gpu_group_create(n_dev, NULL, 0, &g)) != CSOLVE_OK ||
      (t_create = now_s(), rc = csolve_gpu_group_load(g, &m)) != CSOLVE_OK ||
      (t_load = now_s(), stream && (rc = csolve_gpu_group_set_solution_sink(g, solution_sink, &st)) != CSOLVE_OK) ||
      (rc = csolve_gpu_group_solve(g, &opt, &res, NULL)) != CSOLVE_OK) {
    print_fatal("%s", csolve_last_error());
  }
  if (timing) {
    fprintf(stderr, "[csolve_gpu] flatten + options %.1f ms, CUDA init (device count) %.1f ms, contexts + segments %.1f ms, "
                    "model upload + workspace %.1f ms, search %.1f ms (device: expand %.2f + search %.2f ms)\n",
            1e3 * (t_flat - t_begin), 1e3 * (t_count - t_flat), 1e3 * (t_create - t_count), 1e3 * (t_load - t_create),
            1e3 * (now_s() - t_load), res.expand_ms, res.kernel_ms);
  }
  int32_t *vals = malloc((size ? size : 1) * sizeof(int32_t));
  for (int32_t i = 0; i < res.n_stored; i++) {
    int32_t key = 0;
    csolve_gpu_group_get_solution(g, i, vals, &key);
    print_assignment(&st, vals, m.obj_var >= 0 ? vals[m.obj_var] : 0);
  }
  free(vals);
  shared()->solutions = res.solutions;
  if (m.obj_var >= 0 && res.has_solution) {
    shared()->objective_best = res.best;
  }
  fprintf(stdout, "#1: CALLS: %lu, CUTS: %lu, PROPS: %lu, CONFL: %lu, SOLUTIONS: %lu\n",
          (unsigned long)res.nodes, (unsigned long)res.cuts, (unsigned long)res.props,
          (unsigned long)res.conflicts, (unsigned long)res.solutions);
  if (res.timed_out) {
    shared()->timeout = true;
    fprintf(stdout, "TIMEOUT\n");
  }
  if (!res.has_solution) {
    fprintf(stdout, "NO SOLUTION FOUND\n");
  }
  csolve_gpu_group_destroy(g);
  csolve_flat_model_release(&m);
  strategy_var_order_init(size, env);      /* the caller's env_free() / later code finds the heap as it left it */
}
#endif
