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
/* The drop-in: same signature and side effects as src/csolve.c:398. */
void solve(size_t size, struct env_t *env, struct constr_t *constr) {
  csolve_flat_model m;
  int rc = csolve_flatten_reference(size, env, constr, &m);
  if (rc != CSOLVE_OK) {
    print_fatal("cannot flatten model for the GPU path: %d", rc);
  }
  csolve_gpu_config cfg = { .device = 0 };
  if ((rc = csolve_gpu_init(&cfg)) != CSOLVE_OK) {
    print_fatal("%s", csolve_last_error());
  }
  csolve_gpu_problem *p = NULL;
  if ((rc = csolve_gpu_load(&m, &p)) != CSOLVE_OK) {
    print_fatal("%s", csolve_last_error());
  }
  const char *maxsol = getenv("CSOLVE_GPU_MAX_PRINT");
  csolve_solve_options opt;
  memset(&opt, 0, sizeof(opt));
  opt.order = CSOLVE_ORDER_NONE;
  opt.part_count = 1;
  opt.max_solutions = maxsol ? atoi(maxsol) : (1 << 20);
  csolve_gpu_result res;
  if ((rc = csolve_gpu_solve(p, &opt, &res)) != CSOLVE_OK) {
    print_fatal("%s", csolve_last_error());
  }
  /* print the stored assignments in the reference's format; for MIN/MAX the
   * stored sequence is the chain of improving incumbents */
  int32_t *vals = malloc(size * sizeof(int32_t));
  for (int32_t i = 0; i < res.n_stored; i++) {
    csolve_gpu_get_solution(p, i, vals);
    fprintf(stdout, "#1: SOLUTION: ");
    for (size_t v = 0; v < size; v++) {
      fprintf(stdout, "%s = %d, ", env[v].key, vals[v]);
    }
    int32_t best = 0;
    if (m.obj_var >= 0) best = vals[m.obj_var];
    fprintf(stdout, "BEST: %d\n", best);
  }
  free(vals);
  shared()->solutions = res.solutions;
  if (m.obj_var >= 0 && res.has_solution) {
    shared()->objective_best = res.best;
  }
  fprintf(stdout, "#1: CALLS: %lu, CUTS: %lu, PROPS: %lu, SOLUTIONS: %lu\n",
          (unsigned long)res.nodes, (unsigned long)res.cuts, (unsigned long)res.props,
          (unsigned long)res.solutions);
  if (res.timed_out) {
    fprintf(stdout, "TIMEOUT\n");
  }
  if (!res.has_solution) {
    fprintf(stdout, "NO SOLUTION FOUND\n");
  }
  csolve_gpu_unload(p);
  csolve_flat_model_release(&m);
}
#endif
