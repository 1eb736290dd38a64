/*
 * csolve_oracle.c -- CPU restatement of the reference's search hot path.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use this file; the product
 * (csolve_b200/) never links, imports or calls it.
 *
 * It restates, in plain C over the flat model of include/csolve_b200.h, the
 * algorithm of the reference (jeuneS2/csolve, /root/reference/src):
 *   arithmetic        arith.c:27-85
 *   evaluation        eval.c:27-245
 *   contraction       propagate.c:57-376   (recursive, as the reference)
 *   fixpoint engine   propagate.c:488-538  (recursion + prop_tag, without the in-search
 *                                           normalise/patch loop of :521-535, which does
 *                                           not change results)
 *   trail             util.c:137-173
 *   variable order    strategy.c:79-246
 *   objective         objective.c:35-136
 *   search            csolve.c:279-338 (step_*), :398-476 (solve) with -c false -r 0 -j 1
 *
 * Pinning: tests/test_oracle_*.py check it against the value tables of the reference's
 * own unit tests (test/test_arith.c, test/test_eval.c, test/test_propagate.c,
 * test/test_csolve.c, test/test_objective.c) and against the compiled reference
 * (oracle/_ref): identical CALLS/CUTS/PROPS/solutions on the example instances and
 * identical node transitions on sampled nodes (fixtures under tests/golden/).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "csolve_b200.h"

#define DMIN INT32_MIN
#define DMAX INT32_MAX
#define PROP_ERROR (-1)

typedef struct { int32_t lo, hi; } val_t;

typedef struct { int32_t var; val_t old; } trail_t;

typedef struct orc {
  int V, N, C, W, objective, obj_var;
  uint8_t *op; int32_t *l, *r, *cfirst, *wptr, *widx;
  int64_t *prio0;
  val_t *root;
  /* search state */
  val_t *dom;
  int64_t *prio;
  uint64_t *ctag; uint64_t tag;
  trail_t *trail; size_t trail_len, trail_cap;
  uint64_t props;
  /* variable heap (strategy.c:124-246) */
  int *heap; int heap_size; int *hpos;
  int order, prefer_failing;
  int32_t best;
} orc;

/* ---- arith.c ------------------------------------------------------------------- */
int32_t orc_neg(int32_t a) {
  if (a == DMIN) return DMAX;
  if (a == DMAX) return DMIN;
  return -a;
}
int32_t orc_add(int32_t a, int32_t b) {
  if (a == DMIN || b == DMIN) return DMIN;
  if (a == DMAX || b == DMAX) return DMAX;
  int32_t c = (int32_t)((uint32_t)a + (uint32_t)b);
  if (((a ^ b) & DMIN) == 0 && ((c ^ a) & DMIN) != 0) return a < 0 ? DMIN : DMAX;
  return c;
}
int32_t orc_mul(int32_t a, int32_t b) {
  if (a == DMIN) return b < 0 ? DMAX : DMIN;
  if (b == DMIN) return a < 0 ? DMAX : DMIN;
  if (a == DMAX) return b < 0 ? DMIN : DMAX;
  if (b == DMAX) return a < 0 ? DMIN : DMAX;
  int64_t c = (int64_t)a * (int64_t)b;
  int32_t hi = (int32_t)(c >> 32), lo = (int32_t)c;
  if (hi != (lo >> 31)) return hi < 0 ? DMIN : DMAX;
  return lo;
}
int32_t orc_min(int32_t a, int32_t b) { return a < b ? a : b; }
int32_t orc_max(int32_t a, int32_t b) { return a > b ? a : b; }

static int is_value(val_t v) { return v.lo == v.hi; }
static int is_true(val_t v) { return v.lo > 0 || v.hi < 0; }
static int is_false(val_t v) { return v.lo == 0 && v.hi == 0; }
static val_t mkv(int32_t lo, int32_t hi) { val_t v; v.lo = lo; v.hi = hi; return v; }

/* ---- eval.c ----------------------------------------------------------------------- */
static val_t ev(const orc *o, int n) {
  switch (o->op[n]) {
  case CSOLVE_OP_VAR: return o->dom[o->l[n]];
  case CSOLVE_OP_CONST: return mkv(o->l[n], o->r[n]);
  case CSOLVE_OP_EQ: {
    val_t a = ev(o, o->l[n]), b = ev(o, o->r[n]);
    if (a.lo == DMIN || a.hi == DMAX || b.lo == DMIN || b.hi == DMAX) return mkv(0, 1);
    if (a.hi == b.hi && a.lo == b.lo && a.hi == a.lo) return mkv(1, 1);
    if (a.hi < b.lo || a.lo > b.hi) return mkv(0, 0);
    return mkv(0, 1);
  }
  case CSOLVE_OP_LT: {
    val_t a = ev(o, o->l[n]), b = ev(o, o->r[n]);
    if (a.lo == DMIN || a.hi == DMAX || b.lo == DMIN || b.hi == DMAX) return mkv(0, 1);
    if (a.hi < b.lo) return mkv(1, 1);
    if (a.lo >= b.hi) return mkv(0, 0);
    return mkv(0, 1);
  }
  case CSOLVE_OP_NEG: { val_t a = ev(o, o->l[n]); return mkv(orc_neg(a.hi), orc_neg(a.lo)); }
  case CSOLVE_OP_ADD: {
    val_t a = ev(o, o->l[n]), b = ev(o, o->r[n]);
    return mkv(orc_add(a.lo, b.lo), orc_add(a.hi, b.hi));
  }
  case CSOLVE_OP_MUL: {
    val_t a = ev(o, o->l[n]), b = ev(o, o->r[n]);
    int32_t ll = orc_mul(a.lo, b.lo), lh = orc_mul(a.lo, b.hi), hl = orc_mul(a.hi, b.lo), hh = orc_mul(a.hi, b.hi);
    return mkv(orc_min(orc_min(ll, lh), orc_min(hl, hh)), orc_max(orc_max(ll, lh), orc_max(hl, hh)));
  }
  case CSOLVE_OP_NOT: {
    val_t a = ev(o, o->l[n]);
    if (is_true(a)) return mkv(0, 0);
    if (is_false(a)) return mkv(1, 1);
    return mkv(0, 1);
  }
  case CSOLVE_OP_AND: {
    val_t a = ev(o, o->l[n]);
    if (is_false(a)) return mkv(0, 0);
    val_t b = ev(o, o->r[n]);
    if (is_false(b)) return mkv(0, 0);
    if (is_true(a) && is_true(b)) return mkv(1, 1);
    return mkv(0, 1);
  }
  default: { /* OR */
    val_t a = ev(o, o->l[n]);
    if (is_true(a)) return mkv(1, 1);
    val_t b = ev(o, o->r[n]);
    if (is_true(b)) return mkv(1, 1);
    if (is_false(a) && is_false(b)) return mkv(0, 0);
    return mkv(0, 1);
  }
  }
}

/* ---- strategy.c: comparison + heap ---------------------------------------------------- */
static int64_t var_cmp(const orc *o, int e1, int e2) {
  val_t v1 = o->dom[e1], v2 = o->dom[e2];
  int64_t cmp = 0;
  switch (o->order) {
  case CSOLVE_ORDER_SMALLEST_DOMAIN: {
    int32_t d1 = orc_add(v1.lo, orc_neg(v1.hi)), d2 = orc_add(v2.hi, orc_neg(v2.lo));
    cmp = orc_add(d2, d1); break;
  }
  case CSOLVE_ORDER_LARGEST_DOMAIN: {
    int32_t d1 = orc_add(v1.hi, orc_neg(v1.lo)), d2 = orc_add(v2.lo, orc_neg(v2.hi));
    cmp = orc_add(d1, d2); break;
  }
  case CSOLVE_ORDER_SMALLEST_VALUE: cmp = orc_add(v2.lo, orc_neg(v1.lo)); break;
  case CSOLVE_ORDER_LARGEST_VALUE: cmp = orc_add(v1.hi, orc_neg(v2.hi)); break;
  default: cmp = 0;
  }
  if (o->prefer_failing && cmp == 0) {
    /* the reference computes this difference in int64 and returns it as int (strategy.c:116-120) */
    cmp = (int)(o->prio[e1] - o->prio[e2]);
  }
  return cmp;
}
static void hswap(orc *o, int a, int b) {
  int t = o->heap[a]; o->heap[a] = o->heap[b]; o->hpos[o->heap[a]] = a;
  o->heap[b] = t; o->hpos[t] = b;
}
static void hup(orc *o, int pos) {
  while (pos > 0 && var_cmp(o, o->heap[(pos - 1) / 2], o->heap[pos]) < 0) { hswap(o, pos, (pos - 1) / 2); pos = (pos - 1) / 2; }
}
static void hdown(orc *o, int pos) {
  for (;;) {
    int lp = 2 * pos + 1, rp = 2 * pos + 2, best = pos;
    if (lp < o->heap_size && var_cmp(o, o->heap[lp], o->heap[best]) > 0) best = lp;
    if (rp < o->heap_size && var_cmp(o, o->heap[rp], o->heap[best]) > 0) best = rp;
    if (best == pos) break;
    hswap(o, best, pos); pos = best;
  }
}
static void hpush(orc *o, int v) { int pos = o->heap_size++; o->heap[pos] = v; o->hpos[v] = pos; hup(o, pos); }
static int hpop(orc *o) {
  int v = o->heap[0]; o->hpos[v] = -1; --o->heap_size;
  if (o->heap_size > 0) { o->heap[0] = o->heap[o->heap_size]; o->hpos[o->heap[0]] = 0; hdown(o, 0); }
  return v;
}
static void hupdate(orc *o, int v) { if (o->hpos[v] >= 0) { hup(o, o->hpos[v]); hdown(o, o->hpos[v]); } }

/* ---- util.c trail ------------------------------------------------------------------------ */
static void bind_var(orc *o, int var, val_t v) {
  if (o->trail_len == o->trail_cap) {
    o->trail_cap = o->trail_cap ? 2 * o->trail_cap : 1024;
    o->trail = realloc(o->trail, o->trail_cap * sizeof(trail_t));
  }
  o->trail[o->trail_len].var = var; o->trail[o->trail_len].old = o->dom[var];
  o->trail_len++;
  o->dom[var] = v;
}
static void unbind_to(orc *o, size_t depth) {
  while (o->trail_len > depth) { --o->trail_len; o->dom[o->trail[o->trail_len].var] = o->trail[o->trail_len].old; }
}

/* ---- propagate.c ----------------------------------------------------------------------------- */
static int propagate_clauses(orc *o, int var);
static int prop(orc *o, int n, val_t v);

static int prop_term_var(orc *o, int var, val_t v) {
  val_t t = o->dom[var];
  if (t.lo > v.hi || t.hi < v.lo) {
    o->prio[var]++; hupdate(o, var);         /* propagate_term_confl, -c false */
    return PROP_ERROR;
  }
  int32_t lo = orc_max(t.lo, v.lo), hi = orc_min(t.hi, v.hi);
  if (lo != t.lo || hi != t.hi) {
    bind_var(o, var, mkv(lo, hi));
    o->props++;
    int p = propagate_clauses(o, var);       /* propagate_term_recurse */
    if (p == PROP_ERROR) { o->prio[var]++; hupdate(o, var); return PROP_ERROR; }
    return p + 1;
  }
  return 0;
}
static int prop_eq_false_lr(orc *o, int p, val_t pval, val_t val) {
  if (is_value(val) && val.lo != DMIN && val.lo != DMAX) {
    if (val.lo == pval.lo) return prop(o, p, mkv(val.lo + 1, DMAX));
    if (val.lo == pval.hi) return prop(o, p, mkv(DMIN, val.lo - 1));
  }
  return 0;
}
static int prop_add_lr(orc *o, int p, int c, val_t val) {
  val_t cv = ev(o, c);
  return prop(o, p, mkv(orc_add(val.lo, orc_neg(cv.hi)), orc_add(val.hi, orc_neg(cv.lo))));
}
static int prop_mul_lr(orc *o, int p, int c, val_t val) {
  if (val.lo != DMIN && val.hi != DMIN) {
    val_t cv = ev(o, c);
    if (is_value(cv)) {
      if (((val.lo > 0 || val.hi < 0) && cv.lo == 0) ||
          (is_value(val) && cv.lo != 0 && (val.lo % cv.lo) != 0)) return PROP_ERROR;
      if (cv.lo != 0) {
        /* v.hi == DMAX with cv.lo == -1 is fine in C (no INT_MIN / -1 thanks to the guard above) */
        int32_t lo = val.lo / cv.lo, hi = val.hi / cv.lo;
        return prop(o, p, mkv(orc_min(lo, hi), orc_max(lo, hi)));
      }
    }
  }
  return 0;
}
#define CHECK(x) do { if ((x) == PROP_ERROR) return PROP_ERROR; } while (0)

static int prop(orc *o, int n, val_t v) {
  int l = o->l[n], r = o->r[n];
  switch (o->op[n]) {
  case CSOLVE_OP_VAR: return prop_term_var(o, l, v);
  case CSOLVE_OP_CONST: {
    /* anonymous terminal: a single value, so only the emptiness test can fire */
    if (l > v.hi || r < v.lo) return PROP_ERROR;
    return 0;
  }
  case CSOLVE_OP_EQ:
    if (is_true(v)) {
      int p = prop(o, r, ev(o, l)); CHECK(p);
      int q = prop(o, l, ev(o, r)); CHECK(q);
      return p + q;
    }
    if (is_false(v)) {
      val_t lv = ev(o, l), rv = ev(o, r);
      int p = prop_eq_false_lr(o, r, rv, lv); CHECK(p);
      int q = prop_eq_false_lr(o, l, lv, rv); CHECK(q);
      return p + q;
    }
    return 0;
  case CSOLVE_OP_LT:
    if (is_true(v)) {
      val_t lv = ev(o, l); int p = 0, q = 0;
      if (lv.lo != DMIN && lv.lo != DMAX) { p = prop(o, r, mkv(lv.lo + 1, DMAX)); CHECK(p); }
      val_t rv = ev(o, r);
      if (rv.hi != DMIN && rv.hi != DMAX) { q = prop(o, l, mkv(DMIN, rv.hi - 1)); CHECK(q); }
      return p + q;
    }
    if (is_false(v)) {
      val_t lv = ev(o, l);
      int p = prop(o, r, mkv(DMIN, lv.hi)); CHECK(p);
      val_t rv = ev(o, r);
      int q = prop(o, l, mkv(rv.lo, DMAX)); CHECK(q);
      return p + q;
    }
    return 0;
  case CSOLVE_OP_NEG: return prop(o, l, mkv(orc_neg(v.hi), orc_neg(v.lo)));
  case CSOLVE_OP_ADD: { int p = prop_add_lr(o, r, l, v); CHECK(p); int q = prop_add_lr(o, l, r, v); CHECK(q); return p + q; }
  case CSOLVE_OP_MUL: { int p = prop_mul_lr(o, r, l, v); CHECK(p); int q = prop_mul_lr(o, l, r, v); CHECK(q); return p + q; }
  case CSOLVE_OP_NOT:
    if (is_true(v)) return prop(o, l, mkv(0, 0));
    if (is_false(v)) return prop(o, l, mkv(1, 1));
    return 0;
  case CSOLVE_OP_AND:
  case CSOLVE_OP_OR: {
    int is_and = o->op[n] == CSOLVE_OP_AND;
    int both = is_and ? is_true(v) : is_false(v), either = is_and ? is_false(v) : is_true(v);
    if (both) { int p = prop(o, r, v); CHECK(p); int q = prop(o, l, v); CHECK(q); return p + q; }
    if (either) {
      int p = 0, q = 0;
      val_t lv = ev(o, l);
      if (is_and ? is_true(lv) : is_false(lv)) { p = prop(o, r, v); CHECK(p); }
      val_t rv = ev(o, r);
      if (is_and ? is_true(rv) : is_false(rv)) { q = prop(o, l, v); CHECK(q); }
      return p + q;
    }
    return 0;
  }
  default: return PROP_ERROR;
  }
}

/* propagate.c:488-538 */
static int propagate_clauses(orc *o, int var) {
  uint64_t tag = ++o->tag;
  int r = 0;
  for (int i = o->wptr[var], e = o->wptr[var + 1]; i < e; i++) {
    int c = o->widx[i];
    if (o->ctag[c] > tag) continue;
    o->ctag[c] = tag;
    int p = prop(o, o->cfirst[c + 1] - 1, mkv(1, 1));
    CHECK(p);
    r += p;
  }
  return r;
}

/* objective.c:101-126 */
static void objective_update_val(orc *o) {
  if (o->objective == CSOLVE_OBJ_MIN) {
    int32_t b = orc_add(o->best, orc_neg(1));
    if (o->dom[o->obj_var].hi > b) o->dom[o->obj_var].hi = b;
  } else if (o->objective == CSOLVE_OBJ_MAX) {
    int32_t b = orc_add(o->best, 1);
    if (o->dom[o->obj_var].lo < b) o->dom[o->obj_var].lo = b;
  }
}
/* objective.c:62-78 */
static int objective_better(const orc *o) {
  if (o->objective == CSOLVE_OBJ_MIN) return o->dom[o->obj_var].lo < o->best;
  if (o->objective == CSOLVE_OBJ_MAX) return o->dom[o->obj_var].hi > o->best;
  return 1;
}
static void objective_update_best(orc *o) {
  if (o->objective == CSOLVE_OBJ_MIN) o->best = o->dom[o->obj_var].lo;
  else if (o->objective == CSOLVE_OBJ_MAX) o->best = o->dom[o->obj_var].hi;
}
/* csolve.c:247-253 */
static int check_assignment(orc *o, int var) {
  return propagate_clauses(o, var) == PROP_ERROR ||
         (o->obj_var >= 0 && propagate_clauses(o, o->obj_var) == PROP_ERROR);
}
/* update_solution's test: is_true(eval(root WAND)) (csolve.c:226, eval.c:221-245) */
static int root_true(const orc *o) {
  for (int c = 0; c < o->C; c++) if (!is_true(ev(o, o->cfirst[c + 1] - 1))) return 0;
  return 1;
}

/* ---- construction ----------------------------------------------------------------------------- */
orc *orc_create(const csolve_flat_model *m) {
  orc *o = calloc(1, sizeof(orc));
  o->V = m->n_vars; o->N = m->n_nodes; o->C = m->n_clauses; o->W = m->n_watch;
  o->objective = m->objective; o->obj_var = m->obj_var;
#define DUP(dst, src, n, T) do { o->dst = malloc(((n) > 0 ? (n) : 1) * sizeof(T)); memcpy(o->dst, m->src, (n) * sizeof(T)); } while (0)
  DUP(op, node_op, o->N, uint8_t); DUP(l, node_l, o->N, int32_t); DUP(r, node_r, o->N, int32_t);
  DUP(cfirst, clause_first, o->C + 1, int32_t); DUP(wptr, watch_ptr, o->V + 1, int32_t); DUP(widx, watch_idx, o->W, int32_t);
  DUP(prio0, var_prio, o->V, int64_t);
#undef DUP
  o->root = malloc(o->V * sizeof(val_t));
  for (int v = 0; v < o->V; v++) o->root[v] = mkv(m->var_lo[v], m->var_hi[v]);
  o->dom = malloc(o->V * sizeof(val_t));
  o->prio = malloc(o->V * sizeof(int64_t));
  o->ctag = calloc(o->C > 0 ? o->C : 1, sizeof(uint64_t));
  o->heap = malloc(o->V * sizeof(int)); o->hpos = malloc(o->V * sizeof(int));
  return o;
}
void orc_destroy(orc *o) {
  if (!o) return;
  free(o->op); free(o->l); free(o->r); free(o->cfirst); free(o->wptr); free(o->widx); free(o->prio0);
  free(o->root); free(o->dom); free(o->prio); free(o->ctag); free(o->heap); free(o->hpos); free(o->trail);
  free(o);
}
static void reset(orc *o, int order, int prefer_failing) {
  memcpy(o->dom, o->root, o->V * sizeof(val_t));
  memcpy(o->prio, o->prio0, o->V * sizeof(int64_t));
  memset(o->ctag, 0, (o->C > 0 ? o->C : 1) * sizeof(uint64_t));
  o->tag = 0; o->trail_len = 0; o->props = 0;
  o->order = order; o->prefer_failing = prefer_failing;
  o->heap_size = 0;
  for (int v = 0; v < o->V; v++) o->hpos[v] = -1;
  o->best = o->objective == CSOLVE_OBJ_MIN ? DMAX : (o->objective == CSOLVE_OBJ_MAX ? DMIN : 0);  /* objective.c:38-50 */
}

/* ---- one node transition (csolve.c:448-457) from an arbitrary state ------------------------------- */
int orc_node(orc *o, const int32_t *dom_in, int var, int32_t val, int32_t best, int32_t *dom_out) {
  reset(o, CSOLVE_ORDER_NONE, 1);
  for (int v = 0; v < o->V; v++) o->dom[v] = mkv(dom_in[2 * v], dom_in[2 * v + 1]);
  o->best = best;
  if (!is_value(o->dom[var])) bind_var(o, var, mkv(val, val));     /* step_enter, csolve.c:301-303 */
  objective_update_val(o);
  int failed = check_assignment(o, var);
  for (int v = 0; v < o->V; v++) { dom_out[2 * v] = o->dom[v].lo; dom_out[2 * v + 1] = o->dom[v].hi; }
  return failed;
}
int orc_leaf_true(orc *o, const int32_t *dom_in) {
  for (int v = 0; v < o->V; v++) o->dom[v] = mkv(dom_in[2 * v], dom_in[2 * v + 1]);
  return root_true(o);
}

/* unit-vector hooks: propagate [vlo,vhi] into the root of clause 0 / evaluate it */
int orc_prop_root(orc *o, const int32_t *dom_in, int32_t vlo, int32_t vhi, int32_t *dom_out) {
  reset(o, CSOLVE_ORDER_NONE, 1);
  for (int v = 0; v < o->V; v++) o->dom[v] = mkv(dom_in[2 * v], dom_in[2 * v + 1]);
  int r = prop(o, o->cfirst[1] - 1, mkv(vlo, vhi));
  for (int v = 0; v < o->V; v++) { dom_out[2 * v] = o->dom[v].lo; dom_out[2 * v + 1] = o->dom[v].hi; }
  return r;
}
void orc_eval_root(orc *o, const int32_t *dom_in, int32_t *out2) {
  for (int v = 0; v < o->V; v++) o->dom[v] = mkv(dom_in[2 * v], dom_in[2 * v + 1]);
  val_t r = ev(o, o->cfirst[1] - 1);
  out2[0] = r.lo; out2[1] = r.hi;
}

typedef struct orc_result {
  uint64_t solutions, calls, cuts, props;
  int32_t best, has_solution, hit_limit, pad;
} orc_result;

/* value order of a level (csolve.c:323-338), seed = 0 */
static int step_check(uint32_t iter, val_t b) { return iter <= (uint32_t)(b.hi - b.lo); }
static int32_t step_val(uint32_t i, val_t b) { return (i & 1u) ? (int32_t)(b.hi - (i >> 1)) : (int32_t)(b.lo + (i >> 1)); }

typedef struct { size_t bind_depth; int var; int active; uint32_t iter; val_t bounds; } step_t;

/*
 * solve() of the reference (csolve.c:398-476) with -c false, -r 0 (no Luby restarts), -j 1:
 * dynamic priority heap, prio-- on success / prio++ on failure, restart from level 0 after
 * every improving solution in MIN/MAX mode. CALLS/CUTS/PROPS are the reference's counters.
 * first_solution (optional, V values) receives the first accepted assignment; for MIN/MAX the
 * last (optimal) one.
 */
int orc_solve_reference(orc *o, int order, int prefer_failing, uint64_t max_calls, orc_result *res, int32_t *solution) {
  reset(o, order, prefer_failing);
  memset(res, 0, sizeof(*res));
  int V = o->V;
  for (int v = 0; v < V; v++) hpush(o, v);                      /* strategy_var_order_init */
  step_t *steps = calloc(V > 0 ? V : 1, sizeof(step_t));
  int level = 0;
  for (;;) {
    if (o->objective == CSOLVE_OBJ_ANY && res->solutions > 0) break;   /* found_any */
    if (level == V) {
      int updated = 0;
      if (root_true(o) && objective_better(o)) {
        objective_update_best(o);
        if (solution && (o->objective != CSOLVE_OBJ_ALL || res->solutions == 0))
          for (int v = 0; v < V; v++) solution[v] = o->dom[v].lo;
        res->solutions++;
        updated = 1;
      }
      if (updated && o->objective != CSOLVE_OBJ_ALL) {
        /* level--; RESTART(): unwind to level 0 */
        level--;
        for (int i = level; i != -1; --i) { unbind_to(o, steps[i].bind_depth); hpush(o, steps[i].var); steps[i].active = 0; }
        level = 0;
        continue;
      }
      if (level != 0) { level--; continue; }
      break;
    }
    step_t *s = &steps[level];
    if (!s->active) {
      int var = hpop(o);
      s->active = 1; s->var = var; s->bounds = o->dom[var]; s->iter = 0;
    } else {
      unbind_to(o, s->bind_depth);
      s->iter++;
    }
    if (!step_check(s->iter, s->bounds)) {
      hpush(o, s->var); s->active = 0;
      if (level != 0) { level--; continue; }
      break;
    }
    s->bind_depth = o->trail_len;
    if (!is_value(o->dom[s->var])) { int32_t x = step_val(s->iter, s->bounds); bind_var(o, s->var, mkv(x, x)); }
    /* objective_update_val writes <obj> directly; the reference does not trail it either */
    objective_update_val(o);
    res->calls++;
    if (max_calls && res->calls >= max_calls) { res->hit_limit = 1; break; }
    int failed = check_assignment(o, s->var);
    if (failed) { res->cuts++; o->prio[s->var]++; }
    else { o->prio[s->var]--; level++; }
  }
  free(steps);
  res->props = o->props;
  res->best = o->best;
  res->has_solution = res->solutions > 0;
  return 0;
}

/*
 * The search tree the device path explores: same node transition, same value order, but the
 * branching variable of a level is a function of the node only (static priority order for
 * ORDER_NONE; otherwise the domain-based rule with ties broken by higher parse-time priority,
 * then lower index) and there are no restarts. In ALL mode, solutions / nodes / cuts are
 * therefore independent of how the tree is traversed and must match the device's counters
 * exactly. Plain copy-on-branch DFS.
 */
static int select_var(const orc *o, const val_t *dom, const uint8_t *assigned, int order, const int *static_order, int level) {
  if (order == CSOLVE_ORDER_NONE) return static_order[level];
  int bestv = -1; uint64_t bestk = ~(uint64_t)0;
  for (int v = 0; v < o->V; v++) {
    if (assigned[v]) continue;
    uint32_t primary;
    switch (order) {
    case CSOLVE_ORDER_SMALLEST_DOMAIN: primary = (uint32_t)dom[v].hi - (uint32_t)dom[v].lo; break;
    case CSOLVE_ORDER_LARGEST_DOMAIN: primary = ~((uint32_t)dom[v].hi - (uint32_t)dom[v].lo); break;
    case CSOLVE_ORDER_SMALLEST_VALUE: primary = (uint32_t)dom[v].lo ^ 0x80000000u; break;
    default: primary = ~((uint32_t)dom[v].hi ^ 0x80000000u); break;
    }
    int64_t p = o->prio0[v];
    if (p > INT32_MAX) p = INT32_MAX;
    if (p < INT32_MIN) p = INT32_MIN;
    uint32_t secondary = ~((uint32_t)(int32_t)p ^ 0x80000000u);
    uint64_t k = ((uint64_t)primary << 32) | secondary;
    if (k < bestk) { bestk = k; bestv = v; }
  }
  return bestv;
}

typedef struct { orc *o; int order; int *static_order; uint8_t *assigned; orc_result *res; uint64_t max_calls; int32_t *solution;
                 uint32_t part, n_parts, counter; int split_level; } tree_ctx;

static void tree_rec(tree_ctx *t, int level) {
  orc *o = t->o;
  int V = o->V;
  if (t->res->hit_limit) return;
  if (o->objective == CSOLVE_OBJ_ANY && t->res->solutions > 0) return;
  int var = select_var(o, o->dom, t->assigned, t->order, t->static_order, level);
  val_t bounds = o->dom[var];
  t->assigned[var] = 1;
  for (uint32_t it = 0; step_check(it, bounds); it++) {
    /* this process's share: the (node, subtree) pairs at the split level are dealt round-robin in DFS order; the
     * levels above it are walked by every part and counted by part 0 only */
    if (level == t->split_level && t->n_parts > 1 && t->counter++ % t->n_parts != t->part) continue;
    const int counted = t->n_parts <= 1 || level >= t->split_level || t->part == 0;
    size_t depth = o->trail_len;
    val_t obj_saved = mkv(0, 0);
    if (o->obj_var >= 0) obj_saved = o->dom[o->obj_var];
    if (!is_value(o->dom[var])) { int32_t x = step_val(it, bounds); bind_var(o, var, mkv(x, x)); }
    objective_update_val(o);
    t->res->calls += counted;
    if (t->max_calls && t->res->calls >= t->max_calls) t->res->hit_limit = 1;
    int failed = 0;
    if (o->obj_var >= 0 && o->dom[o->obj_var].lo > o->dom[o->obj_var].hi) failed = 1;  /* device rule: empty <obj> fails */
    if (!failed) failed = check_assignment(o, var);
    if (failed) {
      t->res->cuts += counted;
    } else if (level + 1 == V) {
      if (root_true(o) && objective_better(o)) {
        objective_update_best(o);
        if (t->solution && (o->objective != CSOLVE_OBJ_ALL || t->res->solutions == 0))
          for (int v = 0; v < V; v++) t->solution[v] = o->dom[v].lo;
        t->res->solutions += counted;
      }
    } else {
      tree_rec(t, level + 1);
    }
    unbind_to(o, depth);
    if (o->obj_var >= 0) o->dom[o->obj_var] = obj_saved;
    if (t->res->hit_limit) break;
    if (o->objective == CSOLVE_OBJ_ANY && t->res->solutions > 0) break;
  }
  t->assigned[var] = 0;
}

/* part / n_parts / split_level: the nodes of level split_level (with the subtrees below them) are dealt round-robin
 * to n_parts processes, so that they cover the tree between them; ALL mode and unsatisfiable ANY models: solutions,
 * calls and cuts of the parts add up to the whole tree's (props do not: the levels above the split are replayed). */
int orc_solve_tree_part(orc *o, int order, uint64_t max_calls, uint32_t part, uint32_t n_parts, int split_level,
                        orc_result *res, int32_t *solution) {
  reset(o, order, 0);
  memset(res, 0, sizeof(*res));
  int V = o->V;
  int *so = malloc(V * sizeof(int));
  /* stable sort by priority descending */
  for (int v = 0; v < V; v++) so[v] = v;
  for (int i = 1; i < V; i++) {
    int x = so[i], j = i - 1;
    while (j >= 0 && o->prio0[so[j]] < o->prio0[x]) { so[j + 1] = so[j]; j--; }
    so[j + 1] = x;
  }
  tree_ctx t; t.o = o; t.order = order; t.static_order = so; t.assigned = calloc(V, 1);
  t.res = res; t.max_calls = max_calls; t.solution = solution; t.part = part; t.n_parts = n_parts;
  t.counter = 0; t.split_level = split_level;
  if (V > 0) tree_rec(&t, 0);
  free(so); free(t.assigned);
  res->props = o->props; res->best = o->best; res->has_solution = res->solutions > 0;
  return 0;
}

int orc_solve_tree(orc *o, int order, uint64_t max_calls, orc_result *res, int32_t *solution) {
  return orc_solve_tree_part(o, order, max_calls, 0, 1, 0, res, solution);
}
