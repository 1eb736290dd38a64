// front.cpp -- built-in host front end: csolve input text -> root-normalised
// flat model (csolve_flat_model).
//
// This is the host side of the boundary (CPU code, runs once per instance):
//   parse            grammar of src/parser.y:94-284 via csolve_grammar.h
//   root propagate   src/propagate.c:474-485 with env==NULL terminals (src/propagate.c:57-87)
//   root normalise   src/normalize.c:67-316 (one pass, as driven by src/parser.y:60-69)
//   env_generate     src/parser_support.c:245-257
//   clauses_init     src/parser_support.c:339-396 (watch lists)
//   flatten          SURVEY.md §8a-a20
// The search itself never runs here: there is no CPU fallback for the hot path.
#include "csolve_b200.h"
#include "front.hpp"

#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace csolve_front {

static const int32_t DMIN = INT32_MIN;
static const int32_t DMAX = INT32_MAX;

// ---- saturating arithmetic (src/arith.c:27-85) ------------------------------
static int32_t sneg(int32_t a) { return a == DMIN ? DMAX : a == DMAX ? DMIN : -a; }
static int32_t sadd(int32_t a, int32_t b) {
  if (a == DMIN || b == DMIN) return DMIN;
  if (a == DMAX || b == DMAX) return DMAX;
  int64_t c = (int64_t)a + (int64_t)b;
  if (c < DMIN) return DMIN;   // only possible when both are negative
  if (c > DMAX) return DMAX;   // only possible when both are positive
  return (int32_t)c;
}
static int32_t smul(int32_t a, int32_t b) {
  if (a == DMIN) return b < 0 ? DMAX : DMIN;
  if (b == DMIN) return a < 0 ? DMAX : DMIN;
  if (a == DMAX) return b < 0 ? DMIN : DMAX;
  if (b == DMAX) return a < 0 ? DMIN : DMAX;
  int64_t c = (int64_t)a * (int64_t)b;
  if (c < DMIN) return DMIN;
  if (c > DMAX) return DMAX;
  return (int32_t)c;
}
static int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
static int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

struct Val { int32_t lo, hi; };
static bool single(Val v) { return v.lo == v.hi; }
static bool truthy(Val v) { return v.lo > 0 || v.hi < 0; }
static bool falsy(Val v) { return v.lo == 0 && v.hi == 0; }

enum Kind : uint8_t { K_TERM, K_EQ, K_LT, K_NEG, K_ADD, K_MUL, K_NOT, K_AND, K_OR, K_WAND };

struct Node;
struct Elem { Node *constr; Node *orig; };
struct Node {
  Kind kind;
  Val val{0, 0};        // K_TERM
  int var = -1;         // K_TERM: variable index once env_generate() ran
  Node *l = nullptr, *r = nullptr;
  std::vector<Elem> elems;  // K_WAND
};

struct Var { std::string name; Node *term; int64_t prio = 0; std::vector<Elem *> clauses; };

static const int PROP_ERROR = -1;

struct Front {
  std::deque<Node> arena;
  std::vector<Var> vars;
  std::unordered_map<std::string, int> by_name;
  std::unordered_map<const Node *, int> by_term;
  Node *root = nullptr;
  Node *obj_term = nullptr;
  int objective = CSOLVE_OBJ_ANY;
  bool compute_weights = true;
  int objective_override = -1;
  bool invalid_op = false;       // WAND met where the reference would die with "invalid operation"
  std::string error;
  int error_code = 0;

  Node *mk(Kind k, Node *l = nullptr, Node *r = nullptr) {
    arena.emplace_back();
    Node *n = &arena.back();
    n->kind = k; n->l = l; n->r = r;
    return n;
  }
  Node *mk_term(Val v) { Node *n = mk(K_TERM); n->val = v; return n; }
  void add_var(const std::string &name, Node *term) {
    by_name[name] = (int)vars.size();
    by_term[term] = (int)vars.size();
    vars.push_back(Var{name, term, 0, {}});
  }

  // ---- eval (src/eval.c:27-277) ---------------------------------------------
  Val eval(const Node *n) {
    switch (n->kind) {
    case K_TERM: return n->val;
    case K_EQ: case K_LT: {
      Val a = eval(n->l), b = eval(n->r);
      if (a.lo == DMIN || a.hi == DMAX || b.lo == DMIN || b.hi == DMAX) return {0, 1};
      if (n->kind == K_EQ) {
        if (a.hi == b.hi && a.lo == b.lo && a.hi == a.lo) return {1, 1};
        if (a.hi < b.lo || a.lo > b.hi) return {0, 0};
      } else {
        if (a.hi < b.lo) return {1, 1};
        if (a.lo >= b.hi) return {0, 0};
      }
      return {0, 1};
    }
    case K_NEG: { Val a = eval(n->l); return {sneg(a.hi), sneg(a.lo)}; }
    case K_ADD: { Val a = eval(n->l), b = eval(n->r); return {sadd(a.lo, b.lo), sadd(a.hi, b.hi)}; }
    case K_MUL: {
      Val a = eval(n->l), b = eval(n->r);
      int32_t ll = smul(a.lo, b.lo), lh = smul(a.lo, b.hi), hl = smul(a.hi, b.lo), hh = smul(a.hi, b.hi);
      return {imin(imin(ll, lh), imin(hl, hh)), imax(imax(ll, lh), imax(hl, hh))};
    }
    case K_NOT: { Val a = eval(n->l); if (truthy(a)) return {0, 0}; if (falsy(a)) return {1, 1}; return {0, 1}; }
    case K_AND: {
      Val a = eval(n->l); if (falsy(a)) return {0, 0};
      Val b = eval(n->r); if (falsy(b)) return {0, 0};
      if (truthy(a) && truthy(b)) return {1, 1};
      return {0, 1};
    }
    case K_OR: {
      Val a = eval(n->l); if (truthy(a)) return {1, 1};
      Val b = eval(n->r); if (truthy(b)) return {1, 1};
      if (falsy(a) && falsy(b)) return {0, 0};
      return {0, 1};
    }
    case K_WAND: {
      bool all = true;
      for (const Elem &e : n->elems) {
        Val v = eval(e.constr);
        if (falsy(v)) return {0, 0};
        if (!truthy(v)) all = false;
      }
      if (all) return {1, 1};
      return {0, 1};
    }
    }
    return {0, 1};
  }

  // ---- root propagation: terminals have no env yet, so a narrowing is a plain
  //      overwrite counted as one change (src/propagate.c:57-87) -----------------
  int prop(Node *n, Val v) {
    switch (n->kind) {
    case K_TERM: {
      Val t = n->val;
      if (t.lo > v.hi || t.hi < v.lo) return PROP_ERROR;
      int32_t lo = imax(t.lo, v.lo), hi = imin(t.hi, v.hi);
      if (lo != t.lo || hi != t.hi) { n->val = {lo, hi}; return 1; }
      return 0;
    }
    case K_EQ:
      if (truthy(v)) {
        int p = prop(n->r, eval(n->l)); if (p == PROP_ERROR) return p;
        int q = prop(n->l, eval(n->r)); if (q == PROP_ERROR) return q;
        return p + q;
      }
      if (falsy(v)) {
        Val lv = eval(n->l), rv = eval(n->r);
        int p = prop_neq(n->r, rv, lv); if (p == PROP_ERROR) return p;
        int q = prop_neq(n->l, lv, rv); if (q == PROP_ERROR) return q;
        return p + q;
      }
      return 0;
    case K_LT:
      if (truthy(v)) {
        Val lv = eval(n->l);
        int p = 0, q = 0;
        if (lv.lo != DMIN && lv.lo != DMAX) { p = prop(n->r, {lv.lo + 1, DMAX}); if (p == PROP_ERROR) return p; }
        Val rv = eval(n->r);
        if (rv.hi != DMIN && rv.hi != DMAX) { q = prop(n->l, {DMIN, rv.hi - 1}); if (q == PROP_ERROR) return q; }
        return p + q;
      }
      if (falsy(v)) {
        Val lv = eval(n->l);
        int p = prop(n->r, {DMIN, lv.hi}); if (p == PROP_ERROR) return p;
        Val rv = eval(n->r);
        int q = prop(n->l, {rv.lo, DMAX}); if (q == PROP_ERROR) return q;
        return p + q;
      }
      return 0;
    case K_NEG: return prop(n->l, {sneg(v.hi), sneg(v.lo)});
    case K_ADD: {
      int p = prop_add(n->r, n->l, v); if (p == PROP_ERROR) return p;
      int q = prop_add(n->l, n->r, v); if (q == PROP_ERROR) return q;
      return p + q;
    }
    case K_MUL: {
      int p = prop_mul(n->r, n->l, v); if (p == PROP_ERROR) return p;
      int q = prop_mul(n->l, n->r, v); if (q == PROP_ERROR) return q;
      return p + q;
    }
    case K_NOT:
      if (truthy(v)) return prop(n->l, {0, 0});
      if (falsy(v)) return prop(n->l, {1, 1});
      return 0;
    case K_AND:
      if (truthy(v)) return prop_both(n, v);
      if (falsy(v)) return prop_either(n, v, true);
      return 0;
    case K_OR:
      if (falsy(v)) return prop_both(n, v);
      if (truthy(v)) return prop_either(n, v, false);
      return 0;
    case K_WAND: {
      int r = 0;
      if (truthy(v)) {
        for (Elem &e : n->elems) {
          int p = prop(e.constr, v); if (p == PROP_ERROR) return p;
          r += p;
        }
      }
      return r;
    }
    }
    return 0;
  }
  // src/propagate.c:104-119
  int prop_neq(Node *p, Val pv, Val o) {
    if (single(o) && o.lo != DMIN && o.lo != DMAX) {
      if (o.lo == pv.lo) return prop(p, {o.lo + 1, DMAX});
      if (o.lo == pv.hi) return prop(p, {DMIN, o.lo - 1});
    }
    return 0;
  }
  // src/propagate.c:222-230
  int prop_add(Node *p, Node *c, Val v) {
    Val cv = eval(c);
    return prop(p, {sadd(v.lo, sneg(cv.hi)), sadd(v.hi, sneg(cv.lo))});
  }
  // src/propagate.c:249-271
  int prop_mul(Node *p, Node *c, Val v) {
    if (v.lo != DMIN && v.hi != DMIN) {
      Val cv = eval(c);
      if (single(cv)) {
        if (((v.lo > 0 || v.hi < 0) && cv.lo == 0) ||
            (single(v) && cv.lo != 0 && cdiv_rem(v.lo, cv.lo) != 0)) return PROP_ERROR;
        if (cv.lo != 0) {
          int32_t a = cdiv(v.lo, cv.lo), b = cdiv(v.hi, cv.lo);
          return prop(p, {imin(a, b), imax(a, b)});
        }
      }
    }
    return 0;
  }
  // C's truncating / and % on int32 (INT32_MIN / -1 is excluded by the v.lo != DMIN guard
  // for v.lo; v.hi == DMAX / -1 is fine). Done in 64 bit to stay defined.
  static int32_t cdiv(int32_t a, int32_t b) { return (int32_t)((int64_t)a / (int64_t)b); }
  static int32_t cdiv_rem(int32_t a, int32_t b) { return (int32_t)((int64_t)a % (int64_t)b); }
  // src/propagate.c:304-316
  int prop_both(Node *n, Val v) {
    int p = prop(n->r, v); if (p == PROP_ERROR) return p;
    int q = prop(n->l, v); if (q == PROP_ERROR) return q;
    return p + q;
  }
  // src/propagate.c:320-343
  int prop_either(Node *n, Val v, bool neutral_is_true) {
    int p = 0, q = 0;
    Val lv = eval(n->l);
    if (neutral_is_true ? truthy(lv) : falsy(lv)) { p = prop(n->r, v); if (p == PROP_ERROR) return p; }
    Val rv = eval(n->r);
    if (neutral_is_true ? truthy(rv) : falsy(rv)) { q = prop(n->l, v); if (q == PROP_ERROR) return q; }
    return p + q;
  }
  // src/propagate.c:474-485
  int propagate_root(Node *n, size_t limit) {
    int r = 0, p = 0;
    size_t i = 0;
    do {
      p = prop(n, {1, 1});
      if (p == PROP_ERROR) return p;
      r += p;
    } while (p != 0 && i++ < limit);
    return r;
  }

  // ---- normalisation (src/normalize.c) ------------------------------------------
  static bool is_const(const Node *n) { return n->kind == K_TERM && single(n->val); }
  Node *upd(Node *n, Node *l, Node *r) {
    if (l != n->l || r != n->r) return mk(n->kind, l, r);
    return n;
  }
  Node *upd1(Node *n, Node *l) {
    if (l != n->l) return mk(n->kind, l, nullptr);
    return n;
  }
  // src/normalize.c:66-74: fold to a constant when the value is a single point
  Node *fold(Node *n) {
    Val v = eval(n);
    if (single(v)) return mk_term(v);
    return n;
  }
  Node *norm(Node *n) {
    switch (n->kind) {
    case K_TERM: return n;
    case K_EQ: {
      Node *e = fold(n); if (e != n) return e;
      Node *l = norm(n->l), *r = norm(n->r);
      if (l == r) return mk_term({1, 1});
      return upd(n, l, r);
    }
    case K_LT: return norm_lt(n);
    case K_ADD: return norm_arith(n, 0);
    case K_MUL: return norm_arith(n, 1);
    case K_NEG: case K_NOT: {
      Node *e = fold(n); if (e != n) return e;
      Node *l = norm(n->l);
      if (l->kind == n->kind) return l->l;   // --x / !!x
      return upd1(n, l);
    }
    case K_AND: return norm_logic(n, true, K_OR);
    case K_OR: return norm_logic(n, false, K_AND);
    case K_WAND:
      for (Elem &el : n->elems) {
        Node *c = norm(el.constr);
        if (c != el.constr) el.constr = c;   // patch(); committed at root (src/parser.y:76)
      }
      return n;
    }
    return n;
  }
  // src/normalize.c:102-158
  Node *norm_lt(Node *n) {
    Node *e = fold(n); if (e != n) return e;
    Node *l = norm(n->l), *r = norm(n->r);
    if (l == r) return mk_term({0, 0});
    if (l->kind == K_NEG && r->kind == K_NEG) return upd(n, r->l, l->l);
    if (is_const(l)) {
      if (r->kind == K_ADD && is_const(r->r)) {
        Node *c = mk(K_NEG, r->r, nullptr);
        c = norm(upd(r, l, c));
        return upd(n, c, r->l);
      }
      if (r->kind == K_NEG) return upd(n, r->l, norm(upd1(r, l)));
    }
    if (is_const(r)) {
      if (l->kind == K_ADD && is_const(l->r)) {
        Node *c = mk(K_NEG, l->r, nullptr);
        c = norm(upd(l, r, c));
        return upd(n, l->l, c);
      }
      if (l->kind == K_NEG) return upd(n, norm(upd1(l, r)), l->l);
    }
    return upd(n, l, r);
  }
  // src/normalize.c:161-192
  Node *norm_arith(Node *n, int32_t neutral) {
    Node *e = fold(n); if (e != n) return e;
    Node *l = norm(n->l), *r = norm(n->r);
    if (is_const(l)) return upd(n, r, l);
    if (is_const(r) && r->val.lo == neutral) return l;
    if (r->kind == n->kind && is_const(r->r)) return upd(n, upd(r, l, r->l), r->r);
    if (l->kind == n->kind && is_const(l->r)) return upd(n, l->l, upd(l, r, l->r));
    return upd(n, l, r);
  }
  // src/normalize.c:229-268
  Node *norm_logic(Node *n, bool neutral_is_true, Kind inverse) {
    Node *e = fold(n); if (e != n) return e;
    Node *l = norm(n->l), *r = norm(n->r);
    if (l == r) return l;
    if (l->kind == K_TERM && (neutral_is_true ? truthy(l->val) : falsy(l->val))) return r;
    if (r->kind == K_TERM && (neutral_is_true ? truthy(r->val) : falsy(r->val))) return l;
    if (l->kind == K_NOT && r->kind == K_NOT) {
      Node *c = mk(inverse, l->l, r->l);
      return upd1(l, c);
    }
    return upd(n, l, r);
  }

  // ---- weights (src/parser_support.c:182-242) --------------------------------------
  int count_vars(const Node *n) {
    switch (n->kind) {
    case K_TERM: return single(n->val) ? 0 : 1;
    case K_NEG: case K_NOT: return count_vars(n->l);
    case K_WAND: invalid_op = true; return 0;
    default: return count_vars(n->l) + count_vars(n->r);
    }
  }
  void weighten(const Node *n, int32_t w) {
    switch (n->kind) {
    case K_TERM:
      if (!single(n->val)) { auto it = by_term.find(n); if (it != by_term.end()) vars[it->second].prio += w; }
      return;
    case K_NEG: case K_NOT: weighten(n->l, w); return;
    case K_WAND: invalid_op = true; return;
    default: weighten(n->r, w); weighten(n->l, w); return;
    }
  }

  // ---- watch lists (src/parser_support.c:339-396) -------------------------------------
  void clauses_init(Node *n, Elem *clause) {
    switch (n->kind) {
    case K_TERM:
      if (!single(n->val) && clause != nullptr && n->var >= 0) {
        auto &cl = vars[n->var].clauses;
        bool seen = false;
        for (Elem *e : cl) if (e == clause) { seen = true; break; }
        if (!seen) cl.push_back(clause);
      }
      return;
    case K_WAND:
      for (Elem &el : n->elems) {
        Elem *c = clause;
        if (clause == nullptr && el.constr->kind != K_WAND) c = &el;
        clauses_init(el.constr, c);
      }
      return;
    case K_NEG: case K_NOT: clauses_init(n->l, clause); return;
    default: clauses_init(n->r, clause); clauses_init(n->l, clause); return;
    }
  }
};

}  // namespace csolve_front

// ---- grammar builder glue ---------------------------------------------------------
using csolve_front::Front;
using csolve_front::Node;

#define CSG_EXPR Node *
#define CSG_NULL nullptr

static Node *csg_num(void *ctx, int32_t v) { return ((Front *)ctx)->mk_term({v, v}); }
static Node *csg_ident(void *ctx, const char *name) {
  Front *f = (Front *)ctx;
  auto it = f->by_name.find(name);
  if (it != f->by_name.end()) return f->vars[it->second].term;
  Node *t = f->mk_term({INT32_MIN, INT32_MAX});
  f->add_var(name, t);
  return t;
}
static Node *csg_expr(void *ctx, int op, Node *l, Node *r) {
  Front *f = (Front *)ctx;
  csolve_front::Kind k;
  switch (op) {
  case '=': k = csolve_front::K_EQ; break;
  case '<': k = csolve_front::K_LT; break;
  case '-': k = csolve_front::K_NEG; break;
  case '+': k = csolve_front::K_ADD; break;
  case '*': k = csolve_front::K_MUL; break;
  case '!': k = csolve_front::K_NOT; break;
  case '&': k = csolve_front::K_AND; break;
  default:  k = csolve_front::K_OR; break;
  }
  return f->mk(k, l, r);
}
static void csg_weighten(void *ctx, Node *e, int weight_class) {
  Front *f = (Front *)ctx;
  if (f->compute_weights) {
    int n = f->count_vars(e);
    f->weighten(e, weight_class / (n > 1 ? n : 1));
  }
}
static Node *csg_wand_new(void *ctx) { return ((Front *)ctx)->mk(csolve_front::K_WAND); }
static void csg_wand_append(void *, Node *w, Node *e) { w->elems.push_back({e, e}); }
static Node *csg_objective(void *ctx, int kind, Node *e) {
  Front *f = (Front *)ctx;
  if (f->objective_override >= 0 && (kind == CSOLVE_OBJ_ANY || kind == CSOLVE_OBJ_ALL)) {
    kind = f->objective_override;
  }
  f->objective = kind;
  Node *c;
  if (kind == CSOLVE_OBJ_MIN || kind == CSOLVE_OBJ_MAX) {
    // objective_init(): <obj> = [MIN+1, MAX-1] (src/objective.c:35-37), added after the
    // objective expression's own variables (src/parser.y:118-129)
    f->obj_term = f->mk_term({INT32_MIN + 1, INT32_MAX - 1});
    f->add_var("<obj>", f->obj_term);
    c = kind == CSOLVE_OBJ_MIN ? f->mk(csolve_front::K_EQ, e, f->obj_term)
                               : f->mk(csolve_front::K_EQ, f->obj_term, e);
  } else {
    c = f->mk_term({1, 1});
  }
  f->root = f->mk(csolve_front::K_WAND);
  f->root->elems.push_back({c, c});
  return c;
}
static void csg_constraint(void *ctx, Node *e) { ((Front *)ctx)->root->elems.push_back({e, e}); }
static void csg_error(void *ctx, int is_lexer, int ch, const char *msg, unsigned line) {
  Front *f = (Front *)ctx;
  char buf[256];
  if (is_lexer) snprintf(buf, sizeof(buf), "invalid input `%c' in line %u", ch, line);  // src/csolve.h:514
  else snprintf(buf, sizeof(buf), "%s in line %u", msg, line);                           // src/csolve.h:516
  f->error = buf;
  f->error_code = CSOLVE_ERR_SYNTAX;
}

#include "csolve_grammar.h"

// ---- flat model owner ---------------------------------------------------------------
struct csolve_model {
  csolve_flat_model flat{};
  std::vector<uint8_t> node_op;
  std::vector<int32_t> node_l, node_r, clause_first, watch_ptr, watch_idx, var_lo, var_hi;
  std::vector<int64_t> var_prio;
  std::vector<std::string> names;
  std::vector<const char *> name_ptrs;
};

namespace csolve_front {

static thread_local std::string g_last_error;
void set_last_error(const std::string &s) { g_last_error = s; }
const char *last_error() { return g_last_error.c_str(); }

struct Flattener {
  csolve_model *m;
  std::unordered_map<const Elem *, int> clause_id;
  bool unsupported = false;

  int tree(const Node *n) {
    if (n->kind == K_TERM) {
      if (n->var >= 0) { m->node_op.push_back(CSOLVE_OP_VAR); m->node_l.push_back(n->var); m->node_r.push_back(-1); }
      else { m->node_op.push_back(CSOLVE_OP_CONST); m->node_l.push_back(n->val.lo); m->node_r.push_back(n->val.hi); }
      return (int)m->node_op.size() - 1;
    }
    uint8_t op;
    switch (n->kind) {
    case K_EQ: op = CSOLVE_OP_EQ; break;   case K_LT: op = CSOLVE_OP_LT; break;
    case K_NEG: op = CSOLVE_OP_NEG; break; case K_ADD: op = CSOLVE_OP_ADD; break;
    case K_MUL: op = CSOLVE_OP_MUL; break; case K_NOT: op = CSOLVE_OP_NOT; break;
    case K_AND: op = CSOLVE_OP_AND; break; case K_OR: op = CSOLVE_OP_OR; break;
    default:
      unsupported = true;
      m->node_op.push_back(CSOLVE_OP_CONST); m->node_l.push_back(1); m->node_r.push_back(1);
      return (int)m->node_op.size() - 1;
    }
    int l = tree(n->l);
    int r = (op == CSOLVE_OP_NEG || op == CSOLVE_OP_NOT) ? -1 : tree(n->r);
    m->node_op.push_back(op); m->node_l.push_back(l); m->node_r.push_back(r);
    return (int)m->node_op.size() - 1;
  }
  void wand(Node *w) {
    for (Elem &e : w->elems) {
      if (e.constr->kind == K_WAND) { wand(e.constr); continue; }
      const Node *c = e.constr;
      if (c->kind == K_TERM && c->var < 0 && c->val.lo == 1 && c->val.hi == 1) continue;
      clause_id[&e] = (int)m->clause_first.size() - 1;
      tree(c);
      m->clause_first.push_back((int32_t)m->node_op.size());
    }
  }
};

int build(const char *text, size_t len, const csolve_front_options *opt, csolve_model **out) {
  Front f;
  if (opt != nullptr) {
    f.compute_weights = opt->compute_weights != 0;
    f.objective_override = opt->objective_override;
  }
  int rc = csg_parse(&f, text, len);
  if (rc != 0 || f.root == nullptr) {
    set_last_error(f.error.empty() ? "syntax error" : f.error);
    return CSOLVE_ERR_SYNTAX;
  }
  if (f.invalid_op) {
    set_last_error("invalid operation: 41");   // src/csolve.h:530, OP_WAND = 'A'
    return CSOLVE_ERR_INVALID;
  }
  size_t size = f.vars.size();

  // root phase, src/parser.y:57-69 (normalize() returns the same WAND pointer, so the
  // do/while there runs exactly once)
  int p = f.propagate_root(f.root, size);
  if (p != PROP_ERROR) {
    f.norm(f.root);
    p = f.propagate_root(f.root, size);
  }
  if (p == PROP_ERROR) {
    set_last_error("INFEASIBLE PROBLEM");
    return CSOLVE_ERR_INFEASIBLE;
  }
  // env_generate, src/parser_support.c:245-257
  for (size_t i = 0; i < size; i++) {
    Val v = f.vars[i].term->val;
    if (v.lo == DMIN || v.hi == DMAX) {
      set_last_error("unbounded variable: " + f.vars[i].name);
      return CSOLVE_ERR_UNBOUNDED;
    }
    f.vars[i].term->var = (int)i;
  }
  f.clauses_init(f.root, nullptr);

  std::unique_ptr<csolve_model> m(new csolve_model);
  Flattener fl{m.get()};
  m->clause_first.push_back(0);
  fl.wand(f.root);
  if (fl.unsupported) {
    set_last_error("all_different nested inside an operator is not supported on the device path");
    return CSOLVE_ERR_UNSUPPORTED;
  }
  m->watch_ptr.push_back(0);
  for (size_t i = 0; i < size; i++) {
    const Var &v = f.vars[i];
    m->var_lo.push_back(v.term->val.lo);
    m->var_hi.push_back(v.term->val.hi);
    m->var_prio.push_back(v.prio);
    m->names.push_back(v.name);
    for (Elem *e : v.clauses) m->watch_idx.push_back(fl.clause_id.at(e));
    m->watch_ptr.push_back((int32_t)m->watch_idx.size());
  }
  for (auto &s : m->names) m->name_ptrs.push_back(s.c_str());
  if (m->watch_idx.empty()) m->watch_idx.push_back(0), m->watch_idx.pop_back();

  csolve_flat_model &fm = m->flat;
  fm.n_vars = (int32_t)size;
  fm.n_nodes = (int32_t)m->node_op.size();
  fm.n_clauses = (int32_t)m->clause_first.size() - 1;
  fm.n_watch = (int32_t)m->watch_idx.size();
  fm.objective = f.objective;
  fm.obj_var = f.obj_term ? f.by_term.at(f.obj_term) : -1;
  fm.node_op = m->node_op.data(); fm.node_l = m->node_l.data(); fm.node_r = m->node_r.data();
  fm.clause_first = m->clause_first.data();
  fm.watch_ptr = m->watch_ptr.data(); fm.watch_idx = m->watch_idx.data();
  fm.var_lo = m->var_lo.data(); fm.var_hi = m->var_hi.data(); fm.var_prio = m->var_prio.data();
  fm.var_name = m->name_ptrs.data();
  *out = m.release();
  return CSOLVE_OK;
}

}  // namespace csolve_front

extern "C" int csolve_model_parse(const char *text, size_t len, const csolve_front_options *opt,
                                  csolve_model **out) {
  if (text == nullptr || out == nullptr) return CSOLVE_ERR_INVALID;
  *out = nullptr;
  return csolve_front::build(text, len, opt, out);
}
extern "C" const csolve_flat_model *csolve_model_flat(const csolve_model *m) { return m ? &m->flat : nullptr; }
extern "C" void csolve_model_free(csolve_model *m) { delete m; }
