// contract.cuh -- interval evaluation and clause contractors of the hot path.
//
// One lane contracts one clause. The functions are templated on a context
// `Cx` that gives access to the search node's domain vector:
//     Dom  cx.dom(v)                 current [lo,hi] of variable v
//     void cx.raise_lo(v, lo)        lo := max(lo, .)   (device: shared-memory atomicMax)
//     void cx.lower_hi(v, hi)        hi := min(hi, .)   (device: shared-memory atomicMin)
// so the same code runs in the kernels (contexts over shared memory) and in
// the host-side unit tests (contexts over a plain array).
//
// Semantics follow the reference exactly:
//   arithmetic   src/arith.c:27-85
//   evaluation   src/eval.c:27-277
//   contraction  src/propagate.c:57-376   
// with one structural difference: a narrowed variable is not propagated
// recursively right away (src/propagate.c:44-54); it is put on the node's
// worklist and its watchers run in a later round. The result is the same
// greatest fixpoint.
#pragma once
#include <stdint.h>
#include "csolve_b200.h"
#include "device_model.h"

#if defined(__CUDACC__)
#define CSOLVE_HD __host__ __device__ __forceinline__
#define CSOLVE_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CSOLVE_HD inline
#define CSOLVE_HD_NOINLINE
#endif

namespace csolve_dev {

static const int32_t DMIN = INT32_MIN;
static const int32_t DMAX = INT32_MAX;

struct Dom { int32_t lo, hi; };

// ---- saturating arithmetic (src/arith.c) --------------------------------------
CSOLVE_HD int32_t sneg(int32_t a) { return a == DMIN ? DMAX : (a == DMAX ? DMIN : -a); }
CSOLVE_HD int32_t sadd(int32_t a, int32_t b) {
  if (a == DMIN || b == DMIN) return DMIN;
  if (a == DMAX || b == DMAX) return DMAX;
  long long c = (long long)a + (long long)b;
  return c < (long long)DMIN ? DMIN : (c > (long long)DMAX ? DMAX : (int32_t)c);
}
CSOLVE_HD int32_t smul(int32_t a, int32_t b) {
  if (a == DMIN) return b < 0 ? DMAX : DMIN;
  if (b == DMIN) return a < 0 ? DMAX : DMIN;
  if (a == DMAX) return b < 0 ? DMIN : DMAX;
  if (b == DMAX) return a < 0 ? DMIN : DMAX;
  long long c = (long long)a * (long long)b;
  return c < (long long)DMIN ? DMIN : (c > (long long)DMAX ? DMAX : (int32_t)c);
}
CSOLVE_HD int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
CSOLVE_HD int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

// truthiness (src/csolve.h:57-67)
CSOLVE_HD bool is_single(Dom v) { return v.lo == v.hi; }
CSOLVE_HD bool is_true(Dom v) { return v.lo > 0 || v.hi < 0; }
CSOLVE_HD bool is_false(Dom v) { return v.lo == 0 && v.hi == 0; }
CSOLVE_HD Dom mk(int32_t lo, int32_t hi) { Dom d; d.lo = lo; d.hi = hi; return d; }

// ---- evaluation: the subtree of a node is a contiguous post-order range, so
//      bottom-up evaluation is one RPN pass (src/eval.c) ------------------------------
CSOLVE_HD Dom eval_binary(int op, Dom a, Dom b) {
  switch (op) {
  case CSOLVE_OP_EQ:
    if (a.lo == DMIN || a.hi == DMAX || b.lo == DMIN || b.hi == DMAX) return mk(0, 1);
    if (a.hi == b.hi && a.lo == b.lo && a.hi == a.lo) return mk(1, 1);
    if (a.hi < b.lo || a.lo > b.hi) return mk(0, 0);
    return mk(0, 1);
  case CSOLVE_OP_LT:
    if (a.lo == DMIN || a.hi == DMAX || b.lo == DMIN || b.hi == DMAX) return mk(0, 1);
    if (a.hi < b.lo) return mk(1, 1);
    if (a.lo >= b.hi) return mk(0, 0);
    return mk(0, 1);
  case CSOLVE_OP_ADD:
    return mk(sadd(a.lo, b.lo), sadd(a.hi, b.hi));
  case CSOLVE_OP_MUL: {
    int32_t ll = smul(a.lo, b.lo), lh = smul(a.lo, b.hi), hl = smul(a.hi, b.lo), hh = smul(a.hi, b.hi);
    return mk(imin(imin(ll, lh), imin(hl, hh)), imax(imax(ll, lh), imax(hl, hh)));
  }
  case CSOLVE_OP_AND:
    if (is_false(a) || is_false(b)) return mk(0, 0);
    if (is_true(a) && is_true(b)) return mk(1, 1);
    return mk(0, 1);
  default: /* CSOLVE_OP_OR */
    if (is_true(a) || is_true(b)) return mk(1, 1);
    if (is_false(a) && is_false(b)) return mk(0, 0);
    return mk(0, 1);
  }
}

template <class Cx>
CSOLVE_HD Dom eval_subtree(Cx &cx, const DevModel &m, int node) {
  Dom st[MAX_DEPTH + 2];
  int sp = 0;
  for (int j = m.node_first[node]; j <= node; ++j) {
    int op = m.node_op[j];
    if (op == CSOLVE_OP_VAR) {
      st[sp++] = cx.dom(m.node_l[j]);
    } else if (op == CSOLVE_OP_CONST) {
      st[sp++] = mk(m.node_l[j], m.node_r[j]);
    } else if (op == CSOLVE_OP_NEG) {
      Dom a = st[sp - 1];
      st[sp - 1] = mk(sneg(a.hi), sneg(a.lo));
    } else if (op == CSOLVE_OP_NOT) {
      Dom a = st[sp - 1];
      st[sp - 1] = is_true(a) ? mk(0, 0) : (is_false(a) ? mk(1, 1) : mk(0, 1));
    } else {
      Dom b = st[--sp];
      Dom a = st[sp - 1];
      st[sp - 1] = eval_binary(op, a, b);
    }
  }
  return st[0];
}

// ---- TERM contractor on a variable (src/propagate.c:57-87) -----------------------
// returns false on an empty intersection (PROP_ERROR)
template <class Cx>
CSOLVE_HD bool contract_var(Cx &cx, int v, int32_t lo, int32_t hi) {
  Dom t = cx.dom(v);
  if (t.lo > hi || t.hi < lo) return false;
  bool ch = false;
  if (lo > t.lo) { cx.raise_lo(v, lo); ch = true; }
  if (hi < t.hi) { cx.lower_hi(v, hi); ch = true; }
  if (ch) cx.count_prop();
  return true;
}

// src/propagate.c:249-271; out: 0 nothing, 1 interval in [*lo,*hi], -1 error
CSOLVE_HD int mul_target(Dom v, Dom cv, int32_t *lo, int32_t *hi) {
  if (v.lo != DMIN && v.hi != DMIN && is_single(cv)) {
    // 32-bit division: the numerators are never INT32_MIN here, so k == -1 cannot overflow (a 64-bit division is a
    // hundred instructions on the device; this function runs for every term of a linear clause in every round)
    const int32_t k = cv.lo;
    if (((v.lo > 0 || v.hi < 0) && k == 0) ||
        (is_single(v) && k != 0 && (v.lo % k) != 0)) {
      return -1;
    }
    if (k != 0) {
      const int32_t a = v.lo / k;
      const int32_t b = v.hi / k;
      *lo = imin(a, b);
      *hi = imax(a, b);
      return 1;
    }
  }
  return 0;
}

struct PFrame { int32_t node, lo, hi, phase; };

// ---- generic clause contractor: propagate VALUE(1) into the clause tree ------------
// An explicit-stack rendering of the mutually recursive propagate_<op>() functions.
// Frames on the stack are pending prop(node, [lo,hi]) calls; `phase` 1 means the right
// child has already been handled and the left child is next (the reference always
// contracts the right operand first, re-evaluates, then the left one).
template <class Cx>
CSOLVE_HD_NOINLINE bool contract_generic(Cx &cx, const DevModel &m, int root, int32_t vlo = 1, int32_t vhi = 1) {
  PFrame st[MAX_DEPTH + 2];
  int sp = 1;
  st[0].node = root; st[0].lo = vlo; st[0].hi = vhi; st[0].phase = 0;
  while (sp > 0) {
    PFrame &f = st[sp - 1];
    const int n = f.node;
    const int op = m.node_op[n];
    const Dom v = mk(f.lo, f.hi);
    const int l = m.node_l[n], r = m.node_r[n];
    switch (op) {
    case CSOLVE_OP_VAR:
      if (!contract_var(cx, l, v.lo, v.hi)) return false;
      sp--;
      break;
    case CSOLVE_OP_CONST:
      // anonymous terminal holding a single value: only the emptiness test matters
      if (l > v.hi || r < v.lo) return false;
      sp--;
      break;
    case CSOLVE_OP_NOT:  // src/propagate.c:289-301
      if (is_true(v)) { f.node = l; f.lo = 0; f.hi = 0; f.phase = 0; }
      else if (is_false(v)) { f.node = l; f.lo = 1; f.hi = 1; f.phase = 0; }
      else sp--;
      break;
    case CSOLVE_OP_NEG:  // src/propagate.c:211-220
      f.node = l; f.lo = sneg(v.hi); f.hi = sneg(v.lo); f.phase = 0;
      break;
    case CSOLVE_OP_EQ:   // src/propagate.c:90-152
      if (is_true(v)) {
        if (f.phase == 0) {
          Dom lv = eval_subtree(cx, m, l);
          f.phase = 1;
          st[sp].node = r; st[sp].lo = lv.lo; st[sp].hi = lv.hi; st[sp].phase = 0; sp++;
        } else {
          Dom rv = eval_subtree(cx, m, r);
          f.node = l; f.lo = rv.lo; f.hi = rv.hi; f.phase = 0;
        }
      } else if (is_false(v)) {
        // both sides are evaluated up front; the left-side target only depends on those
        Dom lv = eval_subtree(cx, m, l), rv = eval_subtree(cx, m, r);
        bool has_l = false, has_r = false;
        int32_t llo = 0, lhi = 0, rlo = 0, rhi = 0;
        if (is_single(lv) && lv.lo != DMIN && lv.lo != DMAX) {
          if (lv.lo == rv.lo) { has_r = true; rlo = lv.lo + 1; rhi = DMAX; }
          else if (lv.lo == rv.hi) { has_r = true; rlo = DMIN; rhi = lv.lo - 1; }
        }
        if (is_single(rv) && rv.lo != DMIN && rv.lo != DMAX) {
          if (rv.lo == lv.lo) { has_l = true; llo = rv.lo + 1; lhi = DMAX; }
          else if (rv.lo == lv.hi) { has_l = true; llo = DMIN; lhi = rv.lo - 1; }
        }
        if (has_l) { f.node = l; f.lo = llo; f.hi = lhi; f.phase = 0; } else { sp--; }
        if (has_r) { st[sp].node = r; st[sp].lo = rlo; st[sp].hi = rhi; st[sp].phase = 0; sp++; }
      } else {
        sp--;
      }
      break;
    case CSOLVE_OP_LT:   // src/propagate.c:155-208
      if (is_true(v)) {
        if (f.phase == 0) {
          Dom lv = eval_subtree(cx, m, l);
          f.phase = 1;
          if (lv.lo != DMIN && lv.lo != DMAX) {
            st[sp].node = r; st[sp].lo = lv.lo + 1; st[sp].hi = DMAX; st[sp].phase = 0; sp++;
          }
        } else {
          Dom rv = eval_subtree(cx, m, r);
          if (rv.hi != DMIN && rv.hi != DMAX) { f.node = l; f.lo = DMIN; f.hi = rv.hi - 1; f.phase = 0; }
          else sp--;
        }
      } else if (is_false(v)) {
        if (f.phase == 0) {
          Dom lv = eval_subtree(cx, m, l);
          f.phase = 1;
          st[sp].node = r; st[sp].lo = DMIN; st[sp].hi = lv.hi; st[sp].phase = 0; sp++;
        } else {
          Dom rv = eval_subtree(cx, m, r);
          f.node = l; f.lo = rv.lo; f.hi = DMAX; f.phase = 0;
        }
      } else {
        sp--;
      }
      break;
    case CSOLVE_OP_ADD:  // src/propagate.c:222-246
      if (f.phase == 0) {
        Dom cv = eval_subtree(cx, m, l);
        f.phase = 1;
        st[sp].node = r; st[sp].lo = sadd(v.lo, sneg(cv.hi)); st[sp].hi = sadd(v.hi, sneg(cv.lo));
        st[sp].phase = 0; sp++;
      } else {
        Dom cv = eval_subtree(cx, m, r);
        f.node = l; f.lo = sadd(v.lo, sneg(cv.hi)); f.hi = sadd(v.hi, sneg(cv.lo)); f.phase = 0;
      }
      break;
    case CSOLVE_OP_MUL: { // src/propagate.c:249-286
      int32_t tlo = 0, thi = 0;
      if (f.phase == 0) {
        int k = mul_target(v, eval_subtree(cx, m, l), &tlo, &thi);
        if (k < 0) return false;
        f.phase = 1;
        if (k > 0) { st[sp].node = r; st[sp].lo = tlo; st[sp].hi = thi; st[sp].phase = 0; sp++; }
      } else {
        int k = mul_target(v, eval_subtree(cx, m, r), &tlo, &thi);
        if (k < 0) return false;
        if (k > 0) { f.node = l; f.lo = tlo; f.hi = thi; f.phase = 0; } else sp--;
      }
      break;
    }
    case CSOLVE_OP_AND:  // src/propagate.c:346-361
    case CSOLVE_OP_OR: { // src/propagate.c:364-376
      const bool both = (op == CSOLVE_OP_AND) ? is_true(v) : is_false(v);
      const bool either = (op == CSOLVE_OP_AND) ? is_false(v) : is_true(v);
      if (both) {
        if (f.phase == 0) {
          f.phase = 1;
          st[sp].node = r; st[sp].lo = v.lo; st[sp].hi = v.hi; st[sp].phase = 0; sp++;
        } else {
          f.node = l; f.phase = 0;
        }
      } else if (either) {
        // push the value to the side whose sibling already is the neutral element
        if (f.phase == 0) {
          Dom lv = eval_subtree(cx, m, l);
          f.phase = 1;
          if ((op == CSOLVE_OP_AND) ? is_true(lv) : is_false(lv)) {
            st[sp].node = r; st[sp].lo = v.lo; st[sp].hi = v.hi; st[sp].phase = 0; sp++;
          }
        } else {
          Dom rv = eval_subtree(cx, m, r);
          if ((op == CSOLVE_OP_AND) ? is_true(rv) : is_false(rv)) { f.node = l; f.phase = 0; }
          else sp--;
        }
      } else {
        sp--;
      }
      break;
    }
    default:
      return false;
    }
  }
  return true;
}

// ---- the same contractor with the operand values of one call memoised ---------------------------------
// propagate_<op>() re-evaluates the sibling subtree at every level (eval_*() from scratch, src/eval.c): on wcet's
// 57-node objective clause, a left-deep sum 16 levels deep, one contraction costs about n^2 / 4 node evaluations.
// Here the value of EVERY node of the clause is computed once, bottom-up over the contiguous post-order range, and
//   * the operand evaluated BEFORE anything is narrowed at a level (the left one) is taken from that table
//     (a variable leaf is re-read: free),
//   * the operand evaluated AFTER the other side was narrowed (the right one) has its own range re-evaluated, which
//     also refreshes the table for the levels below.
// Without repeated variables this is value for value the reference's sequence; when a variable occurs in both
// operands the left value may be the one from the start of the call, i.e. wider, so a single call can narrow less
// than the reference's. The FIXPOINT is the same: at a state where a call changes nothing its table is exact, so it
// acts exactly like the reference's call there, and a weaker sound contractor cannot stop above the greatest common
// fixpoint (tests/test_gpu_parity.py compares post-fixpoint domains with the compiled reference).
static const int MEMO_NODES = 96;    // longest clause the memo table holds; longer ones take contract_generic

template <class Cx>
CSOLVE_HD void eval_range(Cx &cx, const DevModel &m, int from, int to, int first, Dom *val) {
  for (int j = from; j <= to; ++j) {
    const int op = m.node_op[j];
    const int l = m.node_l[j];
    Dom out;
    if (op == CSOLVE_OP_VAR) out = cx.dom(l);
    else if (op == CSOLVE_OP_CONST) out = mk(l, m.node_r[j]);
    else if (op == CSOLVE_OP_NEG) { const Dom a = val[l - first]; out = mk(sneg(a.hi), sneg(a.lo)); }
    else if (op == CSOLVE_OP_NOT) { const Dom a = val[l - first]; out = is_true(a) ? mk(0, 0) : (is_false(a) ? mk(1, 1) : mk(0, 1)); }
    else out = eval_binary(op, val[l - first], val[m.node_r[j] - first]);
    val[j - first] = out;
  }
}

template <class Cx>
CSOLVE_HD_NOINLINE bool contract_generic_memo(Cx &cx, const DevModel &m, int root) {
  const int first = m.node_first[root];
  Dom val[MEMO_NODES];
  eval_range(cx, m, first, root, first, val);
  // value of the operand that is looked at before this level narrows anything
  auto before = [&](int node) -> Dom {
    return m.node_op[node] == CSOLVE_OP_VAR ? cx.dom(m.node_l[node]) : val[node - first];
  };
  // value of the operand that is looked at after its sibling was narrowed
  auto after = [&](int node) -> Dom {
    eval_range(cx, m, m.node_first[node], node, first, val);
    return val[node - first];
  };
  PFrame st[MAX_DEPTH + 2];
  int sp = 1;
  st[0].node = root; st[0].lo = 1; st[0].hi = 1; st[0].phase = 0;
  while (sp > 0) {
    PFrame &f = st[sp - 1];
    const int n = f.node;
    const int op = m.node_op[n];
    const Dom v = mk(f.lo, f.hi);
    const int l = m.node_l[n], r = m.node_r[n];
    switch (op) {
    case CSOLVE_OP_VAR:
      if (!contract_var(cx, l, v.lo, v.hi)) return false;
      sp--;
      break;
    case CSOLVE_OP_CONST:
      if (l > v.hi || r < v.lo) return false;
      sp--;
      break;
    case CSOLVE_OP_NOT:
      if (is_true(v)) { f.node = l; f.lo = 0; f.hi = 0; f.phase = 0; }
      else if (is_false(v)) { f.node = l; f.lo = 1; f.hi = 1; f.phase = 0; }
      else sp--;
      break;
    case CSOLVE_OP_NEG:
      f.node = l; f.lo = sneg(v.hi); f.hi = sneg(v.lo); f.phase = 0;
      break;
    case CSOLVE_OP_EQ:
      if (is_true(v)) {
        if (f.phase == 0) {
          const Dom lv = before(l);
          f.phase = 1;
          st[sp].node = r; st[sp].lo = lv.lo; st[sp].hi = lv.hi; st[sp].phase = 0; sp++;
        } else {
          const Dom rv = after(r);
          f.node = l; f.lo = rv.lo; f.hi = rv.hi; f.phase = 0;
        }
      } else if (is_false(v)) {
        const Dom lv = before(l), rv = before(r);
        bool has_l = false, has_r = false;
        int32_t llo = 0, lhi = 0, rlo = 0, rhi = 0;
        if (is_single(lv) && lv.lo != DMIN && lv.lo != DMAX) {
          if (lv.lo == rv.lo) { has_r = true; rlo = lv.lo + 1; rhi = DMAX; }
          else if (lv.lo == rv.hi) { has_r = true; rlo = DMIN; rhi = lv.lo - 1; }
        }
        if (is_single(rv) && rv.lo != DMIN && rv.lo != DMAX) {
          if (rv.lo == lv.lo) { has_l = true; llo = rv.lo + 1; lhi = DMAX; }
          else if (rv.lo == lv.hi) { has_l = true; llo = DMIN; lhi = rv.lo - 1; }
        }
        if (has_l) { f.node = l; f.lo = llo; f.hi = lhi; f.phase = 0; } else { sp--; }
        if (has_r) { st[sp].node = r; st[sp].lo = rlo; st[sp].hi = rhi; st[sp].phase = 0; sp++; }
      } else {
        sp--;
      }
      break;
    case CSOLVE_OP_LT:
      if (is_true(v)) {
        if (f.phase == 0) {
          const Dom lv = before(l);
          f.phase = 1;
          if (lv.lo != DMIN && lv.lo != DMAX) {
            st[sp].node = r; st[sp].lo = lv.lo + 1; st[sp].hi = DMAX; st[sp].phase = 0; sp++;
          }
        } else {
          const Dom rv = after(r);
          if (rv.hi != DMIN && rv.hi != DMAX) { f.node = l; f.lo = DMIN; f.hi = rv.hi - 1; f.phase = 0; }
          else sp--;
        }
      } else if (is_false(v)) {
        if (f.phase == 0) {
          const Dom lv = before(l);
          f.phase = 1;
          st[sp].node = r; st[sp].lo = DMIN; st[sp].hi = lv.hi; st[sp].phase = 0; sp++;
        } else {
          const Dom rv = after(r);
          f.node = l; f.lo = rv.lo; f.hi = DMAX; f.phase = 0;
        }
      } else {
        sp--;
      }
      break;
    case CSOLVE_OP_ADD:
      if (f.phase == 0) {
        const Dom cv = before(l);
        f.phase = 1;
        st[sp].node = r; st[sp].lo = sadd(v.lo, sneg(cv.hi)); st[sp].hi = sadd(v.hi, sneg(cv.lo));
        st[sp].phase = 0; sp++;
      } else {
        const Dom cv = after(r);
        f.node = l; f.lo = sadd(v.lo, sneg(cv.hi)); f.hi = sadd(v.hi, sneg(cv.lo)); f.phase = 0;
      }
      break;
    case CSOLVE_OP_MUL: {
      int32_t tlo = 0, thi = 0;
      if (f.phase == 0) {
        const int k = mul_target(v, before(l), &tlo, &thi);
        if (k < 0) return false;
        f.phase = 1;
        if (k > 0) { st[sp].node = r; st[sp].lo = tlo; st[sp].hi = thi; st[sp].phase = 0; sp++; }
      } else {
        const int k = mul_target(v, after(r), &tlo, &thi);
        if (k < 0) return false;
        if (k > 0) { f.node = l; f.lo = tlo; f.hi = thi; f.phase = 0; } else sp--;
      }
      break;
    }
    case CSOLVE_OP_AND:
    case CSOLVE_OP_OR: {
      const bool both = (op == CSOLVE_OP_AND) ? is_true(v) : is_false(v);
      const bool either = (op == CSOLVE_OP_AND) ? is_false(v) : is_true(v);
      if (both) {
        if (f.phase == 0) {
          f.phase = 1;
          st[sp].node = r; st[sp].lo = v.lo; st[sp].hi = v.hi; st[sp].phase = 0; sp++;
        } else {
          f.node = l; f.phase = 0;
        }
      } else if (either) {
        if (f.phase == 0) {
          const Dom lv = before(l);
          f.phase = 1;
          if ((op == CSOLVE_OP_AND) ? is_true(lv) : is_false(lv)) {
            st[sp].node = r; st[sp].lo = v.lo; st[sp].hi = v.hi; st[sp].phase = 0; sp++;
          }
        } else {
          const Dom rv = after(r);
          if ((op == CSOLVE_OP_AND) ? is_true(rv) : is_false(rv)) { f.node = l; f.phase = 0; }
          else sp--;
        }
      } else {
        sp--;
      }
      break;
    }
    default:
      return false;
    }
  }
  return true;
}

// clauses the memo table cannot hold (and host-side uses) take the plain rendering
template <class Cx>
CSOLVE_HD bool contract_tree(Cx &cx, const DevModel &m, int root) {
#if defined(__CUDA_ARCH__)
  if (root - m.node_first[root] < MEMO_NODES) return contract_generic_memo(cx, m, root);
#endif
  return contract_generic(cx, m, root);
}

// ---- linear clause  x_obj == konst + SUM (+-) k_i * x_i, one term per lane ------------------------------
// What the nested propagate_eq / propagate_add / propagate_neg / propagate_mul calls do to such a tree
// (src/propagate.c:90-286), flattened: the sum's bounds go to x_obj; term i must lie in x_obj - (sum of the others),
// and is then contracted by its own node's rule (a bare variable: intersect; a MUL node: propagate_mul's division and
// divisibility rules against BOTH operands, whichever order they are written in). A call of the nested version
// narrows the right operand first and re-evaluates; this one works from a snapshot, so a single call may narrow
// less, but at a state that a call leaves unchanged both do the same: same fixpoint, same failures.
// Nothing here can saturate: compile.cpp only emits a LinClause when the magnitudes stay far below 2^30.
struct LinLane { int32_t var, k, flags; Dom X; int32_t tlo, thi; };

template <class Cx>
CSOLVE_HD LinLane lin_lane_load(Cx &cx, const DevModel &m, const LinClause &L, int lane) {
  LinLane t; t.var = -1; t.k = 0; t.flags = 0; t.X = mk(0, 0); t.tlo = 0; t.thi = 0;
  if (lane < L.n_terms) {
#if defined(__CUDA_ARCH__)
    const LinTerm q = cx.lin_term[L.first + lane];        // the kernels' contexts carry the (shared-memory) tables
#else
    const LinTerm q = m.lin_term[L.first + lane];
#endif
    t.var = q.var & LIN_VAR; t.flags = q.var & (LIN_NEG | LIN_MUL); t.k = q.k;
    t.X = cx.dom(t.var);
    const long long a = (long long)t.k * t.X.lo, b = (long long)t.k * t.X.hi;
    int32_t lo = (int32_t)(a < b ? a : b), hi = (int32_t)(a < b ? b : a);
    if (t.flags & LIN_NEG) { const int32_t x = -lo; lo = -hi; hi = x; }
    t.tlo = lo; t.thi = hi;
  }
  return t;
}

// SL / SH: konst + sum of the terms' lower / upper bounds over all lanes; O: x_obj's domain when the call started
template <class Cx>
CSOLVE_HD bool lin_lane_apply(Cx &cx, const LinClause &L, const LinLane &t, int lane, int32_t SL, int32_t SH, Dom O) {
  bool ok = true;
  if (lane == 0) ok = contract_var(cx, L.obj, SL, SH);          // x_obj ∩= eval(sum)
  if (t.var < 0) return ok;
  Dom v = mk(O.lo - (SH - t.thi), O.hi - (SL - t.tlo));         // what this term may be
  if (t.flags & LIN_NEG) v = mk(-v.hi, -v.lo);                  // propagate_neg
  if (!(t.flags & LIN_MUL)) return contract_var(cx, t.var, v.lo, v.hi) && ok;
  int32_t lo = 0, hi = 0;
  // against the variable operand (matters once it is a value): the constant operand must fit
  int r = mul_target(v, t.X, &lo, &hi);
  if (r < 0 || (r > 0 && (t.k > hi || t.k < lo))) return false;
  // against the constant operand: the variable's share
  r = mul_target(v, mk(t.k, t.k), &lo, &hi);
  if (r < 0) return false;
  if (r > 0 && !contract_var(cx, t.var, lo, hi)) return false;
  return ok;
}

// ---- small linear relation  SUM(s_k * x_k) + konst  REL  0  (s_k = +-1, at most four variables), one lane ----------
// The flat form of what propagate_eq / propagate_lt (true and false branch) / propagate_add / propagate_neg do to
// EQ(l, r), LT(l, r), NOT(LT(l, r)) over sums with unit coefficients (src/propagate.c:90-246; derivation in DESIGN.md):
// with A / B the lower / upper bound of D = SUM(s_k * x_k) + konst,
//   D == 0 :  s_k x_k  in  [hi_k - B, lo_k - A]      (lo_k / hi_k: bounds of s_k x_k; the others sum to [A - lo_k, B - hi_k])
//   D <  0 :  only the upper side:  s_k x_k <= lo_k - A - 1
//   D >= 0 :  only the lower side:  s_k x_k >= hi_k - B
// Snapshot semantics (every term against the bounds the call started with): a single call may narrow less than the
// reference's, which re-evaluates after narrowing the right operand; the fixpoint is the same.
template <class Cx>
CSOLVE_HD bool contract_linrel(Cx &cx, const LinRel &R) {
  int32_t lo[4], hi[4];
  int32_t A = R.konst, B = R.konst;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; k++) {
    lo[k] = 0; hi[k] = 0;
    if (k < R.n) {
      const Dom X = cx.dom(R.v[k] & LR_VAR);
      const bool neg = (R.v[k] & LR_NEG) != 0;
      lo[k] = neg ? -X.hi : X.lo;
      hi[k] = neg ? -X.lo : X.hi;
    }
    A += lo[k]; B += hi[k];
  }
  bool ok = true;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; k++) {
    if (k < R.n) {
      // bounds asked of s_k x_k
      int32_t tl = DMIN, th = DMAX;
      if (R.rel == LR_EQ) { tl = hi[k] - B; th = lo[k] - A; }
      else if (R.rel == LR_LT) { th = lo[k] - A - 1; }
      else { tl = hi[k] - B; }
      const bool neg = (R.v[k] & LR_NEG) != 0;
      // ... of x_k itself (an open side stays open)
      const int32_t xl = neg ? (th == DMAX ? DMIN : -th) : tl;
      const int32_t xh = neg ? (tl == DMIN ? DMAX : -tl) : th;
      if (!contract_var(cx, R.v[k] & LR_VAR, xl, xh)) ok = false;
    }
  }
  return ok;
}

// ---- specialised contractors ----------------------------------------------------------
// NOT(EQ(x + c, y)): the false branch of propagate_eq (src/propagate.c:104-134) reached
// through propagate_not (src/propagate.c:289-301) and propagate_add (src/propagate.c:222-246);
// only emitted when nothing can saturate.
template <class Cx>
CSOLVE_HD bool contract_ne_vv(Cx &cx, int x, int y, int32_t c) {
  Dom X = cx.dom(x), Y = cx.dom(y);
  const int32_t Ll = X.lo + c, Lh = X.hi + c;
  bool ok = true;
  if (X.lo == X.hi) {              // left side is a value: contract the right side first
    if (Ll == Y.lo) ok = contract_var(cx, y, Y.lo + 1, DMAX);
    else if (Ll == Y.hi) ok = contract_var(cx, y, DMIN, Y.hi - 1);
    if (!ok) return false;
  }
  if (Y.lo == Y.hi) {              // right side (as evaluated up front) is a value
    if (Y.lo == Ll) ok = contract_var(cx, x, X.lo + 1, DMAX);
    else if (Y.lo == Lh) ok = contract_var(cx, x, DMIN, X.hi - 1);
  }
  return ok;
}

// NOT(EQ(x + k, const)) == (x != c)
template <class Cx>
CSOLVE_HD bool contract_ne_vc(Cx &cx, int x, int32_t c) {
  Dom X = cx.dom(x);
  if (X.lo == c) return contract_var(cx, x, c + 1, DMAX);
  if (X.hi == c) return contract_var(cx, x, DMIN, c - 1);
  return true;
}

// A disjunction of n <= 3 literals (code = var << 1 | negated) over distinct 0/1 variables. On such a tree
// the reference's OR-true / AND-false "push to the side whose sibling is neutral" rule
// (propagate_logic_either, src/propagate.c:320-343) is unit propagation: nothing happens while a literal is
// true or two are undecided; one undecided literal with all others false is made true; all false fails.
template <class Cx>
CSOLVE_HD bool contract_lits(Cx &cx, int n, int32_t l0, int32_t l1, int32_t l2) {
  int n_false = 0, unk = -1;
  const int32_t lit[3] = {l0, l1, l2};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 3; k++) {
    if (k < n) {
      const Dom D = cx.dom(lit[k] >> 1);
      if (D.lo == D.hi) {
        if ((D.lo != 0) != ((lit[k] & 1) != 0)) return true;      // a true literal: the clause holds
        n_false++;
      } else {
        unk = unk < 0 ? lit[k] : -2;                                // -2: more than one undecided
      }
    }
  }
  if (n_false == n) return false;
  if (unk >= 0 && n_false == n - 1) {
    const int32_t val = (unk & 1) ? 0 : 1;
    return contract_var(cx, unk >> 1, val, val);
  }
  return true;
}

// A learned nogood (src/conflict.c: a CONFL node): literals code = var << 1 | value with value in {0,1}.
// propagate_confl (src/propagate.c:395-471): if some literal's variable is a value different from the
// recorded one the nogood does not apply; if every literal but one matches and one variable is not a value,
// the recorded value is removed from that variable's bound; all literals matching is NOT an error.
template <class Cx>
CSOLVE_HD bool contract_nogood(Cx &cx, const int32_t *lits, int n) {
  int32_t unk = -1;
  for (int k = 0; k < n; k++) {
    const int32_t lit = lits[k];
    const Dom D = cx.dom(lit >> 1);
    if (D.lo == D.hi) {
      if (D.lo != (lit & 1)) return true;      // a variable left the recorded value: nothing to infer
    } else {
      if (unk >= 0) return true;               // two variables still open
      unk = lit;
    }
  }
  if (unk < 0) return true;
  const int v = unk >> 1;
  const int32_t val = unk & 1;
  const Dom D = cx.dom(v);
  if (D.lo == val) return contract_var(cx, v, D.lo + 1, DMAX);      // propagate_confl_infer, src/propagate.c:441-456
  if (D.hi == val) return contract_var(cx, v, DMIN, D.hi - 1);
  return true;
}

// One watch record of variable `self` (device_model.h): X is the snapshot of self's domain the
// caller took when it dequeued the variable. NE_VV: the NOT(EQ) clauses between self and one partner;
// both directions of each clause are contracted from the snapshots, exactly like the false branch
// of propagate_eq which evaluates both sides up front (src/propagate.c:122-134).
// A record of a linear clause (WK_GENERIC, n == 2) that reaches this function is interpreted like any generic clause
// (while nogoods are learned: the reason records name single watch records). Otherwise the caller intercepts it with
// wrec_is_linear() and contracts the clause once per round step with the whole warp.
CSOLVE_HD bool wrec_is_linear(uint32_t w0) { return (w0 >> 28) == ((WK_GENERIC << 2) | 2u); }

template <class Cx>
CSOLVE_HD bool contract_watch(Cx &cx, const DevModel &m, int self, Dom X, const WatchRec &rec) {
  const uint32_t kind = wrec_kind(rec.w0);
  const int n = wrec_n(rec.w0);
#if defined(__CUDA_ARCH__)
  if (kind == WK_GENERIC && n == 2) return contract_tree(cx, m, m.clause[cx.lin[wrec_arg(rec.w0)].clause].b);
#else
  if (kind == WK_GENERIC && n == 2) return contract_tree(cx, m, m.clause[m.lin[wrec_arg(rec.w0)].clause].b);
#endif
  if (kind == WK_GENERIC && n == 3) {
#if defined(__CUDA_ARCH__)
    const int4 q0 = reinterpret_cast<const int4 *>(&cx.linrel[wrec_arg(rec.w0)])[0];
    const int4 q1 = reinterpret_cast<const int4 *>(&cx.linrel[wrec_arg(rec.w0)])[1];
    LinRel R; R.rel = q0.x; R.n = q0.y; R.konst = q0.z; R.clause = q0.w; R.v[0] = q1.x; R.v[1] = q1.y; R.v[2] = q1.z; R.v[3] = q1.w;
    return contract_linrel(cx, R);
#else
    return contract_linrel(cx, m.linrel[wrec_arg(rec.w0)]);
#endif
  }
  if (kind == WK_NE_VV) {
    const int y = wrec_arg(rec.w0);
    const Dom Y = cx.dom(y);
    const bool xs = X.lo == X.hi, ys = Y.lo == Y.hi;
    if (!xs && !ys) return true;           // neither side is a value: nothing to contract
    // Predicated form of "if (other == my.lo) lo + 1 else if (other == my.hi) hi - 1" for every
    // offset; the bounds asked for by the offsets of one record are combined and published once.
    int32_t nyl = Y.lo, nyh = Y.hi, nxl = X.lo, nxh = X.hi;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 3; k++) {
      if (k < n) {
        const int32_t c = rec.c[k];
        const int32_t Ll = X.lo + c, Lh = X.hi + c;     // self + c, never saturates (compile.cpp)
        const bool yl = xs && Ll == Y.lo;               // self is the value Ll: partner loses it at a bound
        const bool yh = xs && !yl && Ll == Y.hi;
        nyl = yl ? Y.lo + 1 : nyl;
        nyh = yh ? Y.hi - 1 : nyh;
        const bool xl = ys && Y.lo == Ll;               // partner is the value Y.lo: self loses Y.lo - c
        const bool xh = ys && !xl && Y.lo == Lh;
        nxl = xl ? X.lo + 1 : nxl;
        nxh = xh ? X.hi - 1 : nxh;
      }
    }
    bool ok = true;
    if (nyl != Y.lo || nyh != Y.hi) ok = contract_var(cx, y, nyl, nyh);
    if (nxl != X.lo || nxh != X.hi) ok = contract_var(cx, self, nxl, nxh) && ok;
    return ok;
  }
  if (kind == WK_NE_VC) {
    bool ok = true;
    for (int k = 0; k < n; k++) {
      const int32_t c = rec.c[k];
      if (X.lo == c) ok = contract_var(cx, self, c + 1, DMAX) && ok;
      else if (X.hi == c) ok = contract_var(cx, self, DMIN, c - 1) && ok;
    }
    return ok;
  }
  if (kind == WK_LITS) return contract_lits(cx, n, rec.c[0], rec.c[1], rec.c[2]);
  return contract_tree(cx, m, m.clause[wrec_arg(rec.w0)].b);
}

// ---- "lane owns variable" form ---------------------------------------------------------------------
// Variable i (domain [Xlo,Xhi]) was dequeued; the calling lane owns variable j with domain [lo,hi] and
// mask = lov_pair[i][j] is the set of offsets c of the clauses x_i + c != x_j (bit c + 32).
// Same contraction as contract_watch(WK_NE_VV) seen from i with partner j:
//   i is a value v      : every clause removes v + c from j if it sits on one of j's bounds;
//   j is a value w      : every clause removes w - c from i if it sits on one of i's bounds -- returned as
//                         proposals, combined across lanes by the caller with ballots.
// When both are values and collide, j's domain becomes empty, which the caller detects when it
// dequeues j (the reference fails in the same situation, src/propagate.c:57-66).
struct LovStep { int32_t nlo, nhi; bool plo, phi; };
CSOLVE_HD bool lov_has(unsigned long long mask, int32_t d) {
  const uint32_t s = (uint32_t)(d + 32);
  return s < 64u && ((mask >> s) & 1ull) != 0;
}
CSOLVE_HD LovStep lov_lane_step(unsigned long long mask, int32_t Xlo, int32_t Xhi, int32_t lo, int32_t hi) {
  LovStep r; r.nlo = lo; r.nhi = hi; r.plo = false; r.phi = false;
  if (Xlo == Xhi) {
    if (lov_has(mask, lo - Xlo)) r.nlo = lo + 1;
    if (lov_has(mask, hi - Xlo)) r.nhi = hi - 1;
  } else if (lo == hi) {
    r.plo = lov_has(mask, lo - Xlo);
    r.phi = lov_has(mask, lo - Xhi);
  }
  return r;
}
// ---- forbidden-value sets ---------------------------------------------------------------------------
// When all root domains fit a window of 32 values (bit b = value vbase + b) the same fixpoint is
// reached without re-queuing a variable for every single bound move: F_j collects the values the
// variables that ARE a value (and the constants) forbid for x_j; the contractors of
// src/propagate.c:104-119 remove a forbidden value only while it sits on a bound, i.e. lo moves up
// to the first value >= lo that is not in F_j and hi down to the last value <= hi not in F_j.
// Values strictly inside the interval stay (intervals have no holes), exactly as in the reference.
//
// lov_forbid: variable i became the value w; mask = lov_pair[i][j] (bit c + 32: x_i + c != x_j).
// Forbidden for j: w + c, i.e. bit (w - vbase) + c  ->  mask shifted right by 32 - (w - vbase).
CSOLVE_HD uint32_t lov_forbid(unsigned long long mask, int32_t w, int32_t vbase) {
  return (uint32_t)(mask >> (32 - (w - vbase)));
}
// lov_trim: returns false when no value of [lo,hi] survives (PROP_ERROR)
CSOLVE_HD bool lov_trim(uint32_t F, int32_t vbase, int32_t &lo, int32_t &hi) {
  const int bl = lo - vbase, bh = hi - vbase;                 // 0..31
  uint32_t free_ = ~F & (0xffffffffu << bl) & (0xffffffffu >> (31 - bh));
  if (free_ == 0u) return false;
#if defined(__CUDA_ARCH__)
  lo = vbase + (__ffs((int)free_) - 1);
  hi = vbase + (31 - __clz((int)free_));
#else
  lo = vbase + __builtin_ctz(free_);
  hi = vbase + (31 - __builtin_clz(free_));
#endif
  return true;
}

// forbidden constant c of the dequeued variable (x_i != c): which bound of i it removes
CSOLVE_HD void lov_const_step(int32_t c, int32_t Xlo, int32_t Xhi, bool &plo, bool &phi) {
  const bool cl = c == Xlo;
  plo = plo || cl;
  phi = phi || (!cl && c == Xhi);
}

// One clause contraction = propagate VALUE(1) into clause k (src/propagate.c:514-516).
template <class Cx>
CSOLVE_HD bool contract_clause(Cx &cx, const DevModel &m, const ClauseRec &rec) {
  if ((rec.kind & 0xff) == CK_LITS) return contract_lits(cx, rec.kind >> 8, rec.a, rec.b, rec.c);
  switch (rec.kind) {
  case CK_NE_VV: return contract_ne_vv(cx, rec.a, rec.b, rec.c);
  case CK_NE_VC: return contract_ne_vc(cx, rec.a, rec.c);
  default:       return contract_tree(cx, m, rec.b);
  }
}

// Leaf test for one clause: is_true(eval(clause)) (src/csolve.c:226, src/eval.c:221-245)
template <class Cx>
CSOLVE_HD bool clause_is_true(Cx &cx, const DevModel &m, const ClauseRec &rec) {
  if ((rec.kind & 0xff) == CK_LITS) {
    // eval of the OR / NOT(AND) tree is true iff some literal is a true value (src/eval.c:167-219)
    const int32_t lit[3] = {rec.a, rec.b, rec.c};
    for (int k = 0; k < (rec.kind >> 8); k++) {
      const Dom D = cx.dom(lit[k] >> 1);
      if (D.lo == D.hi && ((D.lo != 0) != ((lit[k] & 1) != 0))) return true;
      if (D.lo != D.hi && (D.lo > 0 || D.hi < 0) && !(lit[k] & 1)) return true;   // positive literal, 0 excluded
    }
    return false;
  }
  switch (rec.kind) {
  case CK_NE_VV: {
    Dom X = cx.dom(rec.a), Y = cx.dom(rec.b);
    return X.hi + rec.c < Y.lo || X.lo + rec.c > Y.hi;
  }
  case CK_NE_VC: {
    Dom X = cx.dom(rec.a);
    return X.hi < rec.c || X.lo > rec.c;
  }
  default:
    return is_true(eval_subtree(cx, m, rec.b));
  }
}

// value tried at iteration i of a level (src/csolve.c:331-338, seed = 0: no restarts)
CSOLVE_HD int32_t step_value(int32_t lo, int32_t hi, uint32_t i) {
  return (i & 1u) ? (int32_t)((uint32_t)hi - (i >> 1)) : (int32_t)((uint32_t)lo + (i >> 1));
}

// objective_update_val (src/objective.c:101-126) applied to <obj>'s domain
CSOLVE_HD Dom objective_tighten(int objective, Dom obj, int32_t best) {
  if (objective == CSOLVE_OBJ_MIN) {
    int32_t b = sadd(best, sneg(1));
    if (obj.hi > b) obj.hi = b;
  } else if (objective == CSOLVE_OBJ_MAX) {
    int32_t b = sadd(best, 1);
    if (obj.lo < b) obj.lo = b;
  }
  return obj;
}

}  // namespace csolve_dev
