// compile.cpp -- host side: validate a csolve_flat_model and compile it into the
// tables the kernels read (device_model.h): subtree ranges, per-clause records with
// the specialised NOT(EQ) forms, the static branching order.
#include "compile.hpp"

#include <algorithm>
#include <cstdlib>
#include <numeric>

namespace csolve_dev {

namespace {

// NOT(EQ(A, B)) operand of the form  var | const | ADD(var, const)
struct Affine { bool ok; bool is_var; int var; int64_t k; };

Affine affine_of(const csolve_flat_model &m, int n) {
  Affine a{false, false, -1, 0};
  int op = m.node_op[n];
  if (op == CSOLVE_OP_VAR) { a.ok = true; a.is_var = true; a.var = m.node_l[n]; return a; }
  if (op == CSOLVE_OP_CONST) {
    if (m.node_l[n] != m.node_r[n]) return a;
    a.ok = true; a.k = m.node_l[n]; return a;
  }
  if (op == CSOLVE_OP_ADD) {
    int l = m.node_l[n], r = m.node_r[n];
    if (m.node_op[l] == CSOLVE_OP_VAR && m.node_op[r] == CSOLVE_OP_CONST && m.node_l[r] == m.node_r[r]) {
      a.ok = true; a.is_var = true; a.var = m.node_l[l]; a.k = m.node_l[r]; return a;
    }
  }
  return a;
}

// Is the tree below n a disjunction of literals?  want_true: the subtree must be true.
//   true:  VAR v -> v ; NOT x -> lits(x, false) ; OR(l, r) -> both
//   false: VAR v -> !v ; NOT x -> lits(x, true) ; AND(l, r) -> both     (De Morgan form of normal_or, src/normalize.c:257-266)
bool collect_lits(const csolve_flat_model &m, int n, bool want_true, std::vector<int> &lits) {
  const int op = m.node_op[n];
  if (op == CSOLVE_OP_VAR) { lits.push_back((m.node_l[n] << 1) | (want_true ? 0 : 1)); return true; }
  if (op == CSOLVE_OP_NOT) return collect_lits(m, m.node_l[n], !want_true, lits);
  if ((op == CSOLVE_OP_OR && want_true) || (op == CSOLVE_OP_AND && !want_true))
    return collect_lits(m, m.node_l[n], want_true, lits) && collect_lits(m, m.node_r[n], want_true, lits);
  return false;
}

// A sum of terms below n:  ADD(a, b) | var | const | MUL(var, const) | MUL(const, var) | NEG(term).
// sign: -1 below an odd number of NEG nodes (only directly above a term: NEG(ADD(..)) is not flattened).
bool collect_terms(const csolve_flat_model &m, int n, std::vector<LinTerm> &terms, int64_t &konst, bool top = true) {
  const int op = m.node_op[n];
  if (op == CSOLVE_OP_ADD) return collect_terms(m, m.node_l[n], terms, konst, false) && collect_terms(m, m.node_r[n], terms, konst, false);
  int t = n;
  int32_t flags = 0;
  if (op == CSOLVE_OP_NEG) { t = m.node_l[n]; flags |= LIN_NEG; }
  const int top_op = m.node_op[t];
  if (top_op == CSOLVE_OP_VAR) { terms.push_back(LinTerm{m.node_l[t] | flags, 1}); return true; }
  if (top_op == CSOLVE_OP_CONST) {
    if (m.node_l[t] != m.node_r[t]) return false;
    konst += (flags & LIN_NEG) ? -(int64_t)m.node_l[t] : (int64_t)m.node_l[t];
    return true;
  }
  if (top_op == CSOLVE_OP_MUL) {
    const int l = m.node_l[t], r = m.node_r[t];
    int v = -1, c = -1;
    if (m.node_op[l] == CSOLVE_OP_VAR && m.node_op[r] == CSOLVE_OP_CONST) { v = l; c = r; }
    else if (m.node_op[l] == CSOLVE_OP_CONST && m.node_op[r] == CSOLVE_OP_VAR) { v = r; c = l; }
    if (v < 0 || m.node_l[c] != m.node_r[c]) return false;
    terms.push_back(LinTerm{m.node_l[v] | flags | LIN_MUL, m.node_l[c]});
    return true;
  }
  (void)top;
  return false;
}

// A sum of variables and constants with unit coefficients below n (ADD / NEG directly above a leaf): terms as
// variable | LR_NEG, constants folded into konst. sign: -1 when the whole sum counts negative (the right operand).
bool collect_unit_terms(const csolve_flat_model &m, int n, int sign, std::vector<int32_t> &terms, int64_t &konst) {
  const int op = m.node_op[n];
  if (op == CSOLVE_OP_ADD) return collect_unit_terms(m, m.node_l[n], sign, terms, konst) && collect_unit_terms(m, m.node_r[n], sign, terms, konst);
  int t = n, s = sign;
  if (op == CSOLVE_OP_NEG) { t = m.node_l[n]; s = -s; }
  if (m.node_op[t] == CSOLVE_OP_VAR) { terms.push_back(m.node_l[t] | (s < 0 ? LR_NEG : 0)); return true; }
  if (m.node_op[t] == CSOLVE_OP_CONST && m.node_l[t] == m.node_r[t]) { konst += s * (int64_t)m.node_l[t]; return true; }
  return false;
}

// every value involved stays far away from the +-infinity sentinels of src/arith.c
const int64_t SAFE = (int64_t)1 << 29;
bool small(int64_t v) { return v > -SAFE && v < SAFE; }

}  // namespace

int compile_model(const csolve_flat_model &m, CompiledModel &out, std::string &err) {
  const int V = m.n_vars, C = m.n_clauses, N = m.n_nodes, W = m.n_watch;
  if (V <= 0 || C < 0 || N < 0 || W < 0) { err = "invalid model sizes"; return CSOLVE_ERR_INVALID; }
  if (m.objective < CSOLVE_OBJ_ANY || m.objective > CSOLVE_OBJ_MAX) { err = "invalid objective function type"; return CSOLVE_ERR_INVALID; }
  const bool opt = m.objective == CSOLVE_OBJ_MIN || m.objective == CSOLVE_OBJ_MAX;
  if (opt != (m.obj_var >= 0) || m.obj_var >= V) { err = "objective variable does not match objective kind"; return CSOLVE_ERR_INVALID; }
  if (m.clause_first[0] != 0 || m.clause_first[C] != N) { err = "clause_first does not cover the nodes"; return CSOLVE_ERR_INVALID; }
  if (m.watch_ptr[0] != 0 || m.watch_ptr[V] != W) { err = "watch_ptr does not cover the watch list"; return CSOLVE_ERR_INVALID; }
  for (int v = 0; v < V; v++) {
    if (m.watch_ptr[v + 1] < m.watch_ptr[v]) { err = "watch_ptr not monotone"; return CSOLVE_ERR_INVALID; }
    if (m.var_lo[v] > m.var_hi[v] || m.var_lo[v] == INT32_MIN || m.var_hi[v] == INT32_MAX) {
      err = "variable domain empty or unbounded"; return CSOLVE_ERR_INVALID;
    }
  }
  for (int w = 0; w < W; w++) {
    if (m.watch_idx[w] < 0 || m.watch_idx[w] >= C) { err = "watch_idx out of range"; return CSOLVE_ERR_INVALID; }
  }

  out.node_op.assign(m.node_op, m.node_op + N);
  out.node_l.assign(m.node_l, m.node_l + N);
  out.node_r.assign(m.node_r, m.node_r + N);
  out.node_first.assign(N, 0);
  out.watch_ptr.assign(m.watch_ptr, m.watch_ptr + V + 1);
  out.watch_idx.assign(m.watch_idx, m.watch_idx + W);
  out.clause.resize(C);
  out.lin.clear(); out.lin_term.clear();
  out.root_dom.resize(2 * (size_t)V);
  for (int v = 0; v < V; v++) { out.root_dom[2 * v] = m.var_lo[v]; out.root_dom[2 * v + 1] = m.var_hi[v]; }

  // subtree ranges + structure check (post-order, children inside the clause and before the parent)
  std::vector<int> depth(N, 1);
  int max_depth = 0, n_generic = 0;
  for (int c = 0; c < C; c++) {
    int b = m.clause_first[c], e = m.clause_first[c + 1];
    if (e <= b) { err = "empty clause"; return CSOLVE_ERR_INVALID; }
    for (int n = b; n < e; n++) {
      int op = m.node_op[n];
      if (op == CSOLVE_OP_VAR) {
        if (m.node_l[n] < 0 || m.node_l[n] >= V) { err = "variable index out of range"; return CSOLVE_ERR_INVALID; }
        out.node_first[n] = n;
      } else if (op == CSOLVE_OP_CONST) {
        if (m.node_l[n] != m.node_r[n]) { err = "anonymous terminal is not a single value"; return CSOLVE_ERR_UNSUPPORTED; }
        out.node_first[n] = n;
      } else if (op == CSOLVE_OP_NEG || op == CSOLVE_OP_NOT) {
        int l = m.node_l[n];
        if (l != n - 1 || l < b) { err = "malformed unary node"; return CSOLVE_ERR_INVALID; }
        out.node_first[n] = out.node_first[l];
        depth[n] = depth[l] + 1;
      } else if (op >= CSOLVE_OP_EQ && op <= CSOLVE_OP_OR) {
        int l = m.node_l[n], r = m.node_r[n];
        if (r != n - 1 || r < b || l < b || l >= r || out.node_first[r] != l + 1) { err = "malformed binary node"; return CSOLVE_ERR_INVALID; }
        out.node_first[n] = out.node_first[l];
        depth[n] = std::max(depth[l], depth[r]) + 1;
      } else {
        err = "invalid operation"; return CSOLVE_ERR_INVALID;
      }
    }
    if (out.node_first[e - 1] != b) { err = "clause is not a single tree"; return CSOLVE_ERR_INVALID; }
    max_depth = std::max(max_depth, depth[e - 1]);

    ClauseRec rec{CK_GENERIC, b, e - 1, 0};
    // NOT(EQ(A, B)) with affine operands
    int root = e - 1;
    if (m.node_op[root] == CSOLVE_OP_NOT && m.node_op[m.node_l[root]] == CSOLVE_OP_EQ) {
      int eq = m.node_l[root];
      Affine A = affine_of(m, m.node_l[eq]), B = affine_of(m, m.node_r[eq]);
      if (A.ok && B.ok && (A.is_var || B.is_var)) {
        bool safe = small(A.k) && small(B.k);
        if (A.is_var) safe = safe && small(m.var_lo[A.var]) && small(m.var_hi[A.var]);
        if (B.is_var) safe = safe && small(m.var_lo[B.var]) && small(m.var_hi[B.var]);
        if (safe) {
          // a variable fixed at root never changes (src/parser_support.c:341): treat it as a constant
          bool a_var = A.is_var && m.var_lo[A.var] != m.var_hi[A.var];
          bool b_var = B.is_var && m.var_lo[B.var] != m.var_hi[B.var];
          int64_t ak = A.k + ((A.is_var && !a_var) ? (int64_t)m.var_lo[A.var] : 0);
          int64_t bk = B.k + ((B.is_var && !b_var) ? (int64_t)m.var_lo[B.var] : 0);
          if (a_var && b_var) {
            if (A.var != B.var) rec = ClauseRec{CK_NE_VV, A.var, B.var, (int32_t)(ak - bk)};
          } else if (a_var) {
            rec = ClauseRec{CK_NE_VC, A.var, 0, (int32_t)(bk - ak)};
          } else if (b_var) {
            rec = ClauseRec{CK_NE_VC, B.var, 0, (int32_t)(ak - bk)};
          }
        }
      }
    }
    if (rec.kind == CK_GENERIC) {
      // a SAT clause: 2..3 literals over distinct variables whose root domains lie in [0,1]. On such a tree
      // the reference's OR / AND / NOT contractors (src/propagate.c:289-376) are exactly unit propagation.
      std::vector<int> lits;
      if (collect_lits(m, root, true, lits) && lits.size() >= 2 && lits.size() <= 3) {
        bool good = true;
        for (size_t i = 0; i < lits.size() && good; i++) {
          const int v = lits[i] >> 1;
          if (m.var_lo[v] < 0 || m.var_hi[v] > 1) good = false;
          for (size_t j = 0; j < i; j++) if ((lits[j] >> 1) == v) good = false;
        }
        if (good) rec = ClauseRec{CK_LITS | ((int32_t)lits.size() << 8), lits[0], lits[1], lits.size() > 2 ? lits[2] : 0};
      }
    }
    if (rec.kind == CK_GENERIC) n_generic++;
    out.clause[c] = rec;
    // EQ(x, sum) / EQ(sum, x) with a linear sum: contracted by the whole warp (the ClauseRec stays generic: the leaf
    // test evaluates the tree). Nothing may be able to saturate (src/arith.c) and x must not occur in the sum.
    if (rec.kind == CK_GENERIC && m.node_op[root] == CSOLVE_OP_EQ && (int)out.lin.size() < MAX_LIN && V <= LIN_VAR) {
      const int l = m.node_l[root], r = m.node_r[root];
      int ov = -1, sum = -1;
      if (m.node_op[l] == CSOLVE_OP_VAR && m.node_op[r] == CSOLVE_OP_ADD) { ov = m.node_l[l]; sum = r; }
      else if (m.node_op[r] == CSOLVE_OP_VAR && m.node_op[l] == CSOLVE_OP_ADD) { ov = m.node_l[r]; sum = l; }
      std::vector<LinTerm> terms;
      int64_t konst = 0;
      if (sum >= 0 && collect_terms(m, sum, terms, konst) && terms.size() >= 2 && terms.size() <= (size_t)MAX_LIN) {
        int64_t span = std::max(std::llabs((int64_t)m.var_lo[ov]), std::llabs((int64_t)m.var_hi[ov])) + std::llabs(konst);
        bool good = true;
        for (const LinTerm &t : terms) {
          const int v = t.var & LIN_VAR;
          if (v == ov) good = false;
          const int64_t bound = std::max(std::llabs((int64_t)m.var_lo[v]), std::llabs((int64_t)m.var_hi[v]));
          if (!small(bound) || !small(t.k)) { good = false; break; }
          span += bound * std::max<int64_t>(1, std::llabs((int64_t)t.k)) + bound;
          if (!small(span)) { good = false; break; }
        }
        if (good) {
          out.lin.push_back(LinClause{ov, (int32_t)terms.size(), (int32_t)out.lin_term.size(), (int32_t)konst, c, 0, 0, 0});
          out.lin_term.insert(out.lin_term.end(), terms.begin(), terms.end());
        }
      }
    }
  }
  std::vector<int> lin_of_clause(C, -1);
  for (size_t i = 0; i < out.lin.size(); i++) lin_of_clause[out.lin[i].clause] = (int)i;
  // Small linear relations: EQ(l, r), LT(l, r), NOT(LT(l, r)) over sums of at most four DISTINCT variables with unit
  // coefficients. What the nested propagate_eq / propagate_lt / propagate_add / propagate_neg calls do to such a tree
  // (src/propagate.c:90-246) has the fixpoint of plain bounds reasoning on the flat relation (same argument as for the
  // linear clause above; checked against the oracle by scripts/diff_fuzz.py and against the reference's node
  // transitions on schedule / wcet). Nothing may be able to saturate.
  out.linrel.clear();
  std::vector<int> linrel_of_clause(C, -1);
  for (int c = 0; c < C; c++) {
    if (out.clause[c].kind != CK_GENERIC || lin_of_clause[c] >= 0 || V > LR_VAR) continue;
    int root = m.clause_first[c + 1] - 1, rel = -1;
    if (m.node_op[root] == CSOLVE_OP_EQ) rel = LR_EQ;
    else if (m.node_op[root] == CSOLVE_OP_LT) rel = LR_LT;
    else if (m.node_op[root] == CSOLVE_OP_NOT && m.node_op[m.node_l[root]] == CSOLVE_OP_LT) { rel = LR_GE; root = m.node_l[root]; }
    if (rel < 0) continue;
    std::vector<int32_t> terms;
    int64_t konst = 0;
    if (!collect_unit_terms(m, m.node_l[root], +1, terms, konst) || !collect_unit_terms(m, m.node_r[root], -1, terms, konst)) continue;
    if (terms.empty() || terms.size() > 4) continue;
    bool good = small(konst);
    int64_t span = std::llabs(konst);
    for (size_t i = 0; i < terms.size() && good; i++) {
      const int v = terms[i] & LR_VAR;
      for (size_t j = 0; j < i; j++) if ((terms[j] & LR_VAR) == v) good = false;
      const int64_t bound = std::max(std::llabs((int64_t)m.var_lo[v]), std::llabs((int64_t)m.var_hi[v]));
      span += 2 * bound + 2;
      if (!small(bound) || !small(span)) good = false;
    }
    if (!good) continue;
    LinRel r{rel, (int32_t)terms.size(), (int32_t)konst, c, {-1, -1, -1, -1}};
    for (size_t i = 0; i < terms.size(); i++) r.v[i] = terms[i];
    linrel_of_clause[c] = (int)out.linrel.size();
    out.linrel.push_back(r);
  }
  if (max_depth > MAX_DEPTH) { err = "clause expression too deep for the device interpreter"; return CSOLVE_ERR_UNSUPPORTED; }

  // dense propagation: few clauses, one lane each (device_model.h)
  out.dense_form.assign(C, int2_t{DF_CLAUSE, 0});
  for (int c = 0; c < C; c++) {
    if (lin_of_clause[c] >= 0) out.dense_form[c] = int2_t{DF_LIN, lin_of_clause[c]};
    else if (linrel_of_clause[c] >= 0) out.dense_form[c] = int2_t{DF_LINREL, linrel_of_clause[c]};
  }
  out.host.dense = (C >= 1 && C <= 32) ? 1 : 0;

  // ---- watch records: per variable, the NOT(EQ) clauses grouped by partner, then the generic ones ----
  out.wrec.clear();
  out.wrec_ptr.assign(V + 1, 0);
  if (V >= (1 << 28) || C >= (1 << 28)) { err = "model too large for the watch record encoding"; return CSOLVE_ERR_UNSUPPORTED; }
  for (int v = 0; v < V; v++) {
    out.wrec_ptr[v] = (int32_t)out.wrec.size();
    struct Group { int partner; std::vector<int32_t> offs; };
    std::vector<Group> vv;                 // NE_VV groups in first-seen order
    std::vector<int32_t> consts;           // NE_VC constants
    std::vector<int32_t> generic, lit_clauses;
    for (int w = m.watch_ptr[v]; w < m.watch_ptr[v + 1]; w++) {
      const int c = m.watch_idx[w];
      const ClauseRec &rec = out.clause[c];
      if (rec.kind == CK_NE_VV && (rec.a == v || rec.b == v)) {
        // clause: a + c != b. Seen from a: self + c != partner; seen from b: self + (-c) != partner
        const int partner = rec.a == v ? rec.b : rec.a;
        const int32_t off = rec.a == v ? rec.c : -rec.c;
        Group *g = nullptr;
        for (auto &x : vv) if (x.partner == partner) { g = &x; break; }
        if (!g) { vv.push_back(Group{partner, {}}); g = &vv.back(); }
        if (std::find(g->offs.begin(), g->offs.end(), off) == g->offs.end()) g->offs.push_back(off);
      } else if (rec.kind == CK_NE_VC && rec.a == v) {
        if (std::find(consts.begin(), consts.end(), rec.c) == consts.end()) consts.push_back(rec.c);
      } else if ((rec.kind & 0xff) == CK_LITS) {
        lit_clauses.push_back(c);
      } else {
        generic.push_back(c);
      }
    }
    auto emit = [&](uint32_t kind, int arg, const std::vector<int32_t> &vals) {
      for (size_t i = 0; i < vals.size(); i += 3) {
        WatchRec r; r.c[0] = r.c[1] = r.c[2] = 0;
        const int n = (int)std::min<size_t>(3, vals.size() - i);
        for (int k = 0; k < n; k++) r.c[k] = vals[i + k];
        r.w0 = (kind << 30) | ((uint32_t)n << 28) | (uint32_t)arg;
        out.wrec.push_back(r);
      }
    };
    for (auto &g : vv) emit(WK_NE_VV, g.partner, g.offs);
    if (!consts.empty()) emit(WK_NE_VC, 0, consts);
    for (int c : lit_clauses) {
      const ClauseRec &rec = out.clause[c];
      WatchRec r; r.c[0] = rec.a; r.c[1] = rec.b; r.c[2] = rec.c;
      r.w0 = ((uint32_t)WK_LITS << 30) | ((uint32_t)(rec.kind >> 8) << 28) | (uint32_t)c;
      out.wrec.push_back(r);
    }
    for (int c : generic) {
      WatchRec r; r.c[0] = r.c[1] = r.c[2] = 0;
      r.w0 = lin_of_clause[c] >= 0 ? (WK_GENERIC << 30) | (2u << 28) | (uint32_t)lin_of_clause[c]
             : linrel_of_clause[c] >= 0 ? (WK_GENERIC << 30) | (3u << 28) | (uint32_t)linrel_of_clause[c]
                                        : (WK_GENERIC << 30) | (1u << 28) | (uint32_t)c;
      out.wrec.push_back(r);
    }
  }
  out.wrec_ptr[V] = (int32_t)out.wrec.size();

  // ---- "lane owns variable" tables: every clause is a NOT(EQ) between two variables or a variable
  //      and a constant, at most 32 variables, offsets in [-32, 31], no objective variable ----
  {
    const int K = (V + 31) / 32;                 // variables per lane
    bool lov = V <= 128 && n_generic == 0 && m.obj_var < 0;
    const int stride = 32 * K;
    out.lov_pair.assign((size_t)V * stride, 0);
    out.lov_cptr.assign(V + 1, 0);
    out.lov_cval.clear();
    for (int v = 0; v < V && lov; v++) {
      out.lov_cptr[v] = (int32_t)out.lov_cval.size();
      for (int w = out.wrec_ptr[v]; w < out.wrec_ptr[v + 1] && lov; w++) {
        const WatchRec &r = out.wrec[w];
        const int n = wrec_n(r.w0);
        if (wrec_kind(r.w0) == WK_NE_VV) {
          unsigned long long &e = out.lov_pair[(size_t)v * stride + wrec_arg(r.w0)];
          for (int k = 0; k < n; k++) {
            if (r.c[k] < -32 || r.c[k] > 31) { lov = false; break; }
            e |= 1ull << (r.c[k] + 32);
          }
        } else if (wrec_kind(r.w0) == WK_NE_VC) {
          for (int k = 0; k < n; k++) out.lov_cval.push_back(r.c[k]);
        } else {
          lov = false;
        }
      }
      if (K == 1 && (int)out.lov_cval.size() - out.lov_cptr[v] > 32) lov = false;   // one lane per constant
    }
    out.lov_cptr[V] = (int32_t)out.lov_cval.size();
    int32_t vmin = INT32_MAX, vmax = INT32_MIN;
    for (int v = 0; v < V; v++) { vmin = std::min(vmin, m.var_lo[v]); vmax = std::max(vmax, m.var_hi[v]); }
    bool bits = lov && (int64_t)vmax - (int64_t)vmin < 32;
    out.host.lov = (lov && K == 1) ? 1 : 0;
    out.host.lovk = (bits && K >= 2) ? K : 0;
    // constants outside the window can never sit on a bound: they are simply not representable (and not needed)
    out.host.lov_bits = (bits && K == 1) ? 1 : 0;
    out.host.lov_vbase = vmin;
    // all-different style networks: offset 0 only -> adjacency bit matrix (device_model.h: lov_adj)
    out.lov_adj.assign((size_t)V * K, 0u);
    bool adj_only = bits && K >= 2;
    for (int v = 0; v < V && adj_only; v++)
      for (int j = 0; j < stride; j++) {
        const unsigned long long e = out.lov_pair[(size_t)v * stride + j];
        if (e == 0ull) continue;
        if (e != (1ull << 32)) { adj_only = false; break; }
        out.lov_adj[(size_t)v * K + (j >> 5)] |= 1u << (j & 31);
      }
    out.host.lov_adj_only = adj_only ? 1 : 0;
    out.lov_fconst.assign(V, 0u);
    if (bits) {
      for (int v = 0; v < V; v++)
        for (int k = out.lov_cptr[v]; k < out.lov_cptr[v + 1]; k++) {
          const int64_t b = (int64_t)out.lov_cval[k] - vmin;
          if (b >= 0 && b < 32) out.lov_fconst[v] |= 1u << b;
        }
    }
  }

  // ---- "bit state" tables (pure SAT) ----
  {
    bool sat = V >= 1 && V <= 1024 && m.obj_var < 0 && C > 0;
    for (int v = 0; v < V && sat; v++) if (m.var_lo[v] < 0 || m.var_hi[v] > 1) sat = false;
    for (int c = 0; c < C && sat; c++) {
      const ClauseRec &rec = out.clause[c];
      if ((rec.kind & 0xff) != CK_LITS) { sat = false; break; }
      const int32_t lit[3] = {rec.a, rec.b, rec.c};
      // a literal over a variable that is a value at root would never be visited by an assignment event
      for (int k = 0; k < (rec.kind >> 8); k++) if (m.var_lo[lit[k] >> 1] == m.var_hi[lit[k] >> 1]) sat = false;
    }
    out.sat_occ_ptr.assign(2 * (size_t)V + 1, 0);
    out.sat_occ.clear();
    if (sat) {
      std::vector<std::vector<int2_t>> occ(2 * (size_t)V);
      for (int c = 0; c < C; c++) {
        const ClauseRec &rec = out.clause[c];
        const int n = rec.kind >> 8;
        const int32_t lit[3] = {rec.a, rec.b, rec.c};
        for (int k = 0; k < n; k++) {
          int2_t o{-1, -1};
          int q = 0;
          for (int j = 0; j < n; j++) if (j != k) { (q == 0 ? o.x : o.y) = lit[j]; q++; }
          // the literal var << 1 | negated is false when var == negated: that assignment event visits the clause
          occ[(size_t)lit[k]].push_back(o);
        }
      }
      for (size_t e = 0; e < occ.size(); e++) {
        out.sat_occ_ptr[e] = (int32_t)out.sat_occ.size();
        out.sat_occ.insert(out.sat_occ.end(), occ[e].begin(), occ[e].end());
      }
      out.sat_occ_ptr[2 * (size_t)V] = (int32_t)out.sat_occ.size();
    }
    out.host.sat = sat ? 1 : 0;
  }

  // static branching order: priority descending, index ascending (the reference's heap with
  // -o none -f true orders by env_t.prio only, src/strategy.c:79-121)
  out.order.resize(V);
  std::iota(out.order.begin(), out.order.end(), 0);
  std::stable_sort(out.order.begin(), out.order.end(), [&](int a, int b) { return m.var_prio[a] > m.var_prio[b]; });
  out.prio.resize(V);
  for (int v = 0; v < V; v++) {
    int64_t p = m.var_prio[v];
    out.prio[v] = (int32_t)std::max<int64_t>(INT32_MIN, std::min<int64_t>(INT32_MAX, p));
  }

  DevModel &h = out.host;
  h.n_vars = V; h.n_clauses = C; h.n_nodes = N; h.n_watch = W;
  h.objective = m.objective; h.obj_var = m.obj_var;
  h.mask_words = (V + 31) / 32;
  h.frame_words = frame_words(V, h.mask_words);
  if (h.lovk) h.frame_words = (h.frame_words + V + 3) & ~3;      // + forbidden-value sets
  h.max_depth = max_depth;
  h.n_generic = n_generic;
  h.clause = out.clause.data();
  h.watch_ptr = out.watch_ptr.data();
  h.watch_idx = out.watch_idx.data();
  h.wrec = out.wrec.data();
  h.lov_pair = out.lov_pair.data(); h.lov_cptr = out.lov_cptr.data(); h.lov_cval = out.lov_cval.data();
  h.n_lov_cval = (int32_t)out.lov_cval.size();
  h.lov_fconst = out.lov_fconst.data();
  h.lov_adj = out.lov_adj.data();
  h.lov_smem_bytes = (int32_t)((((size_t)V * 32 * 2 + (V + 1) + out.lov_cval.size() + V) * 4 + 15) & ~(size_t)15);   // K == 1 only
  h.wrec_ptr = out.wrec_ptr.data();
  h.n_wrec = (int32_t)out.wrec.size();
  {
    // watch records, their index, and (16-byte aligned behind them) the small tables a contraction reads: linear
    // relations, linear clauses and their terms -- out of global memory each of them costs an L2 round trip per visit
    size_t bytes = out.wrec.size() * sizeof(WatchRec) + (size_t)(V + 1) * sizeof(int32_t);
    bytes = (bytes + 15) & ~(size_t)15;
    bytes += out.linrel.size() * sizeof(LinRel) + out.lin.size() * sizeof(LinClause) + out.lin_term.size() * sizeof(LinTerm);
    if (h.dense) bytes += out.dense_form.size() * sizeof(int2_t);
    h.table_smem_bytes = bytes <= 40 * 1024 ? (int32_t)((bytes + 15) & ~(size_t)15) : 0;
  }
  h.sat_occ_ptr = out.sat_occ_ptr.data();
  h.sat_occ = out.sat_occ.data();
  h.n_sat_occ = (int32_t)out.sat_occ.size();
  {
    // occurrence records + their index + the static order (16-bit) + the root state (two bit vectors)
    const size_t bytes = out.sat_occ.size() * sizeof(int2_t) + (2 * (size_t)V + 1) * sizeof(int32_t) + (size_t)V * 2 + 2 * (size_t)h.mask_words * 4;
    h.sat_smem_bytes = (h.sat && bytes <= 64 * 1024) ? (int32_t)((bytes + 15) & ~(size_t)15) : 0;
  }
  h.n_linrel = (int32_t)out.linrel.size();
  h.linrel = out.linrel.data();
  h.dense_form = out.dense_form.data();
  h.n_lin_term = (int32_t)out.lin_term.size();
  h.n_lin = (int32_t)out.lin.size();
  h.lin = out.lin.data();
  h.lin_term = out.lin_term.data();
  h.node_op = out.node_op.data();
  h.node_l = out.node_l.data();
  h.node_r = out.node_r.data();
  h.node_first = out.node_first.data();
  h.order = out.order.data();
  h.prio = out.prio.data();
  h.root_dom = out.root_dom.data();
  return CSOLVE_OK;
}

}  // namespace csolve_dev
