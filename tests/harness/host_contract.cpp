// host_contract.cpp -- TEST-ONLY harness: runs the product's contractors
// (csolve_b200/csrc/contract.cuh, the code the kernels execute per lane) on the
// host over a plain domain array, with a sequential worklist in place of the warp.
// Not part of the library, not exported through the C ABI, never a fallback.
#include <cstring>
#include <string>
#include <vector>
#include "compile.hpp"
#include "contract.cuh"

using namespace csolve_dev;

struct HostCx {
  int32_t *d;                       // lo,hi pairs
  std::vector<int> *queue;
  std::vector<uint8_t> *queued;
  uint64_t props = 0;
  Dom dom(int v) const { return mk(d[2 * v], d[2 * v + 1]); }
  void mark(int v) { if (!(*queued)[v]) { (*queued)[v] = 1; queue->push_back(v); } }
  void raise_lo(int v, int32_t lo) { if (lo > d[2 * v]) { d[2 * v] = lo; mark(v); } }
  void lower_hi(int v, int32_t hi) { if (hi < d[2 * v + 1]) { d[2 * v + 1] = hi; mark(v); } }
  void count_prop() { props++; }
};

// the warp-cooperative linear contractor with the lanes emulated one after the other (reductions = loops)
static bool host_contract_linear(HostCx &cx, const DevModel &m, int c) {
  const LinClause &L = m.lin[c];
  const Dom O = cx.dom(L.obj);
  LinLane t[32];
  long long SL = L.konst, SH = L.konst;
  for (int lane = 0; lane < 32; lane++) { t[lane] = lin_lane_load(cx, m, L, lane); SL += t[lane].tlo; SH += t[lane].thi; }
  bool ok = true;
  for (int lane = 0; lane < 32; lane++) if (!lin_lane_apply(cx, L, t[lane], lane, (int32_t)SL, (int32_t)SH, O)) ok = false;
  return ok;
}

static CompiledModel g_cm;
static int g_use_wrec = 1;
static std::string g_err;

extern "C" int hc_load(const csolve_flat_model *m, int specialise) {
  int rc = compile_model(*m, g_cm, g_err);
  if (rc != 0) return rc;
  // specialise: 0 = every clause through the interpreter (per-clause watch lists),
  //             1 = compiled watch records (what the kernels run), 2 = specialised per-clause records
  g_use_wrec = specialise == 1;
  if (!specialise) {
    for (size_t c = 0; c < g_cm.clause.size(); c++) {
      g_cm.clause[c] = ClauseRec{CK_GENERIC, m->clause_first[c], m->clause_first[c + 1] - 1, 0};
    }
  }
  return 0;
}
extern "C" const char *hc_error() { return g_err.c_str(); }
extern "C" int hc_n_linear() { return g_cm.host.n_lin; }
extern "C" int hc_n_linrel() { return g_cm.host.n_linrel; }
extern "C" int hc_lov_adj_only() { return g_cm.host.lov_adj_only; }
extern "C" int hc_n_specialised() {
  int n = 0;
  for (auto &c : g_cm.clause) n += c.kind != CK_GENERIC;
  return n;
}

// one node transition (src/csolve.c:448-457); returns 1 if failed
extern "C" int hc_node(const int32_t *dom_in, int var, int32_t val, int32_t best, int32_t *dom_out) {
  const DevModel &m = g_cm.host;
  int V = m.n_vars;
  std::vector<int32_t> d(dom_in, dom_in + 2 * V);
  std::vector<int> queue;
  std::vector<uint8_t> queued(V, 0);
  HostCx cx{d.data(), &queue, &queued};
  if (d[2 * var] != d[2 * var + 1]) { d[2 * var] = val; d[2 * var + 1] = val; }
  cx.mark(var);
  if (m.obj_var >= 0) {
    Dom o = objective_tighten(m.objective, cx.dom(m.obj_var), best);
    d[2 * m.obj_var] = o.lo; d[2 * m.obj_var + 1] = o.hi;
    cx.mark(m.obj_var);
  }
  bool failed = false;
  for (size_t qi = 0; qi < queue.size() && !failed; qi++) {
    int x = queue[qi];
    queued[x] = 0;
    if (d[2 * x] > d[2 * x + 1]) { failed = true; break; }
    if (g_use_wrec) {
      // the path the kernels take: compiled watch records, domain snapshot per dequeued variable
      Dom X = cx.dom(x);
      unsigned lin_hit = 0;
      for (int w = m.wrec_ptr[x]; w < m.wrec_ptr[x + 1]; w++) {
        if (m.n_lin > 0 && wrec_is_linear(m.wrec[w].w0)) { lin_hit |= 1u << wrec_arg(m.wrec[w].w0); continue; }
        if (!contract_watch(cx, m, x, X, m.wrec[w])) { failed = true; break; }
      }
      for (int c = 0; c < m.n_lin && !failed; c++)
        if (((lin_hit >> c) & 1u) && !host_contract_linear(cx, m, c)) failed = true;
    } else {
      for (int w = m.watch_ptr[x]; w < m.watch_ptr[x + 1]; w++) {
        if (!contract_clause(cx, m, m.clause[m.watch_idx[w]])) { failed = true; break; }
      }
    }
  }
  if (!failed) for (int v = 0; v < V; v++) if (d[2 * v] > d[2 * v + 1]) failed = true;
  memcpy(dom_out, d.data(), sizeof(int32_t) * 2 * V);
  return failed ? 1 : 0;
}

// the "lane owns variable" fixpoint with the warp emulated lane by lane (shuffles = array reads,
// ballots = loops); returns -1 when the model is not eligible
extern "C" int hc_node_lov(const int32_t *dom_in, int var, int32_t val, int32_t *dom_out) {
  const DevModel &m = g_cm.host;
  if (!m.lov && !m.lovk) return -1;
  const int V = m.n_vars;
  const int stride = 32 * ((V + 31) / 32);       // row length of lov_pair
  int32_t lo[128], hi[128];
  for (int j = 0; j < 128; j++) { lo[j] = j < V ? dom_in[2 * j] : 0; hi[j] = j < V ? dom_in[2 * j + 1] : 0; }
  if (lo[var] != hi[var]) { lo[var] = val; hi[var] = val; }
  uint32_t changed = var < 32 ? 1u << var : 0u;
  bool failed = false;
  if (m.lov_bits || m.lovk) {
    // forbidden-value-set form (lov_forbid / lov_trim), lanes emulated one after the other
    const int vb = m.lov_vbase;
    // the K-per-lane kernel reads an all-different style network through its adjacency bit matrix (lovk_fixpoint)
    const int K = (V + 31) / 32;
    auto forbid = [&](int i, int j, int32_t w) -> uint32_t {
      if (m.lovk && m.lov_adj_only) return ((m.lov_adj[(size_t)i * K + (j >> 5)] >> (j & 31)) & 1u) ? 1u << (w - vb) : 0u;
      return lov_forbid(m.lov_pair[(size_t)i * stride + j], w, vb);
    };
    uint32_t F[128];
    for (int j = 0; j < V; j++) {
      F[j] = m.lov_fconst[j];
      for (int i = 0; i < V; i++)
        if (dom_in[2 * i] == dom_in[2 * i + 1]) F[j] |= forbid(i, j, dom_in[2 * i]);
    }
    std::vector<int> pend(1, var);
    std::vector<uint8_t> in_pend(V, 0);
    for (size_t qi = 0; qi < pend.size() && !failed; qi++) {
      const int i = pend[qi];
      const int32_t w = lo[i];
      for (int j = 0; j < V; j++) {
        F[j] |= forbid(i, j, w);
        const bool was = lo[j] == hi[j];
        if (!lov_trim(F[j], vb, lo[j], hi[j])) { failed = true; }
        else if (!was && lo[j] == hi[j] && !in_pend[j]) { in_pend[j] = 1; pend.push_back(j); }
      }
    }
    for (int j = 0; j < V; j++) { dom_out[2 * j] = lo[j]; dom_out[2 * j + 1] = hi[j]; }
    return failed ? 1 : 0;
  }
  while (changed && !failed) {
    const int i = __builtin_ctz(changed);
    changed &= changed - 1;
    const int32_t Xlo = lo[i], Xhi = hi[i];
    if (Xlo > Xhi) { failed = true; break; }
    uint32_t bl = 0, bh = 0, chm = 0;
    int32_t nlo[32], nhi[32];
    for (int j = 0; j < 32; j++) {
      nlo[j] = lo[j]; nhi[j] = hi[j];
      if (j >= V) continue;
      LovStep r = lov_lane_step(m.lov_pair[(size_t)i * stride + j], Xlo, Xhi, lo[j], hi[j]);
      bool plo = r.plo, phi = r.phi;
      const int cb = m.lov_cptr[i], ce = m.lov_cptr[i + 1];
      if (cb + j < ce) lov_const_step(m.lov_cval[cb + j], Xlo, Xhi, plo, phi);
      if (r.nlo != lo[j] || r.nhi != hi[j]) chm |= 1u << j;
      nlo[j] = r.nlo; nhi[j] = r.nhi;
      if (plo) bl |= 1u << j;
      if (phi) bh |= 1u << j;
    }
    // constants beyond lane V-1 (a variable may have up to 32 of them)
    for (int j = V; j < 32; j++) {
      bool plo = false, phi = false;
      const int cb = m.lov_cptr[i], ce = m.lov_cptr[i + 1];
      if (cb + j < ce) lov_const_step(m.lov_cval[cb + j], Xlo, Xhi, plo, phi);
      if (plo) bl |= 1u << j;
      if (phi) bh |= 1u << j;
    }
    for (int j = 0; j < 32; j++) { lo[j] = nlo[j]; hi[j] = nhi[j]; }
    if (bl) lo[i] = Xlo + 1;
    if (bh) hi[i] = Xhi - 1;
    if (bl | bh) chm |= 1u << i;
    changed |= chm;
  }
  for (int j = 0; j < V; j++) { dom_out[2 * j] = lo[j]; dom_out[2 * j + 1] = hi[j]; if (lo[j] > hi[j]) failed = true; }
  return failed ? 1 : 0;
}

extern "C" int hc_leaf_true(const int32_t *dom_in) {
  const DevModel &m = g_cm.host;
  std::vector<int32_t> d(dom_in, dom_in + 2 * m.n_vars);
  std::vector<int> queue; std::vector<uint8_t> queued(m.n_vars, 0);
  HostCx cx{d.data(), &queue, &queued};
  for (int c = 0; c < m.n_clauses; c++) if (!clause_is_true(cx, m, m.clause[c])) return 0;
  return 1;
}

// propagate [vlo,vhi] into the root of clause 0 / evaluate it (unit-vector tests)
extern "C" int hc_prop_root(const int32_t *dom_in, int32_t vlo, int32_t vhi, int32_t *dom_out) {
  const DevModel &m = g_cm.host;
  std::vector<int32_t> d(dom_in, dom_in + 2 * m.n_vars);
  std::vector<int> queue; std::vector<uint8_t> queued(m.n_vars, 0);
  HostCx cx{d.data(), &queue, &queued};
  bool ok = contract_generic(cx, m, m.clause[0].b, vlo, vhi);
  memcpy(dom_out, d.data(), sizeof(int32_t) * 2 * m.n_vars);
  return ok ? (int)cx.props : -1;
}
extern "C" void hc_eval_root(const int32_t *dom_in, int32_t *out2) {
  const DevModel &m = g_cm.host;
  std::vector<int32_t> d(dom_in, dom_in + 2 * m.n_vars);
  std::vector<int> queue; std::vector<uint8_t> queued(m.n_vars, 0);
  HostCx cx{d.data(), &queue, &queued};
  Dom v = eval_subtree(cx, m, m.clause[0].b);
  out2[0] = v.lo; out2[1] = v.hi;
}

// unit hook for learned nogoods: returns -1 on failure, else the number of narrowed variables
extern "C" int hc_prop_nogood(const int32_t *dom_in, int n_vars, const int32_t *lits, int n, int32_t *dom_out) {
  std::vector<int32_t> d(dom_in, dom_in + 2 * n_vars);
  std::vector<int> queue; std::vector<uint8_t> queued(n_vars, 0);
  HostCx cx{d.data(), &queue, &queued};
  bool ok = contract_nogood(cx, lits, n);
  memcpy(dom_out, d.data(), sizeof(int32_t) * 2 * n_vars);
  return ok ? (int)cx.props : -1;
}

extern "C" int32_t hc_sneg(int32_t a) { return sneg(a); }
extern "C" int32_t hc_sadd(int32_t a, int32_t b) { return sadd(a, b); }
extern "C" int32_t hc_smul(int32_t a, int32_t b) { return smul(a, b); }
