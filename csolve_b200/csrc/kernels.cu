// kernels.cu -- sm_100a kernels of the csolve search hot path.
//
//   k_search<false>   persistent depth-first search: one warp = one open search node at a
//                     time; the node's domain vector is staged in shared memory, the warp's
//                     DFS stack and the frontier pool live in HBM.     (src/csolve.c:398-476)
//   k_search<true>    batched frontier expansion: every warp takes a frontier frame, tries
//                     all of its values and appends the surviving children to the output
//                     frontier.                                       (src/csolve.c:105-152, 432-457)
//   k_rebalance       between time slices: idle warps receive the upper half of the
//                     shallowest splittable frame of a busy warp (interval bisection, as
//                     worker_spawn does in src/csolve.c:121-149).
//   k_propagate_batch parity hook: B independent node transitions.      (src/csolve.c:448-457)
//
// Inside a node, propagation to fixpoint (src/propagate.c:488-538) is a warp-synchronous
// worklist: the set of changed variables is a bitmask in shared memory; lanes stride over
// the watch list of each changed variable and contract one clause each; narrowed bounds
// are published with shared-memory atomicMax/atomicMin and recorded in the next round's
// mask with atomicOr. No tensor cores: the work is irregular int32 compare/min/max.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "kernels.cuh"
#include "contract.cuh"

namespace csolve_dev {

#define FULL 0xffffffffu

// development: CHK(cond, tag, value) reports a violated bound once per lane and stops the kernel (-DCSOLVE_BOUNDS)
#ifdef CSOLVE_BOUNDS
#define CHK(cond, tag, v) do { if (!(cond)) { printf("[bounds] %s: %lld (block %d thread %d)\n", tag, (long long)(v), blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define CHK(cond, tag, v) do { } while (0)
#endif

// resident blocks per SM the search kernel is compiled for (bounds the registers per thread)
// nodes between two polls of the control block (signal, time slice)
#define POLL_NODES 32

#ifndef CSOLVE_LOV_MIN_BLOCKS
#define CSOLVE_LOV_MIN_BLOCKS 4
#endif
#ifndef CSOLVE_MIN_BLOCKS
#define CSOLVE_MIN_BLOCKS 4
#endif

// ---- domain access of one warp: shared memory, lo/hi interleaved ---------------------------
struct WarpCx {
  int *d;            // shared: 2 * n_vars words
  unsigned *nxt;     // shared: mask of variables narrowed in this round
  unsigned props;    // per-lane PROPS partial
  const LinRel *linrel = nullptr;       // small tables (shared memory when they were staged, else the model's global arrays)
  const LinClause *lin = nullptr;
  const LinTerm *lin_term = nullptr;
  int *rlo = nullptr, *rhi = nullptr;   // learning: which record last raised lo / lowered hi of a variable in this node
  int cur = -1;                         // ... code of the record being contracted (>= 0 watch record, <= -2 nogood -2 - id)
  __device__ __forceinline__ Dom dom(int v) const {
    // one 8-byte load (the pair is 8-byte aligned): bounds published by other lanes are seen, never torn per word
    const unsigned long long q = *reinterpret_cast<const volatile unsigned long long *>(d + 2 * v);
    Dom r; r.lo = (int)(unsigned)q; r.hi = (int)(unsigned)(q >> 32);
    return r;
  }
  __device__ __forceinline__ void mark(int v) { atomicOr(&nxt[v >> 5], 1u << (v & 31)); }
  __device__ __forceinline__ void raise_lo(int v, int32_t lo) {
    if (atomicMax(&d[2 * v], lo) < lo) { mark(v); if (rlo) rlo[v] = cur; }
  }
  __device__ __forceinline__ void lower_hi(int v, int32_t hi) {
    if (atomicMin(&d[2 * v + 1], hi) > hi) { mark(v); if (rhi) rhi[v] = cur; }
  }
  __device__ __forceinline__ void count_prop() { props++; }
};

// one linear clause contracted by the whole warp (contract.cuh: lin_lane_load / lin_lane_apply)
__device__ __forceinline__ bool warp_contract_linear(WarpCx &cx, const DevModel &m, int c, int lane) {
  const int4 q0 = reinterpret_cast<const int4 *>(&cx.lin[c])[0];
  const int4 q1 = reinterpret_cast<const int4 *>(&cx.lin[c])[1];
  LinClause L; L.obj = q0.x; L.n_terms = q0.y; L.first = q0.z; L.konst = q0.w; L.clause = q1.x; L.pad0 = L.pad1 = L.pad2 = 0;
  const Dom O = cx.dom(L.obj);
  const LinLane t = lin_lane_load(cx, m, L, lane);
  const int32_t SL = __reduce_add_sync(FULL, t.tlo) + L.konst, SH = __reduce_add_sync(FULL, t.thi) + L.konst;
  return lin_lane_apply(cx, L, t, lane, SL, SH, O);
}

// per-warp shared memory carve-up
struct LinTables { const LinRel *linrel; const LinClause *lin; const LinTerm *lin_term; const int2 *dense_form; };

struct WarpSmem {
  LinTables lt;    // small tables of the block (set after stage_table)
  int *d;          // 2V: the node being propagated
  int *p;          // 2V: domains of the top frame (state before this level's assignment)
  unsigned *cur;   // mask_words: variables whose watchers run in this round
  unsigned *nxt;   // mask_words
  unsigned *amask; // mask_words: variables assigned on the path to the top frame
  // learning only (src/conflict.c): per-variable reason records of this node, analysis scratch
  int *rlo, *rhi;  // V each
  unsigned *seen;  // mask_words
  int *work;       // 2V: stack of record codes still to expand
  int *nlits;      // NG_MAX_LITS
};

__device__ __host__ __forceinline__ int warp_smem_words(const DevModel &m, bool learn = false) {
  return 4 * m.n_vars + 3 * m.mask_words + (learn ? 4 * m.n_vars + m.mask_words + NG_MAX_LITS : 0);
}

__device__ __forceinline__ WarpSmem carve(const DevModel &m, int *base) {
  WarpSmem s;
  s.d = base;
  s.p = base + 2 * m.n_vars;
  s.cur = (unsigned *)(base + 4 * m.n_vars);
  s.nxt = s.cur + m.mask_words;
  s.amask = s.nxt + m.mask_words;
  s.rlo = (int *)(s.amask + m.mask_words);       // only valid when the region was sized with learn = true
  s.rhi = s.rlo + m.n_vars;
  s.seen = (unsigned *)(s.rhi + m.n_vars);
  s.work = (int *)(s.seen + m.mask_words);
  s.nlits = s.work + 2 * m.n_vars;
  return s;
}

// ---- propagation to fixpoint for the node staged in s.d --------------------------------------
// precondition: s.cur holds the initial worklist, s.nxt is zero, warp converged after __syncwarp.
// wrec / wptr: the compiled watch records (shared memory when the table was staged, else global).
// returns false when the node failed (PROP_ERROR).
struct FailInfo { int rec; int var; };   // learning: record being contracted when the node failed / variable found empty

// LIN: the model has linear clauses that the whole warp contracts (a template parameter, not a run-time test: the
// test and the extra live register cost the 3-SAT kernel 5..10 % when they sat in the common code)
template <bool LEARN = false, bool LIN = false>
__device__ __forceinline__ bool warp_fixpoint(const DevModel &m, WarpSmem &s, const int4 *wrec, const int *wptr,
                                              int lane, unsigned &props, unsigned &visits, int *gprio = nullptr,
                                              const NogoodPool *ng = nullptr, FailInfo *fi = nullptr) {
  WarpCx cx;
  cx.d = s.d; cx.props = 0;
  cx.linrel = s.lt.linrel; cx.lin = s.lt.lin; cx.lin_term = s.lt.lin_term;
  if (LEARN) { cx.rlo = s.rlo; cx.rhi = s.rhi; }
  if (!LEARN && m.dense) {
    // Few clauses (at most 32): every round contracts ALL of them, lane c clause c, then the linear clauses with the
    // whole warp -- no worklists, no per-variable record walks. Same contractors, same greatest fixpoint;
    // a wcet fixpoint went from ~80 clause visits of ~100 warp instructions each to a handful of rounds of ~150.
    cx.nxt = s.nxt;
    for (;;) {
      bool bad = false;
      if (lane < m.n_clauses) {
        const int2 f = s.lt.dense_form[lane];
        if (f.x == DF_LINREL) {
          const int4 q0 = reinterpret_cast<const int4 *>(&cx.linrel[f.y])[0];
          const int4 q1 = reinterpret_cast<const int4 *>(&cx.linrel[f.y])[1];
          LinRel R; R.rel = q0.x; R.n = q0.y; R.konst = q0.z; R.clause = q0.w; R.v[0] = q1.x; R.v[1] = q1.y; R.v[2] = q1.z; R.v[3] = q1.w;
          bad = !contract_linrel(cx, R);
        } else if (f.x == DF_CLAUSE) {
          const int4 q = __ldg(reinterpret_cast<const int4 *>(&m.clause[lane]));
          ClauseRec rec; rec.kind = q.x; rec.a = q.y; rec.b = q.z; rec.c = q.w;
          bad = !contract_clause(cx, m, rec);
        }
        visits++;
      }
      __syncwarp();
      for (int c = 0; LIN && c < m.n_lin; c++) {
        if (!warp_contract_linear(cx, m, c, lane)) bad = true;
        visits++;
      }
      __syncwarp();
      // bounds published by different lanes may have crossed
      for (int v = lane; v < m.n_vars; v += 32) { const Dom X = cx.dom(v); if (X.lo > X.hi) bad = true; }
      if (__any_sync(FULL, bad)) { props += cx.props; return false; }
      unsigned ch = 0;
      for (int w = lane; w < m.mask_words; w += 32) { ch |= s.nxt[w]; s.nxt[w] = 0; }
      __syncwarp();
      if (!__any_sync(FULL, ch != 0u)) break;
    }
    props += cx.props;
    return true;
  }
  bool failed = false;
  int fail_var = -1, fail_rec = -1, empty_var = -1;
  for (;;) {
    cx.nxt = s.nxt;
    bool any = false;
    // The changed variables of this round, 1024 at a time: every lane loads one word of the mask, a ballot finds the
    // non-empty words, and the variables are handed out TWO at a time -- lanes 0..15 contract the watch records of one,
    // lanes 16..31 those of the other (watch lists are short: 13 records on average for 3-SAT, 3(N-1) for N-queens);
    // a single variable gets the whole warp. The fixpoint does not depend on the order.
    for (int wb = 0; wb < m.mask_words; wb += 32) {
      const unsigned myw = wb + lane < m.mask_words ? s.cur[wb + lane] : 0u;
      unsigned nz = __ballot_sync(FULL, myw != 0u);
      if (nz) any = true;
      unsigned bits = 0;
      int wcur = 0;
      unsigned lin_hit = 0;               // linear clauses that watch a variable of this chunk (per lane)
      for (;;) {
        int xA = -1, xB = -1;
        if (bits == 0u && nz != 0u) { wcur = __ffs((int)nz) - 1; nz &= nz - 1; bits = __shfl_sync(FULL, myw, wcur); }
        if (bits == 0u) break;
        xA = ((wb + wcur) << 5) + __ffs((int)bits) - 1;
        bits &= bits - 1;
        if (!LEARN) {
          if (bits == 0u && nz != 0u) { wcur = __ffs((int)nz) - 1; nz &= nz - 1; bits = __shfl_sync(FULL, myw, wcur); }
          if (bits != 0u) { xB = ((wb + wcur) << 5) + __ffs((int)bits) - 1; bits &= bits - 1; }
        }
        const bool two = xB >= 0;
        const int x = (two && lane >= 16) ? xB : xA;
        const int first = two ? (lane & 15) : lane, stride = two ? 16 : 32;
        // snapshot of the dequeued variable; bounds published by different lanes may have
        // crossed since it was queued: an empty domain is a failure
        const Dom X = cx.dom(x);
        if (X.lo > X.hi) { failed = true; empty_var = x; }
        const int b = wptr[x], e = wptr[x + 1];
        for (int i = b + first; i < e; i += stride) {
          const int4 q = wrec[i];
          WatchRec rec; rec.w0 = (uint32_t)q.x; rec.c[0] = q.y; rec.c[1] = q.z; rec.c[2] = q.w;
          if (LEARN) cx.cur = i;
          if (LIN && wrec_is_linear(rec.w0)) lin_hit |= 1u << wrec_arg(rec.w0);   // contracted below, by the warp
          else if (!contract_watch(cx, m, x, X, rec)) { failed = true; fail_var = x; fail_rec = i; }
          visits++;
        }
        if (LEARN) {
          // learned nogoods that mention x (propagate_confl, src/propagate.c:459-471)
          const int cnt = min(*reinterpret_cast<volatile int *>(&ng->watch_n[x]), ng->cap_w);
          for (int i = lane; i < cnt; i += 32) {
            const int id = __ldcg(&ng->watch[(size_t)x * ng->cap_w + i]);
            if (id < 0) continue;
            cx.cur = -2 - id;
            if (!contract_nogood(cx, ng->lits + __ldcg(&ng->start[id]), __ldcg(&ng->len[id]))) {
              failed = true; fail_var = x; fail_rec = -2 - id;
            }
            visits++;
          }
        }
        __syncwarp();
      }
      if (LIN) {
        // each linear clause that watches one of the round's variables: once, by the whole warp
        unsigned dirty = __reduce_or_sync(FULL, lin_hit);
        while (dirty) {
          const int c = __ffs((int)dirty) - 1;
          dirty &= dirty - 1;
          if (!warp_contract_linear(cx, m, c, lane)) { failed = true; fail_var = cx.lin[c].obj; }
          visits++;
        }
      }
    }
    __syncwarp();
    const unsigned fmask = __ballot_sync(FULL, failed);
    if (fmask) {
      // prefer-failing: the variable whose clauses failed gains priority (propagate_term_recurse, src/propagate.c:44-54)
      if (gprio != nullptr && fail_var >= 0) atomicAdd(&gprio[fail_var], 1);
      if (LEARN && fi != nullptr) {
        const int src = __ffs(fmask) - 1;
        fi->rec = __shfl_sync(FULL, fail_rec, src);
        fi->var = __shfl_sync(FULL, empty_var, src);
      }
      props += cx.props;
      return false;
    }
    if (!any) break;
    // next round: cur <- nxt, nxt <- 0
    unsigned *t = s.cur; s.cur = s.nxt; s.nxt = t;
    for (int w = lane; w < m.mask_words; w += 32) s.nxt[w] = 0;
    __syncwarp();
  }
  props += cx.props;
  return true;
}

// ---- conflict analysis (src/conflict.c:327-362), executed by lane 0 after a failed node ---------------------
// The nogood is the set of (variable = value) facts the failure rests on: the decision of this node and the
// values of variables that were already a value before this node (src/conflict.c:188-197: "bound at a lower
// level" terminals are taken as they are); variables narrowed inside this node are replaced by the variables of
// the record that narrowed them (src/conflict.c:289-307). Only 0/1 values can be recorded
// (src/conflict.c:173-179): anything else abandons the analysis, as in the reference.
// Differences to the reference, none of which affects results: no back-jump (the search stays chronological,
// the nogood prunes later nodes), a variable that became a value at an earlier level is always taken as a
// literal (the reference expands the reasons of the failing variable itself one level further).
struct Analysis {
  const DevModel &m; const WarpSmem &s; const int4 *wrec; const int *wptr; const NogoodPool &ng;
  int decision_var, n_lits, n_work; bool ok;

  __device__ void visit(int y) {
    if (!ok) return;
    if (s.seen[y >> 5] & (1u << (y & 31))) return;
    s.seen[y >> 5] |= 1u << (y & 31);
    const int dl = s.d[2 * y], dh = s.d[2 * y + 1], pl = s.p[2 * y], ph = s.p[2 * y + 1];
    if (y == decision_var || (dl == pl && dh == ph)) {
      // the decision, or a value fixed before this node
      const int lo = y == decision_var ? dl : pl, hi = y == decision_var ? dh : ph;
      if (lo != hi || lo < 0 || lo > 1 || n_lits >= NG_MAX_LITS) { ok = false; return; }
      s.nlits[n_lits++] = (y << 1) | lo;
      return;
    }
    // narrowed in this node: its bounds must be explainable by records alone
    if (__ldg(&m.root_dom[2 * y]) < 0 || __ldg(&m.root_dom[2 * y + 1]) > 1 || pl != __ldg(&m.root_dom[2 * y]) || ph != __ldg(&m.root_dom[2 * y + 1])) { ok = false; return; }
    if (dl != pl) push(s.rlo[y]);
    if (dh != ph) push(s.rhi[y]);
  }
  __device__ void push(int code) {
    if (code == -1 || n_work >= 2 * m.n_vars) { ok = false; return; }    // narrowed without a record (objective tightening)
    s.work[n_work++] = code;
  }
  __device__ int owner_of(int r) const {      // variable whose watch list holds record r
    int a = 0, b = m.n_vars;
    while (b - a > 1) { const int c = (a + b) >> 1; if (wptr[c] <= r) a = c; else b = c; }
    return a;
  }
  __device__ void expand(int code) {
    if (code <= -2) {
      const int id = -2 - code;
      const int st = __ldcg(&ng.start[id]), n = __ldcg(&ng.len[id]);
      for (int k = 0; k < n && ok; k++) visit(__ldcg(&ng.lits[st + k]) >> 1);
      return;
    }
    const int4 q = wrec[code];
    const uint32_t kind = wrec_kind((uint32_t)q.x);
    if (kind == WK_NE_VV) { visit(owner_of(code)); visit(wrec_arg((uint32_t)q.x)); }
    else if (kind == WK_NE_VC) { visit(owner_of(code)); }
    else if (kind == WK_LITS) {
      const int n = wrec_n((uint32_t)q.x);
      visit(q.y >> 1); if (n > 1) visit(q.z >> 1); if (n > 2) visit(q.w >> 1);
    } else {
      const int wn = wrec_n((uint32_t)q.x);
      const int ci = wn == 2 ? m.lin[wrec_arg((uint32_t)q.x)].clause : (wn == 3 ? m.linrel[wrec_arg((uint32_t)q.x)].clause : wrec_arg((uint32_t)q.x));
      const ClauseRec c = m.clause[ci];
      for (int j = c.a; j <= c.b && ok; j++) if (m.node_op[j] == CSOLVE_OP_VAR) visit(m.node_l[j]);
    }
  }
};

// Returns the number of literals of the nogood that was stored (they stay in s.nlits), 0 when none was.
__device__ __noinline__ int learn_nogood(const DevModel &m, const WarpSmem &s, const int4 *wrec, const int *wptr,
                                         const NogoodPool &ng, int decision_var, FailInfo fi) {
  atomicAdd(&ng.counters[2], 1);
  for (int w = 0; w < m.mask_words; w++) s.seen[w] = 0;
  Analysis an{m, s, wrec, wptr, ng, decision_var, 0, 0, true};
  if (fi.var >= 0) {
    // a variable whose bounds crossed: explained by the records that moved its two bounds
    const int y = fi.var;
    s.seen[y >> 5] |= 1u << (y & 31);
    if (__ldg(&m.root_dom[2 * y]) < 0 || __ldg(&m.root_dom[2 * y + 1]) > 1) an.ok = false;
    if (y == decision_var) an.ok = false;     // rare (needs two lanes racing on the decision variable): not analysed
    if (an.ok) { if (s.d[2 * y] != s.p[2 * y]) an.push(s.rlo[y]); if (s.d[2 * y + 1] != s.p[2 * y + 1]) an.push(s.rhi[y]); }
  } else if (fi.rec != -1) {
    an.push(fi.rec);
  } else {
    an.ok = false;
  }
  while (an.ok && an.n_work > 0) an.expand(s.work[--an.n_work]);
  if (!an.ok || an.n_lits == 0) { atomicAdd(&ng.counters[3], 1); return 0; }
  const int n = an.n_lits;
  const int st = atomicAdd(&ng.counters[1], n);
  if (st + n > ng.cap_lits) { atomicAdd(&ng.counters[4], 1); return 0; }
  const int id = atomicAdd(&ng.counters[0], 1);
  if (id >= ng.cap_ng) { atomicAdd(&ng.counters[4], 1); return 0; }
  for (int k = 0; k < n; k++) ng.lits[st + k] = s.nlits[k];
  ng.start[id] = st; ng.len[id] = n;
  __threadfence();
  // the nogood is watched by every variable it mentions (src/conflict.c:354-358)
  for (int k = 0; k < n; k++) {
    const int v = s.nlits[k] >> 1;
    const int slot = atomicAdd(&ng.watch_n[v], 1);
    if (slot < ng.cap_w) __stcg(&ng.watch[(size_t)v * ng.cap_w + slot], id);
  }
  return n;
}

// stage the watch-record table and the small linear tables into shared memory (whole block); returns the pointers to use
__device__ __forceinline__ void stage_table(const DevModel &m, int *smem, const int4 *&wrec, const int *&wptr, LinTables &lt) {
  lt.linrel = m.linrel; lt.lin = m.lin; lt.lin_term = m.lin_term; lt.dense_form = reinterpret_cast<const int2 *>(m.dense_form);
  if (m.table_smem_bytes > 0) {
    int4 *dst = reinterpret_cast<int4 *>(smem);
    const int4 *src = reinterpret_cast<const int4 *>(m.wrec);
    for (int i = threadIdx.x; i < m.n_wrec; i += blockDim.x) dst[i] = __ldg(&src[i]);
    int *pdst = smem + 4 * m.n_wrec;
    for (int i = threadIdx.x; i <= m.n_vars; i += blockDim.x) pdst[i] = __ldg(&m.wrec_ptr[i]);
    int *t = smem + ((4 * m.n_wrec + m.n_vars + 1 + 3) & ~3);
    const int n0 = 8 * m.n_linrel, n1 = 8 * m.n_lin, n2 = 2 * m.n_lin_term;
    for (int i = threadIdx.x; i < n0; i += blockDim.x) t[i] = __ldg(reinterpret_cast<const int *>(m.linrel) + i);
    for (int i = threadIdx.x; i < n1; i += blockDim.x) t[n0 + i] = __ldg(reinterpret_cast<const int *>(m.lin) + i);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) t[n0 + n1 + i] = __ldg(reinterpret_cast<const int *>(m.lin_term) + i);
    if (m.dense) {
      for (int i = threadIdx.x; i < 2 * m.n_clauses; i += blockDim.x) t[n0 + n1 + n2 + i] = __ldg(reinterpret_cast<const int *>(m.dense_form) + i);
      lt.dense_form = reinterpret_cast<const int2 *>(t + n0 + n1 + n2);
    }
    lt.linrel = reinterpret_cast<const LinRel *>(t);
    lt.lin = reinterpret_cast<const LinClause *>(t + n0);
    lt.lin_term = reinterpret_cast<const LinTerm *>(t + n0 + n1);
    __syncthreads();
    wrec = dst; wptr = pdst;
  } else {
    wrec = reinterpret_cast<const int4 *>(m.wrec); wptr = m.wrec_ptr;
  }
}

// leaf test: every clause evaluates to true (src/csolve.c:226, src/eval.c:221-245)
__device__ bool warp_all_true(const DevModel &m, WarpSmem &s, int lane) {
  WarpCx cx; cx.d = s.d; cx.nxt = s.nxt; cx.props = 0;
  cx.linrel = s.lt.linrel; cx.lin = s.lt.lin; cx.lin_term = s.lt.lin_term;
  bool ok = true;
  for (int c = lane; c < m.n_clauses; c += 32) {
    const int4 q = __ldg(reinterpret_cast<const int4 *>(&m.clause[c]));
    ClauseRec rec; rec.kind = q.x; rec.a = q.y; rec.b = q.z; rec.c = q.w;
    if (!clause_is_true(cx, m, rec)) ok = false;
  }
  return __all_sync(FULL, ok);
}

// ---- branching variable for the next level (src/strategy.c:79-121) -----------------------------
// ORDER_NONE is static (priority order); the other orders look at the node's domains.
// Ties: higher parse-time priority, then lower index.
__device__ int warp_select_var(const DevModel &m, const WarpSmem &s, int lane, int order, int next_level, int cur_var,
                               const int *gprio = nullptr) {
  if (order == CSOLVE_ORDER_NONE && gprio == nullptr) return __ldg(&m.order[next_level]);
  unsigned long long bestk = ~0ull;
  int bestv = 0x7fffffff;
  for (int v = lane; v < m.n_vars; v += 32) {
    if (v == cur_var || (s.amask[v >> 5] & (1u << (v & 31)))) continue;   // already has a level
    const int lo = s.d[2 * v], hi = s.d[2 * v + 1];
    unsigned primary;
    switch (order) {
    case CSOLVE_ORDER_SMALLEST_DOMAIN: primary = (unsigned)hi - (unsigned)lo; break;
    case CSOLVE_ORDER_LARGEST_DOMAIN:  primary = ~((unsigned)hi - (unsigned)lo); break;
    case CSOLVE_ORDER_SMALLEST_VALUE:  primary = (unsigned)lo ^ 0x80000000u; break;
    case CSOLVE_ORDER_LARGEST_VALUE:   primary = ~((unsigned)hi ^ 0x80000000u); break;
    default:                           primary = 0; break;      // ORDER_NONE: priority only
    }
    const int pr = gprio != nullptr ? __ldcg(&gprio[v]) : __ldg(&m.prio[v]);
    const unsigned secondary = ~((unsigned)pr ^ 0x80000000u);
    const unsigned long long k = ((unsigned long long)primary << 32) | secondary;
    if (k < bestk) { bestk = k; bestv = v; }   // ascending v per lane: first best wins
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ok = __shfl_xor_sync(FULL, bestk, o);
    const int ov = __shfl_xor_sync(FULL, bestv, o);
    if (ok < bestk || (ok == bestk && ov < bestv)) { bestk = ok; bestv = ov; }
  }
  return bestv;
}

// ---- frame <-> shared memory ----------------------------------------------------------------------
__device__ __forceinline__ void load_domains(const DevModel &m, const int *frame, int *d, int lane) {
  const int2 *src = reinterpret_cast<const int2 *>(frame + frame_dom_offset(m.mask_words));
  int2 *dst = reinterpret_cast<int2 *>(d);
  for (int v = lane; v < m.n_vars; v += 32) dst[v] = __ldcg(&src[v]);
}
__device__ __forceinline__ void store_domains(const DevModel &m, int *frame, const int *d, int lane) {
  int2 *dst = reinterpret_cast<int2 *>(frame + frame_dom_offset(m.mask_words));
  const int2 *src = reinterpret_cast<const int2 *>(d);
  for (int v = lane; v < m.n_vars; v += 32) __stcg(&dst[v], src[v]);
}

__device__ __forceinline__ unsigned mix_hash(unsigned h, unsigned a, unsigned b) {
  h ^= a * 0x9E3779B1u; h = (h << 13) | (h >> 19); h *= 0x85EBCA77u;
  h ^= b * 0xC2B2AE3Du; h = (h << 15) | (h >> 17); h *= 0x27D4EB2Fu;
  return h ^ (h >> 16);
}

// ---- parity instrumentation (SAMPLE instances of the search kernels only) -------------------------------
// A node is identified by (parent domains, variable, value); whether it is recorded depends only on that identity,
// so the recorded set of a deterministic tree does not depend on how the warps happened to traverse it.
__device__ __forceinline__ unsigned sample_dom_hash(const int *p, int V, int lane) {
  unsigned h = 0;
  for (int v = lane; v < V; v += 32) h ^= mix_hash(0x9E3779B9u + (unsigned)v, (unsigned)p[2 * v], (unsigned)p[2 * v + 1]);
  return __reduce_xor_sync(FULL, h);
}
// failed nodes are thinned further (1 in sample_fkeep of the hits): most nodes of a search fail
__device__ __forceinline__ bool sample_hit(const SearchArgs &a, unsigned hpar, int var, int val, bool failed) {
  const unsigned key = mix_hash(hpar, (unsigned)var, (unsigned)val);
  if (key % a.sample_mod != 0u) return false;
  return !failed || (key / a.sample_mod) % a.sample_fkeep == 0u;
}
// reserves a record and writes its header; returns the record (nullptr: buffer full), same on every lane
__device__ __forceinline__ int *sample_begin(const SearchArgs &a, int lane, int flags, int var, int val, int best) {
  int slot = 0;
  if (lane == 0) slot = atomicAdd(a.sample_n, 1);
  slot = __shfl_sync(FULL, slot, 0);
  if (slot >= a.sample_cap) return nullptr;
  int *r = a.sample_rec + (size_t)slot * sample_words(a.m.n_vars);
  if (lane == 0) { r[0] = flags; r[1] = var; r[2] = val; r[3] = best; }
  return r;
}

// write the child frame for level `level + 1` (domains = the node's post-fixpoint state in s.d)
__device__ __forceinline__ void write_child_frame(const DevModel &m, const WarpSmem &s, int *g, int lane,
                                                  int nv, int level1, int best, unsigned hash, int parent_var) {
  if (lane == 0) {
    const int lo = s.d[2 * nv], hi = s.d[2 * nv + 1];
    __stcg(reinterpret_cast<int4 *>(g), make_int4(nv, 0, (int)((unsigned)hi - (unsigned)lo), lo));
    __stcg(reinterpret_cast<int4 *>(g) + 1, make_int4(hi, level1, best, (int)hash));
  }
  for (int w = lane; w < m.mask_words; w += 32) {
    unsigned mk = s.amask[w];
    if (parent_var >= 0 && (parent_var >> 5) == w) mk |= 1u << (parent_var & 31);
    __stcg(&g[FR_MASK + w], (int)mk);
  }
  store_domains(m, g, s.d, lane);
}

// record an accepted leaf (src/csolve.c:222-244): the assignment plus the objective key
__device__ __forceinline__ void store_solution(const SearchArgs &a, const WarpSmem &s, int lane, int key) {
  int slot = 0;
  if (lane == 0) slot = atomicAdd(&a.ctl->n_stored, 1);
  slot = __shfl_sync(FULL, slot, 0);
  // MIN / MAX: the buffer is a ring -- every accepted leaf improves on all earlier ones, so the optimum's witness is
  // among the most recent entries however many incumbents there were
  if (a.m.obj_var >= 0 && a.max_solutions > 0) slot = (int)((unsigned)slot % (unsigned)a.max_solutions);
  if (slot < a.max_solutions) {
    int *dst = a.solbuf + (size_t)slot * (a.m.n_vars + 1);
    for (int v = lane; v < a.m.n_vars; v += 32) dst[v] = s.d[2 * v];
    if (lane == 0) dst[a.m.n_vars] = key;
  }
  __syncwarp();       // reconverge after the predicated stores (see k_search_lov)
}


// ---- shared frame pool (depth-first phase) ---------------------------------------------------------------
// Work comes from two places.
// (1) The expanded root frontier (the first n_initial pool entries): static, claimed in CHUNKS with one atomicAdd --
//     guided self-scheduling: a chunk is 1/4 of an even share of what is left, at most 32 frames, so the atomics are
//     few while the frontier is long and single frames are handed out when it runs out; with a rank partition every
//     lane looks at one frame of the chunk and a ballot keeps the ones whose path hash maps to this rank.
// (2) Behind it the pool is a ring of DONATED frames run as a ticket queue: a warp that ran dry takes a ticket
//     (fetch-and-add on ctl->item_next) and waits for ITS slot's ready flag -- a private address, so thousands of
//     waiters do not contend; a busy warp that sees unserved tickets (item_next > item_count) takes the next one to
//     serve (fetch-and-add on ctl->item_count), writes the frame into that slot and publishes it. Frames are the
//     shallowest untried work of the donor's stack (the bisection worker_spawn does with fork(),
//     src/csolve.c:121-149). Neither side retries on a shared line: a head claimed with atomicCAS serialised the
//     hand-off at about one frame per microsecond (measured), which was the whole tail of a search.
//     ready[slot]: 0 free, 1 frame present, -1 the ticket's holder left at a slice end (the donor that draws this
//     ticket takes another one; k_rebalance clears what is left between slices).
// Returns the claimed slot, or -1 when the warp has to leave the kernel (slice end / stop / nothing left anywhere).
struct Claim {
  int base; unsigned mask;     // claimed, not yet searched frames of the root frontier (bit b = frame base + b)
  bool drained;                // the root frontier has been handed out completely
};

// DEMAND: the waiting loop carries the request to the peer ranks (see there). Compiled out of the lane-owns-variable
// kernel: its trees are dealt evenly by the path hash, and the mere presence of that code in the waiting path moved
// the compiler's choices in the node loop (16-queens: 16.22 -> 15.88 G nodes/s on one GPU, A/B on the same box).
template <bool DEMAND>
__device__ __forceinline__ int claim_frame(const SearchArgs &a, int lane, bool &hungry, Claim &cl, int *blk_hungry, long long t0) {
  if (cl.mask) {
    const int b = __ffs((int)cl.mask) - 1;
    cl.mask &= cl.mask - 1;
    return cl.base + b;
  }
  SearchCtl *ctl = a.ctl;
  const int fw = a.m.frame_words;
  const int ring = a.pool_cap - a.n_initial;       // donated frames live in the slots behind the root frontier
  int slot = -1;
  int tslot = -1;                                  // lane 0: slot of the ticket this warp holds
  unsigned spins = 0, nap = 200;
  for (;;) {
    // code: >= 0 ring slot claimed, -1 leave the kernel, -2 nothing yet (retry), -3 the root frontier is drained
    int code = -2, it = 0, n = 0;
    if (lane == 0) {
      if (tslot < 0 && (spins & 7u) == 0u && *reinterpret_cast<volatile int *>(&ctl->signal) != SIG_RUN) {
        code = -1;
      } else if (!cl.drained) {
        // the frontier's claim counter may live on another GPU (csolve_gpu_comm): system-scope atomic over NVLink
        const int nx = *reinterpret_cast<volatile int *>(&a.front_ctl->init_next);
        if (nx < a.n_initial) {
          n = min(32, max(min(a.part_count, 32), (a.n_initial - nx) / (a.total_warps * 4)));
          it = a.n_peers > 0 ? atomicAdd_system(&a.front_ctl->init_next, n) : atomicAdd(&a.front_ctl->init_next, n);
          if (it >= a.n_initial) n = 0;
        }
        if (n == 0) code = -3;
      } else {
        if (tslot < 0) {
          const int ticket = atomicAdd(&ctl->item_next, 1);
          tslot = a.n_initial + (int)((unsigned)ticket % (unsigned)ring);
          if (!hungry) { hungry = true; atomicAdd(&ctl->hungry, 1); atomicAdd(blk_hungry, 1); }
        }
        if (*reinterpret_cast<volatile int *>(&a.ready[tslot]) == 1) {
          code = tslot;
        } else {
          bool leave = false;
          if ((spins & 3u) == 3u) {
            if (a.n_peers > 0 && *reinterpret_cast<volatile int *>(&a.comm->stop_epoch) == a.epoch) atomicMax(&ctl->signal, SIG_STOP);
            const int n_hungry = *reinterpret_cast<volatile int *>(&ctl->hungry);
            // Ranks of a comm: when this GPU is running out of donors (three quarters of its warps are waiting) a peer
            // is asked to serve some of its tickets over NVLink -- while this kernel keeps running. (Asking earlier,
            // whenever an eighth of the warps held unserved tickets, cost 16-queens on 8 GPUs 17 %: in the second half
            // of such a search every rank is short of frames now and then, and a hand-off over NVLink stalls the donor
            // for several round trips.) The request is a LEVEL in the peer's block (stale requests do not pile up),
            // the peers take turns, and a busy warp over there serves a ticket like a local donor (donation_target).
            if (DEMAND && a.peer_demand && n_hungry * 4 >= a.n_warps * 3 && (spins & 31u) == 3u)
              atomicMax_system(&a.peer_comm[(a.rank + 1 + (int)((unsigned)(tslot + (int)(spins >> 5)) % (unsigned)a.n_peers)) % a.world]->demand[a.rank], n_hungry / a.n_peers + 1);
            if (*reinterpret_cast<volatile int *>(&ctl->signal) != SIG_RUN) leave = true;
            else if (n_hungry >= a.n_warps || clock64() - t0 > a.slice_cycles) {
              // every warp is waiting: nothing left anywhere -- or the time slice is over. Waiters watch the clock too:
              // if fewer blocks are resident than the launch assumed (another context on the device, a profiler),
              // `hungry` can never reach n_warps and the resident warps would wait for ever.
              atomicMax(&ctl->signal, SIG_SLICE_END);
              leave = true;
            }
          }
          if (leave) {
            // give the ticket back -- unless its frame arrived in the meantime
            code = atomicCAS(&a.ready[tslot], 0, -1) == 1 ? tslot : -1;
          } else {
            __nanosleep(nap);
            nap = min(nap * 2u, 3200u);
          }
        }
      }
      spins++;
    }
    code = __shfl_sync(FULL, code, 0);
    if (code == -1) break;
    if (code >= 0) { slot = code; break; }
    if (code == -3) { cl.drained = true; continue; }
    n = __shfl_sync(FULL, n, 0);
    if (n > 0) {
      it = __shfl_sync(FULL, it, 0);
      const int idx = it + lane;
      bool mine = lane < n && idx < a.n_initial;
      if (mine && a.part_count > 1)
        mine = (unsigned)__ldcg(&a.front_pool[(size_t)idx * fw + 7]) % (unsigned)a.part_count == (unsigned)a.part_rank;
      const unsigned mk = __ballot_sync(FULL, mine);
      if (mk) {
        cl.base = it;
        cl.mask = mk & (mk - 1);
        slot = it + __ffs((int)mk) - 1;
        break;
      }
    }
  }
  if (lane == 0 && slot >= 0 && hungry) { hungry = false; atomicSub(&ctl->hungry, 1); atomicSub(blk_hungry, 1); }
  return slot;
}

// should this warp donate now, and into whose ring?  (lane 0 reads the control block; every lane gets the answer)
// -1: no. a.rank: this rank's own ring (tickets taken by its waiting warps that no donor has picked up yet).
// another rank r: rank r ran dry and asked for frames (CommBlock::demand, csolve_gpu_comm) -- one unit of its budget
// is taken here; its tickets are served over NVLink exactly like local ones.
__device__ __forceinline__ int donation_target(const SearchArgs &a, int lane) {
  int tgt = -1;
  if (lane == 0) {
    if (*reinterpret_cast<volatile int *>(&a.ctl->item_next) - *reinterpret_cast<volatile int *>(&a.ctl->item_count) > 0) {
      tgt = a.rank;
    } else if (a.n_peers > 0) {
      for (int r = 0; r < a.world; r++) {
        if (r == a.rank || *reinterpret_cast<volatile int *>(&a.comm->demand[r]) <= 0) continue;
        if (atomicSub(&a.comm->demand[r], 1) <= 0) continue;
        // announce the visit, then make sure the ring is (still) open: its owner waits for announced visitors
        // before it touches the ring again (k_comm_state)
        atomicAdd_system(&a.peer_comm[r]->inflight, 1);
        if (*reinterpret_cast<volatile int *>(&a.peer_comm[r]->ring_open) == a.epoch) { tgt = r; break; }
        atomicSub_system(&a.peer_comm[r]->inflight, 1);
      }
    }
  }
  return __shfl_sync(FULL, tgt, 0);
}
// end of a visit to another rank's ring (whether or not a frame was delivered)
__device__ __forceinline__ void donation_done(const SearchArgs &a, int lane, int tgt) {
  if (tgt != a.rank && lane == 0) { __threadfence_system(); atomicSub_system(&a.peer_comm[tgt]->inflight, 1); }
  __syncwarp();
}

__device__ __forceinline__ int *ring_frame(const SearchArgs &a, int tgt, int slot) {
  return (tgt == a.rank ? a.pool : a.peer_pool[tgt]) + (size_t)slot * a.m.frame_words;
}

// take the next ticket of rank tgt's ring to serve (lane 0); its slot must be free: wait while the previous lap's
// frame is still being copied out, skip tickets whose holder left
__device__ __forceinline__ int reserve_slot(const SearchArgs &a, int lane, int tgt) {
  int s = 0;
  if (lane == 0) {
    const bool local = tgt == a.rank;
    SearchCtl *ctl = local ? a.ctl : a.peer_ctl[tgt];
    int *ready = local ? a.ready : a.peer_ready[tgt];
    for (;;) {
      const int t = local ? atomicAdd(&ctl->item_count, 1) : atomicAdd_system(&ctl->item_count, 1);
      s = a.n_initial + (int)((unsigned)t % (unsigned)(a.pool_cap - a.n_initial));
      int r;
      while ((r = *reinterpret_cast<volatile int *>(&ready[s])) == 1) __nanosleep(100);
      if (r == 0) break;
      __stcg(&ready[s], 0);       // abandoned ticket: nobody waits here any more
    }
  }
  return __shfl_sync(FULL, s, 0);
}

// make the frame written to `slot` visible to the ticket's holder (lane 0; every lane gets the answer).
// false: the holder left while the frame was being written -- the caller donates again to another ticket.
// Another rank's ring: that rank counts as active from here on (it may have been idle, its host polls for this).
__device__ __forceinline__ bool publish_slot(const SearchArgs &a, int lane, int tgt, int slot) {
  int ok = 1;
  if (lane == 0) {
    if (tgt == a.rank) {
      __threadfence();
      ok = atomicCAS(&a.ready[slot], 0, 1) == 0;
      if (!ok) __stcg(&a.ready[slot], 0);
    } else {
      if (atomicExch_system(&a.peer_comm[tgt]->busy_epoch, a.epoch) != a.epoch) atomicAdd_system(&a.peer_comm[0]->active64, 1ull);
      __threadfence_system();
      ok = atomicCAS_system(&a.peer_ready[tgt][slot], 0, 1) == 0;
      if (!ok) __stcg(&a.peer_ready[tgt][slot], 0);
    }
  }
  return __shfl_sync(FULL, ok, 0) != 0;
}

// restarts (src/csolve.c:264-276): the warps report their failed nodes at every poll; the slice ends for everybody when
// the device-wide count passes the limit (the host restarts the search from the root, src/csolve.c:380-385).
// `reported`: this warp's failed nodes already added to the count. Returns true when the limit is reached.
__device__ __forceinline__ bool restart_due(const SearchArgs &a, int lane, unsigned long long cuts, unsigned long long &reported) {
  int due = 0;
  if (lane == 0) {
    const int delta = (int)(cuts - reported);
    const int before = delta > 0 ? atomicAdd(&a.ctl->fails, delta) : *reinterpret_cast<volatile int *>(&a.ctl->fails);
    due = before + delta > a.fail_limit;
    if (due) atomicMax(&a.ctl->signal, SIG_SLICE_END);
  }
  reported = cuts;
  return __shfl_sync(FULL, due, 0) != 0;
}

// frame of a claimed slot: the first n_initial slots are the root frontier (possibly on another GPU), the rest the ring
__device__ __forceinline__ const int *claimed_frame(const SearchArgs &a, int slot) {
  return (slot < a.n_initial ? a.front_pool : a.pool) + (size_t)slot * a.m.frame_words;
}

// ---- ranks of a csolve_gpu_comm: incumbent and first-solution exchange over peer memory -------------------------
// An accepted improving leaf stores its key into every peer's CommBlock (lane r serves peer r: one 64-bit
// system-scope atomic each, 8 bytes over NVLink, no kernel drain, no host in the loop).
__device__ __forceinline__ void comm_push_best(const SearchArgs &a, int lane, int key) {
  if (lane < a.world && lane != a.rank) {
    if (a.m.objective == CSOLVE_OBJ_MIN) atomicMin_system(&a.peer_comm[lane]->rmin64, comm_key_min(a.epoch, key));
    else atomicMax_system(&a.peer_comm[lane]->rmax64, comm_key_max(a.epoch, key));
  }
  __syncwarp();
}
__device__ __forceinline__ void comm_push_stop(const SearchArgs &a, int lane) {
  if (lane < a.world && lane != a.rank) atomicMax_system(&a.peer_comm[lane]->stop_epoch, a.epoch);
  __syncwarp();
}
// poll of this rank's own block (local memory, the peers wrote it): fold a better incumbent into ctl->best, turn a
// peer's "found" into SIG_STOP. Lane 0; called where the control block is polled anyway.
__device__ __forceinline__ void comm_poll(const SearchArgs &a, int lane) {
  if (lane == 0) {
    if (a.m.objective == CSOLVE_OBJ_MIN) {
      const unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(&a.comm->rmin64);
      if (comm_key_is_epoch_min(k, a.epoch)) atomicMin(&a.ctl->best, comm_key_value(k));
    } else if (a.m.objective == CSOLVE_OBJ_MAX) {
      const unsigned long long k = *reinterpret_cast<volatile unsigned long long *>(&a.comm->rmax64);
      if (comm_key_is_epoch_max(k, a.epoch)) atomicMax(&a.ctl->best, comm_key_value(k));
    } else if (a.m.objective == CSOLVE_OBJ_ANY) {
      if (*reinterpret_cast<volatile int *>(&a.comm->stop_epoch) == a.epoch) atomicMax(&a.ctl->signal, SIG_STOP);
    }
  }
  __syncwarp();
}

// ---- the search kernel ----------------------------------------------------------------------------
// Frame header words (device_model.h): var, iter, last, lo | hi, level, best_seen, hash
// CSOLVE_BJ (kernels_bj.cu compiles this file a second time with it, up to the end of this kernel: CSOLVE_BJ_UNIT):
// the LEARN instance of the depth-first phase back-jumps after a conflict, see the block that follows learn_nogood in
// the node loop. A property of the compilation unit, not a template parameter: the instances of this unit keep their
// names and their code.
template <bool EXPAND, bool LEARN, bool LIN = false, bool SAMPLE = false>
__global__ void __launch_bounds__(THREADS_PER_BLOCK, CSOLVE_MIN_BLOCKS)
k_search(const SearchArgs a) {
  extern __shared__ __align__(16) int smem[];
  const DevModel &m = a.m;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * WARPS_PER_BLOCK + wib;
  const int4 *wrec; const int *wptr;
  __shared__ int s_blk_hungry;       // warps of this block waiting for a frame: their neighbours poll for donations faster
  if (threadIdx.x == 0) s_blk_hungry = 0;
  __syncthreads();
  LinTables ltab;
  stage_table(m, smem, wrec, wptr, ltab);
  if (gw >= a.n_warps) return;
  // per-warp regions are padded to 16 bytes so the int2 staging copies stay aligned
  const int wwords = (warp_smem_words(m, LEARN) + 3) & ~3;
  WarpSmem s = carve(m, smem + (m.table_smem_bytes >> 2) + wib * wwords);
  s.lt = ltab;

  const int V = m.n_vars, fw = m.frame_words;
  const bool optimise = m.obj_var >= 0;
  int *stack = a.stacks + (size_t)gw * (V + 1) * fw;
  SearchCtl *ctl = a.ctl;

  int level = a.wstate[gw].level, base = a.wstate[gw].base;
  Claim cl; cl.base = a.wstate[gw].claim_base; cl.mask = a.wstate[gw].claim_mask; cl.drained = false;
  unsigned long long nodes = 0, cuts = 0, sols = 0, refresh = 0, cuts_reported = 0;
  unsigned props = 0, visits = 0;
  const long long t0 = clock64();
  long long waited = 0, lastwork = -1;     // load-balance diagnostics (CSOLVE_DEBUG)
  unsigned claims = 0;

  // The header of the top frame lives in registers and its domains (the state BEFORE this level's
  // assignment) in s.p while the warp iterates over the level's values; HBM is touched only when
  // a frame is pushed, popped, refreshed, parked or fetched from the frontier.
  bool have = false;
  int var = 0, lo = 0, hi = 0, flevel = 0, fbest = 0;
  unsigned iter = 0, last = 0, fhash = 0, poll = 0;
  bool hungry = false;

  for (;;) {

    if (level < base) {
      // out of work: take a frame of the frontier / shared pool
      const int *src;
      int ring_slot = 0;
      if (EXPAND) {
        int it = 0;
        if (lane == 0) it = atomicAdd(&ctl->item_next, 1);
        it = __shfl_sync(FULL, it, 0);
        if (it >= ctl->item_count) break;
        src = a.items + (size_t)it * fw;
      } else {
        const long long w0 = clock64();
        lastwork = w0 - t0;
        const int slot = claim_frame<true>(a, lane, hungry, cl, &s_blk_hungry, t0);
        waited += clock64() - w0;
        if (slot < 0) break;
        claims++;
        src = claimed_frame(a, slot);
        ring_slot = slot;
      }
      const int L = EXPAND ? 0 : __ldcg(&src[FR_LEVEL]);
      int *dst = stack + (size_t)L * fw;
      for (int w = lane; w < fw; w += 32) __stcg(&dst[w], __ldcg(&src[w]));
      if (!EXPAND) {
        __syncwarp();
        if (lane == 0 && ring_slot >= a.n_initial) { __threadfence(); __stcg(&a.ready[ring_slot], 0); }   // slot may be reused
      }
      level = base = L;
      have = false;
      __syncwarp();
    }

    int *f = stack + (size_t)level * fw;
    if (!have) {
      // (re)load the top frame: header, assigned-variable mask, domains
      const int4 h0 = __ldcg(reinterpret_cast<const int4 *>(f));
      const int4 h1 = __ldcg(reinterpret_cast<const int4 *>(f) + 1);
      load_domains(m, f, s.p, lane);
      for (int w = lane; w < m.mask_words; w += 32) s.amask[w] = (unsigned)__ldcg(&f[FR_MASK + w]);
      var = h0.x; iter = (unsigned)h0.y; last = (unsigned)h0.z; lo = h0.w;
      hi = h1.x; flevel = h1.y; fbest = h1.z; fhash = (unsigned)h1.w;
      have = true;
      __syncwarp();
    }

    int best = 0;
    if (optimise) best = EXPAND ? a.frozen_best : *reinterpret_cast<volatile int *>(&ctl->best);

    if (EXPAND && last >= (unsigned)a.expand_branch_max) {
      // too many values to enumerate breadth-first: pass the frame through unchanged
      int slot = 0;
      if (lane == 0) { slot = atomicAdd(&ctl->out_count, 1); atomicAdd(&ctl->passed, 1); }
      slot = __shfl_sync(FULL, slot, 0);
      if (slot < a.out_cap) {
        int *g = a.items_out + (size_t)slot * fw;
        for (int w = lane; w < fw; w += 32) __stcg(&g[w], __ldcg(&f[w]));
      } else if (lane == 0) {
        atomicAdd(&ctl->out_dropped, 1);
      }
      level = base - 1;
      have = false;
      continue;
    }

    if (iter > last) {
      // values exhausted (src/csolve.c:439-442): backtrack; the parent frame is reloaded from HBM
      level--;
      have = false;
      continue;
    }

    if (!EXPAND && optimise && best != fbest) {
      // The incumbent improved since this frame's domains were computed. The reference
      // restarts from level 0 on every improvement (src/csolve.c:418-421); here the frame is
      // re-propagated against the tighter <obj> bound and keeps only its untried values.
      for (int v = lane; v < V; v += 32) reinterpret_cast<int2 *>(s.d)[v] = reinterpret_cast<const int2 *>(s.p)[v];
      for (int w = lane; w < m.mask_words; w += 32) { s.cur[w] = 0; s.nxt[w] = 0; }
      __syncwarp();
      bool ok = true;
      if (lane == 0) {
        Dom o; o.lo = s.d[2 * m.obj_var]; o.hi = s.d[2 * m.obj_var + 1];
        o = objective_tighten(m.objective, o, best);
        s.d[2 * m.obj_var] = o.lo; s.d[2 * m.obj_var + 1] = o.hi;
        s.cur[m.obj_var >> 5] |= 1u << (m.obj_var & 31);
        ok = o.lo <= o.hi;
      }
      ok = __shfl_sync(FULL, ok, 0);
      __syncwarp();
      if (LEARN) { for (int v = lane; v < V; v += 32) { s.rlo[v] = -1; s.rhi[v] = -1; } __syncwarp(); }
      if (ok) ok = warp_fixpoint<LEARN, LIN>(m, s, wrec, wptr, lane, props, visits, nullptr, &a.ng, nullptr);
      refresh++;
      // untried values of the old enumeration form the interval [lo + ceil(iter/2), hi - floor(iter/2)]
      const long long ua = (long long)lo + ((iter + 1) >> 1), ub = (long long)hi - (iter >> 1);
      long long na = ua, nb = ub;
      if (ok) {
        const long long dl = s.d[2 * var], dh = s.d[2 * var + 1];
        na = ua > dl ? ua : dl; nb = ub < dh ? ub : dh;
      }
      if (!ok || na > nb) {
        level--;
        have = false;
        continue;
      }
      // the frame keeps its (now tighter) parent domains, except that the branching variable
      // still ranges over the values this frame owns
      __syncwarp();
      if (lane == 0) { s.d[2 * var] = (int)na; s.d[2 * var + 1] = (int)nb; }
      __syncwarp();
      for (int v = lane; v < V; v += 32) reinterpret_cast<int2 *>(s.p)[v] = reinterpret_cast<const int2 *>(s.d)[v];
      store_domains(m, f, s.d, lane);
      lo = (int)na; hi = (int)nb; iter = 0; last = (unsigned)(nb - na); fbest = best;
      if (lane == 0) {
        __stcg(reinterpret_cast<int4 *>(f), make_int4(var, 0, (int)last, lo));
        __stcg(reinterpret_cast<int4 *>(f) + 1, make_int4(hi, flevel, fbest, (int)fhash));
      }
      __syncwarp();
      continue;
    }

    // ---- one search node: assign var := val, propagate (src/csolve.c:444-457) ----------------
    const int val = step_value(lo, hi, iter);
    iter++;
    for (int v = lane; v < V; v += 32) reinterpret_cast<int2 *>(s.d)[v] = reinterpret_cast<const int2 *>(s.p)[v];
    for (int w = lane; w < m.mask_words; w += 32) { s.cur[w] = 0; s.nxt[w] = 0; }
    __syncwarp();
    bool ok = true;
    if (lane == 0) {
      s.d[2 * var] = val; s.d[2 * var + 1] = val;
      s.cur[var >> 5] |= 1u << (var & 31);
      if (optimise) {
        Dom o; o.lo = s.d[2 * m.obj_var]; o.hi = s.d[2 * m.obj_var + 1];
        o = objective_tighten(m.objective, o, best);
        s.d[2 * m.obj_var] = o.lo; s.d[2 * m.obj_var + 1] = o.hi;
        s.cur[m.obj_var >> 5] |= 1u << (m.obj_var & 31);
        ok = o.lo <= o.hi;
      }
    }
    ok = __shfl_sync(FULL, ok, 0);
    if (LEARN) { for (int v = lane; v < V; v += 32) { s.rlo[v] = -1; s.rhi[v] = -1; } }
    __syncwarp();
    FailInfo fi; fi.rec = -1; fi.var = -1;
    if (ok) ok = warp_fixpoint<LEARN, LIN>(m, s, wrec, wptr, lane, props, visits, a.gprio, &a.ng, &fi);
    nodes++;
#ifdef CSOLVE_BJ
    int n_ng = 0;      // lane 0: literals of the nogood this node contributed (in s.nlits)
#endif
    if (LEARN && !ok) {
      // conflict_create (src/conflict.c:327-362): record why this node failed
      __syncwarp();
#ifdef CSOLVE_BJ
      if (lane == 0) n_ng = learn_nogood(m, s, wrec, wptr, a.ng, var, fi);
#else
      if (lane == 0) learn_nogood(m, s, wrec, wptr, a.ng, var, fi);
#endif
      __syncwarp();
    }
    // prio-- on success, prio++ on failure (src/csolve.c:459-462)
    if (a.gprio != nullptr && lane == 0) atomicAdd(&a.gprio[var], ok ? -1 : 1);
    bool s_hit = false;
    if (SAMPLE) {
      __syncwarp();
      s_hit = sample_hit(a, sample_dom_hash(s.p, V, lane), var, val, !ok);
      if (s_hit && !(ok && flevel + 1 == V)) {
        int *r = sample_begin(a, lane, ok ? 0 : SAMPLE_FAILED, var, val, best);
        if (r != nullptr) for (int w = lane; w < 2 * V; w += 32) { r[4 + w] = s.p[w]; r[4 + 2 * V + w] = s.d[w]; }
      }
      __syncwarp();      // every lane has copied its words of s.p / s.d before the push below overwrites s.p
    }

    if (!ok) {
      cuts++;
#ifdef CSOLVE_BJ
      if (LEARN && !EXPAND && m.objective != CSOLVE_OBJ_ALL) {
        // ---- back-jump (conflict_backtrack, src/csolve.c:350-364; conflict_update, src/conflict.c:311-324) -----------
        // The reference unwinds to 1 + the second-highest level of the nogood's variables, propagates the new nogood
        // there and goes on with a FRESH step at that level (the steps above it are deactivated, the next iteration
        // of solve() picks a variable again). With copy-on-branch frames: frame q of the stack holds the domains before
        // level q's assignment, so "the level variable y became a value at" is the first frame that has y as a value
        // (found by bisection over the frames of this warp's path, lane 0), T = the deepest such frame over the
        // nogood's literals other than the decision; the frames above T are dropped, frame T's domains are
        // propagated again with the nogood's variables on the worklist (the nogood now removes the decision's value
        // there, propagate_confl) and frame T is decided again from scratch. T is kept above the warp's base frame
        // (that frame may own only a part of its variable's values). Dropping and re-deciding frames loses nothing
        // -- the fresh frame T enumerates a whole domain under the same decisions, and what frame T had given away
        // (donation, k_rebalance) is searched by whoever took it -- so the result does not depend on T; what the
        // nogood buys is that the failure is not met again. Every jump stores a nogood, the pool is
        // finite, and with a full pool learn_nogood returns 0: the search is chronological from then on and ends.
        // ALL models never jump (a re-decided frame would count solutions twice: the reference's -c true over-counts).
        int T = level;
        if (lane == 0 && n_ng > 0 && level > base + 1) {
          int t = base + 1;
          const int doff = frame_dom_offset(m.mask_words);
          for (int k = 0; k < n_ng && t < level; k++) {
            const int y = s.nlits[k] >> 1;
            if (y == var) continue;
            int2 dy = __ldcg(reinterpret_cast<const int2 *>(stack + (size_t)t * fw + doff) + y);
            if (dy.x == dy.y) continue;                       // already a value in frame t
            int qa = t, qb = level;                           // not a value in frame qa, a value in frame qb (s.p)
            while (qb - qa > 1) {
              const int qm = (qa + qb) >> 1;
              dy = __ldcg(reinterpret_cast<const int2 *>(stack + (size_t)qm * fw + doff) + y);
              if (dy.x == dy.y) qb = qm; else qa = qm;
            }
            t = qb;
          }
          T = t;
        }
        T = __shfl_sync(FULL, T, 0);
        if (T < level) {
          level = T;
          int *ft = stack + (size_t)T * fw;
          const int4 g1 = __ldcg(reinterpret_cast<const int4 *>(ft) + 1);
          flevel = g1.y; fhash = (unsigned)g1.w;
          // The value of frame T that was being searched (frame T + 1 has T's variable at it). k_rebalance, k_export_frames
          // and the incumbent refresh narrow a frame's own variable to the values the frame still owns, which no longer
          // include that one -- and what is left of its sub-tree is in the frames being dropped.
          const int tvar = __ldcg(&ft[FR_VAR]);
          const int tval = __ldcg(&stack[(size_t)(T + 1) * fw + frame_dom_offset(m.mask_words) + 2 * tvar]);
          load_domains(m, ft, s.d, lane);
          for (int w = lane; w < m.mask_words; w += 32) { s.amask[w] = (unsigned)__ldcg(&ft[FR_MASK + w]); s.cur[w] = 0; s.nxt[w] = 0; }
          for (int v = lane; v < V; v += 32) { s.rlo[v] = -1; s.rhi[v] = -1; }
          __syncwarp();
          bool ok2 = true;
          if (lane == 0) {
            if (tval < s.d[2 * tvar]) s.d[2 * tvar] = tval;             // the re-decided frame covers that value again
            if (tval > s.d[2 * tvar + 1]) s.d[2 * tvar + 1] = tval;
            for (int k = 0; k < n_ng; k++) { const int y = s.nlits[k] >> 1; s.cur[y >> 5] |= 1u << (y & 31); }
            if (optimise) {
              Dom o; o.lo = s.d[2 * m.obj_var]; o.hi = s.d[2 * m.obj_var + 1];
              o = objective_tighten(m.objective, o, best);
              s.d[2 * m.obj_var] = o.lo; s.d[2 * m.obj_var + 1] = o.hi;
              s.cur[m.obj_var >> 5] |= 1u << (m.obj_var & 31);
              ok2 = o.lo <= o.hi;
            }
            atomicAdd(&a.ng.counters[5], 1);
          }
          ok2 = __shfl_sync(FULL, ok2, 0);
          __syncwarp();
          if (ok2) ok2 = warp_fixpoint<LEARN, LIN>(m, s, wrec, wptr, lane, props, visits, nullptr, &a.ng, nullptr);
          if (!ok2) {
            // nothing is left below frame T - 1's current value: it goes on with its next one (T - 1 >= base)
            level = T - 1;
            have = false;
            continue;
          }
          const int nv = warp_select_var(m, s, lane, a.order, flevel, -1, a.gprio);
          fbest = optimise ? best : g1.z;
          write_child_frame(m, s, ft, lane, nv, flevel, fbest, fhash, -1);
          for (int v = lane; v < V; v += 32) reinterpret_cast<int2 *>(s.p)[v] = reinterpret_cast<const int2 *>(s.d)[v];
          __syncwarp();
          lo = s.p[2 * nv]; hi = s.p[2 * nv + 1];
          var = nv; iter = 0; last = (unsigned)hi - (unsigned)lo;
          have = true;
        }
      }
#endif
    } else if (flevel + 1 == V) {
      // all variables assigned: leaf (src/csolve.c:416-424, 222-244)
      const bool leaf_ok = warp_all_true(m, s, lane);
      if (SAMPLE && s_hit) {
        int *r = sample_begin(a, lane, leaf_ok ? SAMPLE_LEAF : 0, var, val, best);
        if (r != nullptr) for (int w = lane; w < 2 * V; w += 32) { r[4 + w] = s.p[w]; r[4 + 2 * V + w] = s.d[w]; }
      }
      if (leaf_ok) {
        bool accepted = true;
        int key = 0;
        if (m.objective == CSOLVE_OBJ_MIN) {
          key = s.d[2 * m.obj_var];
          int old = 0;
          if (lane == 0) old = atomicMin(&ctl->best, key);
          old = __shfl_sync(FULL, old, 0);
          accepted = key < old;
          if (accepted && a.n_peers > 0) comm_push_best(a, lane, key);
        } else if (m.objective == CSOLVE_OBJ_MAX) {
          key = s.d[2 * m.obj_var + 1];
          int old = 0;
          if (lane == 0) old = atomicMax(&ctl->best, key);
          old = __shfl_sync(FULL, old, 0);
          accepted = key > old;
          if (accepted && a.n_peers > 0) comm_push_best(a, lane, key);
        } else if (m.objective == CSOLVE_OBJ_ANY) {
          int old = 0;
          if (lane == 0) old = atomicMax(&ctl->signal, SIG_STOP);
          old = __shfl_sync(FULL, old, 0);
          accepted = old != SIG_STOP;     // first finder wins (found_any(), src/csolve.c:207-209)
          if (accepted && a.n_peers > 0) comm_push_stop(a, lane);
        }
        if (accepted) {
          sols++;
          if (a.inst_solutions != nullptr) {       // batched roots: header word 6 carries the root id
            key = fbest;
            if (lane == 0) atomicAdd(&a.inst_solutions[fbest], 1u);
          }
          store_solution(a, s, lane, key);
        }
      }
    } else {
      const int nv = warp_select_var(m, s, lane, a.order, flevel + 1, var, a.gprio);
      const unsigned chash = mix_hash(fhash, (unsigned)var, (unsigned)val);
      if (EXPAND) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(&ctl->out_count, 1);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot < a.out_cap) {
          write_child_frame(m, s, a.items_out + (size_t)slot * fw, lane, nv, flevel + 1, optimise ? best : fbest, chash, var);
        } else if (lane == 0) {
          atomicAdd(&ctl->out_dropped, 1);
        }
      } else {
        // push: the current frame's iteration state goes to HBM (it is reloaded on backtrack), the
        // child frame is written for rebalancing/parking and becomes the register/shared-resident top
        if (lane == 0) __stcg(&f[FR_ITER], (int)iter);
        write_child_frame(m, s, stack + (size_t)(level + 1) * fw, lane, nv, flevel + 1, optimise ? best : fbest, chash, var);
        if (lane == 0) s.amask[var >> 5] |= 1u << (var & 31);
        for (int v = lane; v < V; v += 32) reinterpret_cast<int2 *>(s.p)[v] = reinterpret_cast<const int2 *>(s.d)[v];
        __syncwarp();
        lo = s.p[2 * nv]; hi = s.p[2 * nv + 1];
        var = nv; iter = 0; last = (unsigned)hi - (unsigned)lo;
        flevel = flevel + 1; fbest = optimise ? best : fbest; fhash = chash;
        level++;
      }
      __syncwarp();
    }

    // ---- park? ----------------------------------------------------------------------------------
    // every few nodes: has a slice end / stop been requested, is the time slice over, is somebody waiting for work?
    // (general-kernel nodes are long, so the control block is polled often; an L2 round trip per node would
    // dominate the short nodes of the lane-owns-variable kernel, which polls every POLL_NODES nodes)
    if (!EXPAND && ((++poll & 3u) == 0 || *reinterpret_cast<volatile int *>(&s_blk_hungry) > 0)) {
      if (a.n_peers > 0) comm_poll(a, lane);
      if (a.fail_limit > 0 && restart_due(a, lane, cuts, cuts_reported)) break;
      if (a.sink_headroom > 0 && *reinterpret_cast<volatile int *>(&ctl->n_stored) > a.max_solutions - a.sink_headroom) {
        if (lane == 0) atomicMax(&ctl->signal, SIG_SLICE_END);      // the solution buffer is nearly full: let the host drain it
        break;
      }
      if (*reinterpret_cast<volatile int *>(&ctl->signal) != SIG_RUN) break;
      if (clock64() - t0 > a.slice_cycles) {
        if (lane == 0) atomicMax(&ctl->signal, SIG_SLICE_END);
        break;
      }
      int tgt = -1;
      if (level >= base && (tgt = donation_target(a, lane)) >= 0) {
        // shallowest frame with at least two untried values; the top frame's header is in registers
        int L = -1;
        unsigned d_iter = 0; int d_lo = 0, d_hi = 0;
        if (lane == 0) {
          for (int q = base; q <= level; ++q) {
            unsigned it2, la2; int lo2, hi2;
            if (q == level) { it2 = iter; la2 = last; lo2 = lo; hi2 = hi; }
            else {
              const int4 g0 = __ldcg(reinterpret_cast<const int4 *>(stack + (size_t)q * fw));
              it2 = (unsigned)g0.y; la2 = (unsigned)g0.z; lo2 = g0.w; hi2 = __ldcg(&stack[(size_t)q * fw + FR_HI]);
            }
            // a frame below the top is given away whole (the warp still has the deeper levels); the top frame is halved
            if (it2 <= la2 && (q < level || la2 - it2 >= 1)) { L = q; d_iter = it2; d_lo = lo2; d_hi = hi2; break; }
          }
        }
        L = __shfl_sync(FULL, L, 0);
        if (L >= 0) {
          d_iter = __shfl_sync(FULL, d_iter, 0); d_lo = __shfl_sync(FULL, d_lo, 0); d_hi = __shfl_sync(FULL, d_hi, 0);
          const long long ua = (long long)d_lo + ((d_iter + 1) >> 1), ub = (long long)d_hi - (d_iter >> 1);
          const long long mid = L < level ? ua - 1 : ua + (ub - ua) / 2;
          int *own = stack + (size_t)L * fw;
          for (;;) {
            const int slot = reserve_slot(a, lane, tgt);
            int *g = ring_frame(a, tgt, slot);
            for (int w = lane; w < fw; w += 32) __stcg(&g[w], __ldcg(&own[w]));
            __syncwarp();
            if (lane == 0) {
              // the donated frame owns [mid + 1, ub]; this warp keeps [ua, mid]; both restart their value iteration
              __stcg(&g[FR_ITER], 0); __stcg(&g[FR_LO], (int)(mid + 1)); __stcg(&g[FR_HI], (int)ub);
              __stcg(&g[FR_LAST], (int)(unsigned)(ub - mid - 1));
            }
            if (publish_slot(a, lane, tgt, slot)) break;
          }
          if (lane == 0) {
            if (L < level) { __stcg(&own[FR_ITER], 1); __stcg(&own[FR_LAST], 0); }     // exhausted: iter > last
            else {
              __stcg(&own[FR_ITER], 0); __stcg(&own[FR_LO], (int)ua); __stcg(&own[FR_HI], (int)mid);
              __stcg(&own[FR_LAST], (int)(unsigned)(mid - ua));
            }
          }
          if (L == level) { iter = 0; lo = (int)ua; hi = (int)mid; last = (unsigned)(mid - ua); }
          __syncwarp();
        }
        donation_done(a, lane, tgt);
      }
    }
  }

  if (lane == 0) {
    if (have && level >= base) __stcg(&stack[(size_t)level * fw + FR_ITER], (int)iter);   // park the top frame
    a.wstate[gw].level = level;
    a.wstate[gw].base = base;
    a.wstate[gw].claim_base = cl.base; a.wstate[gw].claim_mask = cl.mask;
  }
  // flush counters (the slot belongs to this warp; values accumulate across slices)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    props += __shfl_xor_sync(FULL, props, o);
    visits += __shfl_xor_sync(FULL, visits, o);
  }
  if (lane == 0) {
    unsigned long long *c = a.wcount + (size_t)gw * CNT_WIDTH;
    c[CNT_NODES] += nodes; c[CNT_CUTS] += cuts; c[CNT_PROPS] += props;
    c[CNT_VISITS] += visits; c[CNT_SOLUTIONS] += sols; c[CNT_REFRESH] += refresh;
    c[CNT_WAIT] += (unsigned long long)waited; c[CNT_CLAIMS] += claims;
    c[CNT_LASTWORK] = (unsigned long long)(lastwork >= 0 ? lastwork : clock64() - t0);
  }
}

#ifdef CSOLVE_BJ_UNIT
// kernels_bj.cu compiles this file up to here, in a namespace of its own, for the one back-jumping instance of
// k_search: the instances of this unit keep the code they were measured and parity-tested with (an instantiation more
// in the same unit moves the inliner's choices in the others).
}  // namespace csolve_dev (renamed by kernels_bj.cu)
extern "C" const void *csolve_bj_search_kernel(void) {
  return (const void *)csolve_dev::k_search<false, true, false, false>;
}
#else

// =====================================================================================================
// "Lane owns variable" search kernel for pure NOT(EQ) networks with at most 32 variables (N-queens).
// Lane j keeps the bounds of variable j in registers: a node's domain vector never touches shared
// memory, a dequeued variable is broadcast with two shuffles, every lane contracts its own clauses
// with it (lov_lane_step), bounds requested of the dequeued variable are combined with ballots.
// No atomics, no divergence; the worklist is a 32-bit register mask. Same frames, same frontier,
// same rebalancing as k_search.
struct LovTables { const unsigned long long *pair; const int *cptr; const int *cval; const uint32_t *fconst; };

__device__ __forceinline__ LovTables stage_lov(const DevModel &m, int *smem) {
  unsigned long long *pair = reinterpret_cast<unsigned long long *>(smem);
  int *cptr = smem + m.n_vars * 64;
  int *cval = cptr + m.n_vars + 1;
  uint32_t *fconst = reinterpret_cast<uint32_t *>(cval + m.n_lov_cval);
  for (int i = threadIdx.x; i < m.n_vars; i += blockDim.x) fconst[i] = __ldg(&m.lov_fconst[i]);
  for (int i = threadIdx.x; i < m.n_vars * 32; i += blockDim.x) pair[i] = __ldg(&m.lov_pair[i]);
  for (int i = threadIdx.x; i <= m.n_vars; i += blockDim.x) cptr[i] = __ldg(&m.lov_cptr[i]);
  for (int i = threadIdx.x; i < m.n_lov_cval; i += blockDim.x) cval[i] = __ldg(&m.lov_cval[i]);
  __syncthreads();
  LovTables t; t.pair = pair; t.cptr = cptr; t.cval = cval; t.fconst = fconst;
  return t;
}

// propagate to fixpoint; lo/hi: this lane's variable. returns false when the node failed.
// The dequeued variable's bounds are warp-uniform, so "is it a value?" is a uniform branch:
//   value     -> every lane trims its own bounds (registers only), one ballot collects who changed;
//   not value -> lanes whose variable is a value ask for the dequeued variable's bounds to move,
//                two ballots combine the requests and the owning lane applies them.
__device__ __forceinline__ bool lov_fixpoint(const LovTables &t, int V, bool has_consts, int lane, int &lo, int &hi,
                                             unsigned changed, unsigned &props, unsigned &visits) {
  const bool act = lane < V;
  while (changed) {
    const int i = __ffs(changed) - 1;
    changed &= changed - 1;
    const int Xlo = __shfl_sync(FULL, lo, i), Xhi = __shfl_sync(FULL, hi, i);
    if (Xlo > Xhi) return false;                 // bounds requested by different lanes crossed
    const unsigned long long mask = act ? t.pair[i * 32 + lane] : 0ull;
    const LovStep r = lov_lane_step(mask, Xlo, Xhi, lo, hi);
    unsigned chm = 0;
    if (Xlo == Xhi) {
      const bool mych = (r.nlo != lo) | (r.nhi != hi);
      lo = r.nlo; hi = r.nhi;
      chm = __ballot_sync(FULL, mych);
    }
    bool plo = r.plo, phi = r.phi;
    if (has_consts) {
      const int cb = t.cptr[i], ce = t.cptr[i + 1];
      if (cb + lane < ce) lov_const_step(t.cval[cb + lane], Xlo, Xhi, plo, phi);
    }
    if (has_consts || Xlo != Xhi) {
      const unsigned bl = __ballot_sync(FULL, plo), bh = __ballot_sync(FULL, phi);
      if (lane == i) { if (bl) lo = Xlo + 1; if (bh) hi = Xhi - 1; }
      if (bl | bh) chm |= 1u << i;
    }
    changed |= chm;
    props += (chm >> lane) & 1u;                       // per lane; summed over the warp when the counters are flushed
    visits += (unsigned)V;
  }
  return true;
}

// Forbidden-value-set form of the same fixpoint (contract.cuh: lov_forbid / lov_trim). `pend` holds the
// variables that became a value and whose forbidden values have not been distributed yet.
__device__ __forceinline__ bool lov_fixpoint_bits(const LovTables &t, int V, int vbase, int lane, int &lo, int &hi,
                                                  uint32_t &F, unsigned pend, unsigned &props, unsigned &visits) {
  while (pend) {
    // the forbidden values of ALL variables that just became a value are distributed, then every lane trims once
    // (the greatest fixpoint does not depend on the order)
    do {
      const int i = __ffs(pend) - 1;
      pend &= pend - 1;
      const int w = __shfl_sync(FULL, lo, i);
      F |= lov_forbid(t.pair[i * 32 + lane], w, vbase);    // lanes >= V read the zero entries of the table: no branch
      visits++;                                            // counted in variables here, scaled by V when flushed
    } while (pend);
    const bool was = lo == hi;
    const int olo = lo, ohi = hi;
    const bool alive = lov_trim(F, vbase, lo, hi);
    if (__any_sync(FULL, !alive)) return false;
    pend = __ballot_sync(FULL, !was && lo == hi);
    props += (lo != olo || hi != ohi) ? 1u : 0u;      // per lane; summed over the warp when the counters are flushed
  }
  return true;
}
// the same fixpoint entered from a search node: the first round has exactly one variable, the decision var := val, and
// both are warp-uniform -- no bit scan, no shuffle for it
__device__ __forceinline__ bool lov_fixpoint_bits_node(const LovTables &t, int vbase, int lane, int &lo, int &hi,
                                                       uint32_t &F, int var, int val, unsigned &props, unsigned &visits) {
  F |= lov_forbid(t.pair[var * 32 + lane], val, vbase);
  visits++;
  for (;;) {
    const bool was = lo == hi;
    const int olo = lo, ohi = hi;
    const bool alive = lov_trim(F, vbase, lo, hi);
    if (__any_sync(FULL, !alive)) return false;
    unsigned pend = __ballot_sync(FULL, !was && lo == hi);
    props += (lo != olo || hi != ohi) ? 1u : 0u;
    if (pend == 0u) return true;
    do {
      const int i = __ffs(pend) - 1;
      pend &= pend - 1;
      const int w = __shfl_sync(FULL, lo, i);
      F |= lov_forbid(t.pair[i * 32 + lane], w, vbase);
      visits++;
    } while (pend);
  }
}
// forbidden-value set of this lane's variable for a frame whose domains are in sdom (lo,hi pairs)
__device__ __forceinline__ uint32_t lov_rebuild_F(const LovTables &t, int V, int vbase, int lane, const int *sdom) {
  uint32_t F = lane < V ? t.fconst[lane] : 0u;
  for (int i = 0; i < V; i++) {
    const int2 di = reinterpret_cast<const int2 *>(sdom)[i];
    if (di.x == di.y && lane < V) F |= lov_forbid(t.pair[i * 32 + lane], di.x, vbase);
  }
  return F;
}

// Shared-memory frame of the LOV kernel (private to it; converted at the HBM boundary by frame_in / frame_out):
//   word 0  cursor   next value of the branching variable to try        word 4  amask  variables that have a level
//   word 1  rem      values left: the frame owns [cursor, cursor+rem)   word 5  hash   path hash (rank partition)
//   word 2  var      branching variable                                 word 6,7 unused
//   word 3  level
// followed by the 2 * V domain words and the V forbidden-value sets. Values are tried in ASCENDING order here
// (the reference alternates lo, hi, lo+1, ... -- src/csolve.c:331-338 -- which matters only for which solution ANY
// meets first; counts are order-free), so the untried values of a level always form one interval and whole runs
// of values that are already forbidden are counted as failed nodes in one step (one bit scan instead of one loop
// iteration each). The warp's whole DFS stack lives in shared memory during a slice; it is loaded from / parked
// to the HBM frames (device_model.h layout, iter = 0 over the remaining interval) at the slice boundaries, where
// k_rebalance and the host see it.
__device__ __forceinline__ int lov_sframe_words(int V) { return (8 + 3 * V + 3) & ~3; }   // header, domains, value sets; 16-byte aligned

template <bool EXPAND, bool BITS, bool SAMPLE = false>
__global__ void __launch_bounds__(THREADS_PER_BLOCK, CSOLVE_LOV_MIN_BLOCKS)
k_search_lov(const SearchArgs a) {
  extern __shared__ __align__(16) int smem[];
  const DevModel &m = a.m;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * WARPS_PER_BLOCK + wib;
  __shared__ int s_blk_hungry;       // warps of this block waiting for a frame: their neighbours poll for donations faster
  if (threadIdx.x == 0) s_blk_hungry = 0;
  const LovTables T = stage_lov(m, smem);
  if (gw >= a.n_warps) return;

  const int V = m.n_vars, fw = m.frame_words;
  const int dofs = frame_dom_offset(1);
  const int sfw = lov_sframe_words(V);
  int *sst = smem + (m.lov_smem_bytes >> 2) + wib * (EXPAND ? sfw : V * sfw);   // this warp's stack
  int *stack = a.stacks + (size_t)gw * (V + 1) * fw;
  SearchCtl *ctl = a.ctl;
  const bool act = lane < V;
  const bool has_consts = m.n_lov_cval > 0;
  // BITS: inside the kernel a value is its bit index in the value sets (value - lov_vbase, 0..31); zb is added back
  // wherever a value leaves the kernel (HBM frames, stored solutions). Saves the +- vbase of every trim and shift.
  const int zb = BITS ? m.lov_vbase : 0;
  const int vbase = BITS ? 0 : m.lov_vbase;

  int level = a.wstate[gw].level, base = a.wstate[gw].base;
  Claim cl; cl.base = a.wstate[gw].claim_base; cl.mask = a.wstate[gw].claim_mask; cl.drained = false;
  unsigned long long nodes = 0, cuts = 0, sols = 0;
  unsigned n32 = 0, c32 = 0;   // nodes / cuts since the last flush into the 64-bit counters
  unsigned props = 0, visits = 0;
  const long long t0 = clock64();
  long long waited = 0, lastwork = -1;
  unsigned claims = 0, dbg_polls = 0, dbg_wanted = 0, dbg_donated = 0;

  // HBM frame -> shared frame. The untried values of an alternating enumeration at iteration `it` are the
  // interval [lo + ceil(it / 2), hi - floor(it / 2)]: la - it + 1 values.
  auto frame_in = [&](const int *g, int *sf) {
    if (lane == 0) {
      const int4 h0 = __ldcg(reinterpret_cast<const int4 *>(g));
      const int4 h1 = __ldcg(reinterpret_cast<const int4 *>(g) + 1);
      const int mk = __ldcg(&g[FR_MASK]);
      const unsigned it = (unsigned)h0.y, la = (unsigned)h0.z;
      const bool left = it <= la;
      CHK(h0.x >= 0 && h0.x < V, "lov frame_in var", h0.x);
      CHK(h1.y >= 0 && h1.y < V, "lov frame_in level", h1.y);
      CHK(!left || la - it < 32u, "lov frame_in rem", la - it);
      reinterpret_cast<int4 *>(sf)[0] = make_int4((left ? (int)((unsigned)h0.w + ((it + 1) >> 1)) : h0.w) - zb, left ? (int)(la - it + 1u) : 0, h0.x, h1.y);
      reinterpret_cast<int2 *>(sf)[2] = make_int2(mk, h1.w);
    }
    if (act) {
      int2 d = __ldcg(reinterpret_cast<const int2 *>(g + dofs) + lane);
      d.x -= zb; d.y -= zb;
      reinterpret_cast<int2 *>(sf + 8)[lane] = d;
    }
    if (BITS) {
      __syncwarp();
      const uint32_t F = lov_rebuild_F(T, V, vbase, lane, sf + 8);
      if (act) sf[8 + 2 * V + lane] = (int)F;
    }
  };
  // shared frame -> HBM frame (best_seen is unused by pure NOT(EQ) networks); an exhausted frame is iter 1 > last 0
  auto frame_out = [&](const int *sf, int *g) {
    if (lane == 0) {
      const int4 s0 = reinterpret_cast<const int4 *>(sf)[0];
      const int2 s1 = reinterpret_cast<const int2 *>(sf)[2];
      const unsigned rm = (unsigned)s0.y;
      __stcg(reinterpret_cast<int4 *>(g), make_int4(s0.z, rm ? 0 : 1, rm ? (int)(rm - 1u) : 0, s0.x + zb));
      __stcg(reinterpret_cast<int4 *>(g) + 1, make_int4((rm ? (int)((unsigned)s0.x + rm - 1u) : s0.x) + zb, s0.w, 0, s1.y));
      __stcg(&g[FR_MASK], s1.x);
    }
    if (act) {
      int2 d = reinterpret_cast<const int2 *>(sf + 8)[lane];
      d.x += zb; d.y += zb;
      __stcg(reinterpret_cast<int2 *>(g + dofs) + lane, d);
    }
  };

  if (!EXPAND && level >= base) {
    for (int L = base; L <= level; ++L) frame_in(stack + (size_t)L * fw, sst + L * sfw);
    __syncwarp();
  }

  // parity instrumentation (SAMPLE instances only): identity hash of a parent state, one record
  auto s_hash = [&](int l, int h) {
    return __reduce_xor_sync(FULL, act ? mix_hash(0x9E3779B9u + (unsigned)lane, (unsigned)(l + zb), (unsigned)(h + zb)) : 0u);
  };
  auto s_record = [&](int flags, int svar, int sval, int l0, int h0, int l1, int h1) {
    int *r = sample_begin(a, lane, flags, svar, sval + zb, 0);
    if (r != nullptr && act) {
      reinterpret_cast<int2 *>(r + 4)[lane] = make_int2(l0 + zb, h0 + zb);
      reinterpret_cast<int2 *>(r + 4 + 2 * V)[lane] = make_int2(l1 + zb, h1 + zb);
    }
  };
  // values [b0, b1) (bit indices) of the top frame's variable that are already forbidden: failed nodes, counted in bulk
  auto s_skipped = [&](int b0, int b1, int svar, int l0, int h0) {
    const unsigned hp = s_hash(l0, h0);
    for (int b = b0; b < b1; b++)
      if (sample_hit(a, hp, svar, vbase + b + zb, true)) s_record(SAMPLE_FAILED | SAMPLE_COUNTED, svar, vbase + b, l0, h0, l0, h0);
  };

  bool have = false;
  int *sf = sst + level * sfw;        // the top frame; moved with the stack (recomputing it cost 8 instructions per node)
  int var = 0, cur = 0, flevel = 0;   // top frame: branching variable, cursor, level
  unsigned rem = 0;                   // ... values left: [cur, cur + rem)
  unsigned fhash = 0, amask = 0;
  int plo = 0, phi = 0;        // this lane's variable in the top frame (state before the assignment)
  uint32_t pF = 0;             // ... and its forbidden-value set (BITS)
  uint32_t fvarF = 0;          // forbidden-value set of the top frame's branching variable (warp-uniform)
  uint32_t avail = 0;          // values of [cur, cur + rem) that are not in fvarF (BITS)
  unsigned poll = 0;
  // the last level is not searched: with every other variable a value, each value its forbidden set leaves is a solution
#ifdef CSOLVE_NO_COUNT_LAST
  const bool count_last = false;
#else
  const bool count_last = BITS && !EXPAND && m.objective == CSOLVE_OBJ_ALL && a.max_solutions == 0 && V >= 2;
#endif

  bool hungry = false;
  for (;;) {
    if (level < base) {
      // out of work: take a frame of the frontier / shared pool
      const int *src;
      int ring_slot = 0;
      if (EXPAND) {
        int it = 0;
        if (lane == 0) it = atomicAdd(&ctl->item_next, 1);
        it = __shfl_sync(FULL, it, 0);
        if (it >= ctl->item_count) break;
        src = a.items + (size_t)it * fw;
      } else {
        const long long w0 = clock64();
        lastwork = w0 - t0;
        const int slot = claim_frame<false>(a, lane, hungry, cl, &s_blk_hungry, t0);
        waited += clock64() - w0;
        if (slot < 0) break;
        claims++;
        src = claimed_frame(a, slot);
        ring_slot = slot;
      }
      const int L = EXPAND ? 0 : __ldcg(&src[FR_LEVEL]);
      CHK(L >= 0 && L < V, "lov claimed level", L);
      frame_in(src, sst + L * sfw);
      if (!EXPAND) {
        __syncwarp();
        if (lane == 0 && ring_slot >= a.n_initial) { __threadfence(); __stcg(&a.ready[ring_slot], 0); }   // slot may be reused
      }
      level = base = L;
      sf = sst + L * sfw;
      have = false;
      __syncwarp();
    }

    if (!have) {
      CHK(sf >= sst && sf + sfw <= sst + (EXPAND ? 1 : V) * sfw, "lov top frame", (sf - sst) / sfw);
      const int4 h0 = reinterpret_cast<const int4 *>(sf)[0];
      const int2 h1 = reinterpret_cast<const int2 *>(sf)[2];
      CHK(h0.z >= 0 && h0.z < V && h0.w >= 0 && h0.w < V, "lov top frame var/level", h0.z * 1000 + h0.w);
      int2 dj = make_int2(vbase, vbase);       // idle lanes hold a harmless one-value domain
      if (act) dj = reinterpret_cast<const int2 *>(sf + 8)[lane];
      plo = dj.x; phi = dj.y;
      if (BITS) pF = act ? (uint32_t)sf[8 + 2 * V + lane] : 0u;
      cur = h0.x; rem = (unsigned)h0.y; var = h0.z; flevel = h0.w;
      amask = (unsigned)h1.x; fhash = (unsigned)h1.y;
      if (BITS) {
        fvarF = __shfl_sync(FULL, pF, var);
        avail = rem ? (~fvarF & ((rem >= 32u ? 0xffffffffu : ((1u << rem) - 1u)) << (cur - vbase))) : 0u;
      }
      have = true;
      if (EXPAND && rem > (unsigned)a.expand_branch_max) {
        // too many values to enumerate breadth-first: pass the frame through unchanged
        int slot = 0;
        if (lane == 0) { slot = atomicAdd(&ctl->out_count, 1); atomicAdd(&ctl->passed, 1); }
        slot = __shfl_sync(FULL, slot, 0);
        if (slot < a.out_cap) frame_out(sf, a.items_out + (size_t)slot * fw);
        else if (lane == 0) atomicAdd(&ctl->out_dropped, 1);
        level = base - 1;
        have = false;
        continue;
      }
    }

    // ---- next value of the level -------------------------------------------------------------------
    int val = cur;
    if (BITS) {
      // values of [cur, cur + rem) the fixed variables (and constants) do not forbid yet; a forbidden value's node
      // fails at the first trim of its own variable (lov_trim on [val, val]): counted, not executed
      // (`avail` is kept in a register while the frame is the top one)
      if (avail == 0u) {
        if (SAMPLE) s_skipped(cur - vbase, cur - vbase + (int)rem, var, plo, phi);
        n32 += rem; c32 += rem;
        level--; sf -= sfw;
        have = false;
        continue;
      }
      const int b = __ffs((int)avail) - 1;
      avail &= avail - 1u;
      const unsigned skipped = (unsigned)(b - (cur - vbase));
      if (SAMPLE) s_skipped(cur - vbase, b, var, plo, phi);
      n32 += skipped; c32 += skipped;
      val = vbase + b;
      rem -= skipped + 1u;
    } else {
      if (rem == 0u) {
        level--; sf -= sfw;
        have = false;
        continue;
      }
      rem--;
    }
    cur = val + 1;

    // ---- one search node -------------------------------------------------------------------------
    int lo = plo, hi = phi;
    if (lane == var) { lo = val; hi = val; }
    uint32_t F = pF;
    const bool ok = BITS ? lov_fixpoint_bits_node(T, vbase, lane, lo, hi, F, var, val, props, visits)
                         : lov_fixpoint(T, V, has_consts, lane, lo, hi, 1u << var, props, visits);
    n32++;
    bool s_hit = false;
    if (SAMPLE) {
      s_hit = sample_hit(a, s_hash(plo, phi), var, val + zb, !ok);
      if (s_hit && !(ok && flevel + 1 == V)) s_record(ok ? 0 : SAMPLE_FAILED, var, val, plo, phi, lo, hi);
    }

    if (!ok) {
      c32++;
    } else if (flevel + 1 == V) {
      // leaf: is_true(eval(root)) -- every clause x_i + c != x_j / x_i != c holds on the assignment
      bool good = lo == hi;
      if (BITS) good = good && !((F >> (lo - vbase)) & 1u);     // F holds every value the other variables forbid
      for (int i = 0; !BITS && i < V; i++) {
        const int Xi = __shfl_sync(FULL, lo, i);
        const unsigned long long mk = act ? T.pair[i * 32 + lane] : 0ull;
        if (lov_has(mk, lo - Xi)) good = false;      // x_i + c == x_j for a clause of the pair
        const int cb = T.cptr[i], ce = T.cptr[i + 1];
        if (cb + lane < ce && T.cval[cb + lane] == Xi) good = false;
      }
      const bool leaf_ok = __all_sync(FULL, good || !act);
      if (SAMPLE && s_hit) s_record(leaf_ok ? SAMPLE_LEAF : 0, var, val, plo, phi, lo, hi);
      if (leaf_ok) {
        bool accepted = true;
        if (m.objective == CSOLVE_OBJ_ANY) {
          int old = 0;
          if (lane == 0) old = atomicMax(&ctl->signal, SIG_STOP);
          old = __shfl_sync(FULL, old, 0);
          accepted = old != SIG_STOP;
          if (accepted && a.n_peers > 0) comm_push_stop(a, lane);
        }
        if (accepted) {
          sols++;
          int slot = 0;
          if (lane == 0) slot = atomicAdd(&ctl->n_stored, 1);
          slot = __shfl_sync(FULL, slot, 0);
          if (slot < a.max_solutions) {
            int *dst = a.solbuf + (size_t)slot * (V + 1);
            if (act) dst[lane] = lo + zb;
            if (lane == 0) dst[V] = 0;
          }
          // Reconverge here: without it the lanes that skipped the stores ran ahead into the shuffles of the poll /
          // donation code and paired with the wrong ones (measured on B200: garbage split points, lost nodes,
          // illegal addresses -- only with max_solutions > 0, i.e. when this block stores anything).
          __syncwarp();
        }
      }
    } else {
      // branching variable of the next level
      int nv;
      if (a.order == CSOLVE_ORDER_NONE) {
        nv = __ldg(&m.order[flevel + 1]);
      } else {
        unsigned long long bestk = ~0ull;
        int bestv = 0x7fffffff;
        if (act && lane != var && !(amask & (1u << lane))) {
          unsigned primary;
          switch (a.order) {
          case CSOLVE_ORDER_SMALLEST_DOMAIN: primary = (unsigned)hi - (unsigned)lo; break;
          case CSOLVE_ORDER_LARGEST_DOMAIN:  primary = ~((unsigned)hi - (unsigned)lo); break;
          case CSOLVE_ORDER_SMALLEST_VALUE:  primary = (unsigned)lo ^ 0x80000000u; break;
          default:                           primary = ~((unsigned)hi ^ 0x80000000u); break;
          }
          const unsigned secondary = ~((unsigned)__ldg(&m.prio[lane]) ^ 0x80000000u);
          bestk = ((unsigned long long)primary << 32) | secondary;
          bestv = lane;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long ok2 = __shfl_xor_sync(FULL, bestk, o);
          const int ov = __shfl_xor_sync(FULL, bestv, o);
          if (ok2 < bestk || (ok2 == bestk && ov < bestv)) { bestk = ok2; bestv = ov; }
        }
        nv = bestv;
      }
      const int nlo = __shfl_sync(FULL, lo, nv), nhi = __shfl_sync(FULL, hi, nv);
      const unsigned nrem = (unsigned)nhi - (unsigned)nlo + 1u;
      const unsigned nmask = amask | (1u << var);
      if (EXPAND) {
        // the path hash decides which rank searches the frame (only frames of the expanded frontier are partitioned)
        const unsigned chash = mix_hash(fhash, (unsigned)var, (unsigned)val);
        int slot = 0;
        if (lane == 0) slot = atomicAdd(&ctl->out_count, 1);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot < a.out_cap) {
          int *g = a.items_out + (size_t)slot * fw;
          if (lane == 0) {
            __stcg(reinterpret_cast<int4 *>(g), make_int4(nv, 0, (int)(nrem - 1u), nlo + zb));
            __stcg(reinterpret_cast<int4 *>(g) + 1, make_int4(nhi + zb, flevel + 1, 0, (int)chash));
            __stcg(&g[FR_MASK], (int)nmask);
          }
          if (act) __stcg(reinterpret_cast<int2 *>(g + dofs) + lane, make_int2(lo + zb, hi + zb));
        } else if (lane == 0) {
          atomicAdd(&ctl->out_dropped, 1);
        }
      } else if (count_last && flevel + 2 == V) {
        // nv is the last variable: F_nv holds everything the V - 1 values forbid (the clause tables are symmetric),
        // so each remaining value is an accepted leaf and each forbidden one a failed node -- counted, not searched
        const uint32_t Fn = __shfl_sync(FULL, F, nv);
        const int good = __popc(~Fn & ((nrem >= 32u ? 0xffffffffu : ((1u << nrem) - 1u)) << (nlo - vbase)));
        if (SAMPLE) {
          // the nodes of the last level: forbidden values fail, every other value is an accepted leaf
          const unsigned hp = s_hash(lo, hi);
          for (int b = nlo - vbase; b < nlo - vbase + (int)nrem; b++) {
            const bool bad = (Fn >> b) & 1u;
            if (!sample_hit(a, hp, nv, vbase + b + zb, bad)) continue;
            s_record(SAMPLE_COUNTED | (bad ? SAMPLE_FAILED : SAMPLE_LEAF), nv, vbase + b, lo, hi,
                     lane == nv ? vbase + b : lo, lane == nv ? vbase + b : hi);
          }
        }
        n32 += nrem; c32 += nrem - (unsigned)good; sols += (unsigned)good;
        props += (lane == nv && nrem > 1u) ? (unsigned)good : 0u;  // each of those nodes narrows nv to its value (branch-free:
                                                                   // a divergent branch here kept the warp split far into the loop)
      } else {
        // push: everything stays in shared memory / registers. When the frame has no value left to try after this one
        // (BITS: what is left of its interval is forbidden -- failed nodes, counted here) the child takes its place
        // instead of going on top of it: no pop back into an exhausted frame, no reload of it.
        const bool last_value = BITS ? avail == 0u : rem == 0u;
        if (BITS && last_value) {
          if (SAMPLE) s_skipped(cur - vbase, cur - vbase + (int)rem, var, plo, phi);
          n32 += rem; c32 += rem;
        }
        int *nf = last_value ? sf : sf + sfw;
        CHK(nf >= sst && nf + sfw <= sst + V * sfw, "lov push frame", (nf - sst) / sfw);
        if (lane == 0) {
          if (!last_value) reinterpret_cast<int2 *>(sf)[0] = make_int2(cur, (int)rem);
          reinterpret_cast<int4 *>(nf)[0] = make_int4(nlo, (int)nrem, nv, flevel + 1);
          reinterpret_cast<int2 *>(nf)[2] = make_int2((int)nmask, (int)fhash);
        }
        if (act) reinterpret_cast<int2 *>(nf + 8)[lane] = make_int2(lo, hi);
        if (BITS && act) nf[8 + 2 * V + lane] = (int)F;
        amask = nmask;
        plo = lo; phi = hi; pF = F;
        if (BITS) {
          fvarF = __shfl_sync(FULL, F, nv);
          avail = ~fvarF & ((nrem >= 32u ? 0xffffffffu : ((1u << nrem) - 1u)) << (nlo - vbase));
        }
        var = nv; cur = nlo; rem = nrem;
        flevel = flevel + 1;
        level += last_value ? 0 : 1;
        sf = nf;
        __syncwarp();
      }
    }

    // every POLL_NODES nodes -- every 4 while a warp of this block is waiting for work (shared-memory flag: no L2 trip)
#ifdef CSOLVE_OLD_POLL
    if (!EXPAND && (++poll & (*reinterpret_cast<volatile int *>(&s_blk_hungry) > 0 ? 3u : (unsigned)(POLL_NODES - 1))) == 0) {
#else
    if (!EXPAND && (++poll & 3u) == 0 &&
        ((poll & (unsigned)(POLL_NODES - 1)) == 0 || *reinterpret_cast<volatile int *>(&s_blk_hungry) > 0)) {
#endif
      nodes += n32; cuts += c32; n32 = 0; c32 = 0;
      dbg_polls++;
      if (a.n_peers > 0) comm_poll(a, lane);
      if (a.sink_headroom > 0 && *reinterpret_cast<volatile int *>(&ctl->n_stored) > a.max_solutions - a.sink_headroom) {
        if (lane == 0) atomicMax(&ctl->signal, SIG_SLICE_END);      // the solution buffer is nearly full: let the host drain it
        break;
      }
      if (*reinterpret_cast<volatile int *>(&ctl->signal) != SIG_RUN) break;
      if (clock64() - t0 > a.slice_cycles) {
        if (lane == 0) atomicMax(&ctl->signal, SIG_SLICE_END);
        break;
      }
      int tgt = -1;
      if (level >= base && (tgt = donation_target(a, lane)) >= 0) {
        // shallowest frame with at least two untried values (the top frame's cursor is in registers)
        // BITS: only values that are not forbidden yet count -- both halves get real work
        dbg_wanted++;
        int L = -1, d_cur = 0;
        unsigned d_rem = 0, keep = 0;
        if (lane == 0) {
          // a frame below the top is given away whole (the warp still has the deeper levels); the top frame is halved
          for (int q = base; q <= level; ++q) {
            const int *qf = sst + q * sfw;
            const bool top = q == level;
            const unsigned rm = top ? rem : (unsigned)qf[1];
            if (rm < (top ? 2u : 1u)) continue;
            const int cq = top ? cur : qf[0];
            unsigned kp = top ? (rm - 1u) / 2u + 1u : 0u;
            if (BITS) {
              const uint32_t av = top ? avail : (~(uint32_t)qf[8 + 2 * V + qf[2]] & ((rm >= 32u ? 0xffffffffu : ((1u << rm) - 1u)) << (cq - vbase)));
              const int cnt = __popc(av);
              if (cnt < (top ? 2 : 1)) continue;
              if (top) kp = (unsigned)((int)__fns(av, 0, cnt / 2) - (cq - vbase)) + 1u;    // up to the (cnt / 2)-th allowed value
            }
            L = q; d_rem = rm; d_cur = cq; keep = kp;
            break;
          }
        }
        L = __shfl_sync(FULL, L, 0);
        if (L >= 0) {
          d_cur = __shfl_sync(FULL, d_cur, 0); d_rem = __shfl_sync(FULL, d_rem, 0); keep = __shfl_sync(FULL, keep, 0);
          // this warp keeps the lower part [d_cur, d_cur + keep), the donated frame owns the rest
          const unsigned give = d_rem - keep;
          const int glo = (int)((unsigned)d_cur + keep);
          CHK(give >= 1u && give <= 32u && keep <= 32u, "lov donate give", give);
          int *own = sst + L * sfw;
          for (;;) {
            const int slot = reserve_slot(a, lane, tgt);
            int *g = ring_frame(a, tgt, slot);
            frame_out(own, g);
            __syncwarp();
            if (lane == 0) {
              __stcg(&g[FR_ITER], 0); __stcg(&g[FR_LO], glo + zb); __stcg(&g[FR_HI], (int)((unsigned)glo + give - 1u) + zb);
              __stcg(&g[FR_LAST], (int)(give - 1u));
            }
            if (publish_slot(a, lane, tgt, slot)) break;
          }
          if (lane == 0) { own[0] = d_cur; own[1] = (int)keep; }
          if (L == level) {
            rem = keep;
            const unsigned e = (unsigned)(cur - vbase) + keep;          // first bit that no longer belongs to the frame
            if (BITS && e < 32u) avail &= (1u << e) - 1u;
          }
          dbg_donated++;
          __syncwarp();
        }
        donation_done(a, lane, tgt);
      }
    }
  }
  nodes += n32; cuts += c32;

  if (!EXPAND && level >= base) {
    // park: the stack goes back to HBM for k_rebalance / the next slice
    if (have && lane == 0) { sst[level * sfw] = cur; sst[level * sfw + 1] = (int)rem; }
    __syncwarp();
    for (int L = base; L <= level; ++L) frame_out(sst + L * sfw, stack + (size_t)L * fw);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) props += __shfl_xor_sync(FULL, props, o);
  if (lane == 0) {
    a.wstate[gw].level = level;
    a.wstate[gw].base = base;
    a.wstate[gw].claim_base = cl.base; a.wstate[gw].claim_mask = cl.mask;
    unsigned long long *c = a.wcount + (size_t)gw * CNT_WIDTH;
    c[CNT_NODES] += nodes; c[CNT_CUTS] += cuts; c[CNT_PROPS] += props;
    c[CNT_VISITS] += BITS ? (unsigned long long)visits * (unsigned)V : visits; c[CNT_SOLUTIONS] += sols;
    c[CNT_WAIT] += (unsigned long long)waited; c[CNT_CLAIMS] += claims;
    c[CNT_POLLS] += dbg_polls; c[CNT_WANTED] += dbg_wanted; c[CNT_DONATED] += dbg_donated;
    c[CNT_LASTWORK] = (unsigned long long)(lastwork >= 0 ? lastwork : clock64() - t0);
  }
}

__global__ void __launch_bounds__(THREADS_PER_BLOCK)
k_propagate_batch_lov(const DevModel m, int n_nodes, const int32_t *dom_in, const int32_t *var, const int32_t *val,
                      int32_t *dom_out, uint8_t *failed) {
  extern __shared__ __align__(16) int smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const LovTables T = stage_lov(m, smem);
  const int V = m.n_vars;
  const bool act = lane < V;
  const int n_warps = gridDim.x * WARPS_PER_BLOCK;
  for (int b = blockIdx.x * WARPS_PER_BLOCK + wib; b < n_nodes; b += n_warps) {
    int2 dj = make_int2(0, 0);
    if (act) dj = __ldg(reinterpret_cast<const int2 *>(dom_in + (size_t)b * 2 * V) + lane);
    int lo = dj.x, hi = dj.y;
    const int x = var[b];
    if (lane == x && lo != hi) { lo = val[b]; hi = val[b]; }
    unsigned props = 0, visits = 0;
    bool ok;
    if (m.lov_bits) {
      // forbidden-value sets of the incoming state, then the transition
      if (!act) { lo = m.lov_vbase; hi = m.lov_vbase; }
      uint32_t F = act ? T.fconst[lane] : 0u;
      for (int i = 0; i < V; i++) {
        const int li = __shfl_sync(FULL, dj.x, i), hi_i = __shfl_sync(FULL, dj.y, i);
        if (li == hi_i && act) F |= lov_forbid(T.pair[i * 32 + lane], li, m.lov_vbase);
      }
      ok = lov_fixpoint_bits(T, V, m.lov_vbase, lane, lo, hi, F, 1u << x, props, visits);
    } else {
      ok = lov_fixpoint(T, V, m.n_lov_cval > 0, lane, lo, hi, 1u << x, props, visits);
    }
    ok = ok && !__any_sync(FULL, act && lo > hi);
    if (act) reinterpret_cast<int2 *>(dom_out + (size_t)b * 2 * V)[lane] = make_int2(lo, hi);
    if (lane == 0) failed[b] = ok ? 0 : 1;
  }
}

// =====================================================================================================
// K variables per lane: pure NOT(EQ) networks with 33..128 variables whose values fit a 32-value window
// (sudoku: 81 cells, values 1..9). Lane j owns variables j, j + 32, ... in registers together with their
// forbidden-value sets; same fixpoint as k_search_lov<.,true>. The DFS stack stays in HBM (frames carry the
// value sets behind the domains); levels whose variable already is a value are counted and skipped: in this
// form their propagation provably changes nothing (the variable's forbidden values were distributed when it
// became a value).
template <int K>
struct LovK {
  int lo[K], hi[K];
  uint32_t F[K];
};

template <int K>
__device__ __forceinline__ bool lovk_fixpoint(const DevModel &m, int lane, LovK<K> &x, uint32_t (&pend)[K],
                                              unsigned &props, unsigned &visits) {
  const int V = m.n_vars, vbase = m.lov_vbase;
  for (;;) {
    uint32_t any = 0;
#pragma unroll
    for (int q = 0; q < K; q++) any |= pend[q];
    if (any == 0u) break;
    // the forbidden values of ALL variables that just became a value are distributed (one shuffle and K table words
    // each), then every lane trims its K variables once -- a sudoku node fixes many cells per round
#pragma unroll
    for (int q = 0; q < K; q++) {
      uint32_t pq = pend[q];
      while (pq) {
        const int bit = __ffs((int)pq) - 1;
        pq &= pq - 1;
        const int i = q * 32 + bit;
        const int w = __shfl_sync(FULL, x.lo[q], bit);
        if (m.lov_adj_only) {
          // all-different style network: one broadcast word per 32 variables says who shares a clause with variable i,
          // and the only value they lose is w itself
          const uint32_t wbit = 1u << (w - vbase);
#pragma unroll
          for (int q2 = 0; q2 < K; q2++)
            x.F[q2] |= ((__ldg(&m.lov_adj[i * K + q2]) >> lane) & 1u) ? wbit : 0u;
        } else {
#pragma unroll
          for (int q2 = 0; q2 < K; q2++)     // v >= V: zero entries of the table
            x.F[q2] |= lov_forbid(__ldg(&m.lov_pair[(size_t)i * (32 * K) + lane + 32 * q2]), w, vbase);
        }
        visits += (unsigned)V;
      }
    }
    bool dead = false;
    unsigned changed = 0;
#pragma unroll
    for (int q = 0; q < K; q++) {
      const bool act = lane + 32 * q < V;
      const bool was = x.lo[q] == x.hi[q];
      const int olo = x.lo[q], ohi = x.hi[q];
      if (!lov_trim(x.F[q], vbase, x.lo[q], x.hi[q])) dead = true;
      pend[q] = __ballot_sync(FULL, act && !was && x.lo[q] == x.hi[q]);
      changed += (x.lo[q] != olo || x.hi[q] != ohi) ? 1u : 0u;
    }
    if (__any_sync(FULL, dead)) return false;
    props += changed;       // per lane; the callers that report it sum over the warp
  }
  return true;
}

template <int K>
__device__ __forceinline__ void lovk_load(const DevModel &m, const int *frame, int lane, LovK<K> &x) {
  const int dofs = frame_dom_offset(m.mask_words);
#pragma unroll
  for (int q = 0; q < K; q++) {
    const int v = lane + 32 * q;
    int2 d = make_int2(m.lov_vbase, m.lov_vbase);
    uint32_t F = 0;
    if (v < m.n_vars) {
      d = __ldcg(reinterpret_cast<const int2 *>(frame + dofs) + v);
      F = (uint32_t)__ldcg(&frame[dofs + 2 * m.n_vars + v]);
    }
    x.lo[q] = d.x; x.hi[q] = d.y; x.F[q] = F;
  }
}
template <int K>
__device__ __forceinline__ void lovk_store(const DevModel &m, int *frame, int lane, const LovK<K> &x) {
  const int dofs = frame_dom_offset(m.mask_words);
#pragma unroll
  for (int q = 0; q < K; q++) {
    const int v = lane + 32 * q;
    if (v < m.n_vars) {
      __stcg(reinterpret_cast<int2 *>(frame + dofs) + v, make_int2(x.lo[q], x.hi[q]));
      __stcg(&frame[dofs + 2 * m.n_vars + v], (int)x.F[q]);
    }
  }
}

// next branching variable among those without a level; -1 if none. amask: uniform words.
template <int K>
__device__ __forceinline__ int lovk_select(const DevModel &m, int lane, const LovK<K> &x, const uint32_t (&amask)[K],
                                           int order, int level1) {
  if (order == CSOLVE_ORDER_NONE) return level1 < m.n_vars ? __ldg(&m.order[level1]) : -1;
  unsigned long long bestk = ~0ull;
  int bestv = 0x7fffffff;
#pragma unroll
  for (int q = 0; q < K; q++) {
    const int v = lane + 32 * q;
    if (v >= m.n_vars || (amask[q] >> lane) & 1u) continue;
    unsigned primary;
    switch (order) {
    case CSOLVE_ORDER_SMALLEST_DOMAIN: primary = (unsigned)x.hi[q] - (unsigned)x.lo[q]; break;
    case CSOLVE_ORDER_LARGEST_DOMAIN:  primary = ~((unsigned)x.hi[q] - (unsigned)x.lo[q]); break;
    case CSOLVE_ORDER_SMALLEST_VALUE:  primary = (unsigned)x.lo[q] ^ 0x80000000u; break;
    default:                           primary = ~((unsigned)x.hi[q] ^ 0x80000000u); break;
    }
    const unsigned secondary = ~((unsigned)__ldg(&m.prio[v]) ^ 0x80000000u);
    const unsigned long long key = ((unsigned long long)primary << 32) | secondary;
    if (key < bestk) { bestk = key; bestv = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ok2 = __shfl_xor_sync(FULL, bestk, o);
    const int ov = __shfl_xor_sync(FULL, bestv, o);
    if (ok2 < bestk || (ok2 == bestk && ov < bestv)) { bestk = ok2; bestv = ov; }
  }
  return bestv == 0x7fffffff ? -1 : bestv;
}

// bounds of variable v (warp-uniform) from the register file of its owner lane
template <int K>
__device__ __forceinline__ int2 lovk_bounds(const LovK<K> &x, int v) {
  int l = x.lo[0], h = x.hi[0];
#pragma unroll
  for (int q = 1; q < K; q++) if ((v >> 5) == q) { l = x.lo[q]; h = x.hi[q]; }
  return make_int2(__shfl_sync(FULL, l, v & 31), __shfl_sync(FULL, h, v & 31));
}

// forbidden-value set of variable v (warp-uniform) from its owner lane
template <int K>
__device__ __forceinline__ uint32_t lovk_F(const LovK<K> &x, int v) {
  uint32_t f = x.F[0];
#pragma unroll
  for (int q = 1; q < K; q++) if ((v >> 5) == q) f = x.F[q];
  return __shfl_sync(FULL, f, v & 31);
}

#ifndef CSOLVE_LOVK_MIN_BLOCKS
#define CSOLVE_LOVK_MIN_BLOCKS 3
#endif
template <bool EXPAND, int K, bool SAMPLE = false>
__global__ void __launch_bounds__(THREADS_PER_BLOCK, CSOLVE_LOVK_MIN_BLOCKS)
k_search_lovk(const SearchArgs a) {
  const DevModel &m = a.m;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * WARPS_PER_BLOCK + wib;
  __shared__ int s_blk_hungry;       // warps of this block waiting for a frame: their neighbours poll for donations faster
  if (threadIdx.x == 0) s_blk_hungry = 0;
  __syncthreads();
  if (gw >= a.n_warps) return;
  const int V = m.n_vars, fw = m.frame_words;
  int *stack = a.stacks + (size_t)gw * (V + 1) * fw;
  SearchCtl *ctl = a.ctl;

  int level = a.wstate[gw].level, base = a.wstate[gw].base;
  Claim cl; cl.base = a.wstate[gw].claim_base; cl.mask = a.wstate[gw].claim_mask; cl.drained = false;
  unsigned long long nodes = 0, cuts = 0, sols = 0;
  unsigned props = 0, visits = 0, poll = 0;
  const long long t0 = clock64();
  bool have = false, hungry = false;
  int var = 0, flo = 0, fhi = 0, flevel = 0, ftag = 0;
  unsigned iter = 0, last = 0, fhash = 0;
  uint32_t amask[K];
  uint32_t fvarF = 0;              // forbidden-value set of the top frame's branching variable (warp-uniform)
  LovK<K> P;                       // top frame: state before this level's assignment
#pragma unroll
  for (int q = 0; q < K; q++) { amask[q] = 0; P.lo[q] = P.hi[q] = m.lov_vbase; P.F[q] = 0; }

  // parity instrumentation (SAMPLE instances only)
  auto s_hash = [&](const LovK<K> &z) {
    unsigned h = 0;
#pragma unroll
    for (int q = 0; q < K; q++)
      if (lane + 32 * q < V) h ^= mix_hash(0x9E3779B9u + (unsigned)(lane + 32 * q), (unsigned)z.lo[q], (unsigned)z.hi[q]);
    return __reduce_xor_sync(FULL, h);
  };
  auto s_record = [&](int flags, int svar, int sval, const LovK<K> &z0, const LovK<K> &z1) {
    int *r = sample_begin(a, lane, flags, svar, sval, 0);
    if (r == nullptr) return;
#pragma unroll
    for (int q = 0; q < K; q++) {
      const int v = lane + 32 * q;
      if (v < V) {
        reinterpret_cast<int2 *>(r + 4)[v] = make_int2(z0.lo[q], z0.hi[q]);
        reinterpret_cast<int2 *>(r + 4 + 2 * V)[v] = make_int2(z1.lo[q], z1.hi[q]);
      }
    }
  };

  for (;;) {
    if (level < base) {
      const int *src;
      int ring_slot = 0;
      if (EXPAND) {
        int it = 0;
        if (lane == 0) it = atomicAdd(&ctl->item_next, 1);
        it = __shfl_sync(FULL, it, 0);
        if (it >= ctl->item_count) break;
        src = a.items + (size_t)it * fw;
      } else {
        const int slot = claim_frame<true>(a, lane, hungry, cl, &s_blk_hungry, t0);
        if (slot < 0) break;
        src = claimed_frame(a, slot);
        ring_slot = slot;
      }
      const int L = EXPAND ? 0 : __ldcg(&src[FR_LEVEL]);
      int *dst = stack + (size_t)L * fw;
      for (int w = lane; w < fw; w += 32) __stcg(&dst[w], __ldcg(&src[w]));
      if (!EXPAND) {
        __syncwarp();
        if (lane == 0 && ring_slot >= a.n_initial) { __threadfence(); __stcg(&a.ready[ring_slot], 0); }
      }
      level = base = L;
      have = false;
      __syncwarp();
    }

    int *f = stack + (size_t)level * fw;
    if (!have) {
      const int4 h0 = __ldcg(reinterpret_cast<const int4 *>(f));
      const int4 h1 = __ldcg(reinterpret_cast<const int4 *>(f) + 1);
#pragma unroll
      for (int q = 0; q < K; q++) amask[q] = (uint32_t)__ldcg(&f[FR_MASK + q]);
      lovk_load<K>(m, f, lane, P);
      fvarF = lovk_F<K>(P, h0.x);
      var = h0.x; iter = (unsigned)h0.y; last = (unsigned)h0.z; flo = h0.w;
      fhi = h1.x; flevel = h1.y; ftag = h1.z; fhash = (unsigned)h1.w;
      have = true;
    }

    if (EXPAND && last >= (unsigned)a.expand_branch_max) {
      int slot = 0;
      if (lane == 0) { slot = atomicAdd(&ctl->out_count, 1); atomicAdd(&ctl->passed, 1); }
      slot = __shfl_sync(FULL, slot, 0);
      if (slot < a.out_cap) {
        int *g = a.items_out + (size_t)slot * fw;
        for (int w = lane; w < fw; w += 32) __stcg(&g[w], __ldcg(&f[w]));
      } else if (lane == 0) {
        atomicAdd(&ctl->out_dropped, 1);
      }
      level = base - 1; have = false;
      continue;
    }
    if (iter > last) { level--; have = false; continue; }

    // ---- one search node -------------------------------------------------------------------------
    const int val = step_value(flo, fhi, iter);
    iter++;
    if ((fvarF >> (val - m.lov_vbase)) & 1u) {      // already forbidden: a failed node (see k_search_lov)
      if (SAMPLE && sample_hit(a, s_hash(P), var, val, true)) s_record(SAMPLE_FAILED | SAMPLE_COUNTED, var, val, P, P);
      nodes++; cuts++;
      continue;
    }
    LovK<K> x = P;
    uint32_t pend[K];
#pragma unroll
    for (int q = 0; q < K; q++) {
      pend[q] = 0;
      if ((var >> 5) == q) {
        pend[q] = 1u << (var & 31);
        if (lane == (var & 31)) { x.lo[q] = val; x.hi[q] = val; }
      }
    }
    const bool ok = lovk_fixpoint<K>(m, lane, x, pend, props, visits);
    nodes++;
    bool s_hit = false;
    unsigned s_hx = 0;
    if (SAMPLE) {
      s_hit = sample_hit(a, s_hash(P), var, val, !ok);
      if (!ok && s_hit) s_record(SAMPLE_FAILED, var, val, P, x);
      if (ok) s_hx = s_hash(x);
    }

    if (!ok) {
      cuts++;
    } else {
      // levels of variables that already are a value: one node each, nothing to propagate (see above)
      uint32_t am[K];
#pragma unroll
      for (int q = 0; q < K; q++) am[q] = amask[q] | (((var >> 5) == q) ? (1u << (var & 31)) : 0u);
      int lev = flevel + 1;
      int nv = -1;
      int2 nb = make_int2(0, 0);
      unsigned hsh = mix_hash(fhash, (unsigned)var, (unsigned)val);
      if (a.order == CSOLVE_ORDER_SMALLEST_DOMAIN) {
        // in this order every variable that is a value and has no level yet is selected before any other one: all
        // their levels are counted at once (a sudoku node fixes 5..20 cells; selecting them one by one cost a 64-bit
        // warp reduction each). The path hash folds them in order-independently.
        unsigned cnt = 0, fold = 0;
#pragma unroll
        for (int q = 0; q < K; q++) {
          const int v = lane + 32 * q;
          const bool fixed = v < V && !((am[q] >> lane) & 1u) && x.lo[q] == x.hi[q];
          const unsigned mk = __ballot_sync(FULL, fixed);
          if (SAMPLE) {
            for (unsigned rest = mk; rest; rest &= rest - 1u) {
              const int bit = __ffs((int)rest) - 1;
              const int fv = __shfl_sync(FULL, x.lo[q], bit);
              if (sample_hit(a, s_hx, bit + 32 * q, fv, false)) s_record(SAMPLE_COUNTED, bit + 32 * q, fv, x, x);
            }
          }
          cnt += (unsigned)__popc(mk);
          am[q] |= mk;
          fold ^= fixed ? mix_hash(0x9E3779B9u, (unsigned)v, (unsigned)x.lo[q]) : 0u;
        }
        if (cnt) {
          nodes += cnt;
          lev += (int)cnt;
          hsh = mix_hash(hsh, __reduce_xor_sync(FULL, fold), cnt);
        }
      }
      while (lev < V) {
        nv = lovk_select<K>(m, lane, x, am, a.order, lev);
        nb = lovk_bounds<K>(x, nv);
        if (nb.x != nb.y) break;
        if (SAMPLE && sample_hit(a, s_hx, nv, nb.x, false)) s_record(SAMPLE_COUNTED, nv, nb.x, x, x);
        nodes++;                                   // the level of a fixed variable
        hsh = mix_hash(hsh, (unsigned)nv, (unsigned)nb.x);
#pragma unroll
        for (int q = 0; q < K; q++) if ((nv >> 5) == q) am[q] |= 1u << (nv & 31);
        lev++;
      }
      if (lev >= V) {
        // leaf: every variable is a value and none of them is forbidden by the others
        bool good = true;
#pragma unroll
        for (int q = 0; q < K; q++)
          if (lane + 32 * q < V && (x.lo[q] != x.hi[q] || ((x.F[q] >> (x.lo[q] - m.lov_vbase)) & 1u))) good = false;
        const bool leaf_ok = __all_sync(FULL, good);
        if (SAMPLE && s_hit) s_record(leaf_ok ? SAMPLE_LEAF : 0, var, val, P, x);
        if (leaf_ok) {
          bool accepted = true;
          if (m.objective == CSOLVE_OBJ_ANY) {
            int old = 0;
            if (lane == 0) old = atomicMax(&ctl->signal, SIG_STOP);
            old = __shfl_sync(FULL, old, 0);
            accepted = old != SIG_STOP;
            if (accepted && a.n_peers > 0) comm_push_stop(a, lane);
          }
          if (accepted) {
            sols++;
            if (a.inst_solutions != nullptr && lane == 0) atomicAdd(&a.inst_solutions[ftag], 1u);
            int slot = 0;
            if (lane == 0) slot = atomicAdd(&ctl->n_stored, 1);
            slot = __shfl_sync(FULL, slot, 0);
            if (slot < a.max_solutions) {
              int *dst = a.solbuf + (size_t)slot * (V + 1);
#pragma unroll
              for (int q = 0; q < K; q++) if (lane + 32 * q < V) dst[lane + 32 * q] = x.lo[q];
              if (lane == 0) dst[V] = ftag;
            }
            __syncwarp();     // reconverge after the predicated stores (see k_search_lov)
          }
        }
      } else {
        if (SAMPLE && s_hit) s_record(0, var, val, P, x);
        int *g;
        if (EXPAND) {
          int slot = 0;
          if (lane == 0) slot = atomicAdd(&ctl->out_count, 1);
          slot = __shfl_sync(FULL, slot, 0);
          g = slot < a.out_cap ? a.items_out + (size_t)slot * fw : nullptr;
          if (g == nullptr && lane == 0) atomicAdd(&ctl->out_dropped, 1);
        } else {
          if (lane == 0) __stcg(&f[FR_ITER], (int)iter);
          g = stack + (size_t)(level + 1) * fw;
        }
        const unsigned nlast = (unsigned)nb.y - (unsigned)nb.x;
        if (g != nullptr) {
          if (lane == 0) {
            __stcg(reinterpret_cast<int4 *>(g), make_int4(nv, 0, (int)nlast, nb.x));
            __stcg(reinterpret_cast<int4 *>(g) + 1, make_int4(nb.y, lev, ftag, (int)hsh));
          }
          if (lane < K) {
            uint32_t mv = am[0];
#pragma unroll
            for (int q = 1; q < K; q++) if (lane == q) mv = am[q];
            __stcg(&g[FR_MASK + lane], (int)mv);
          }
          lovk_store<K>(m, g, lane, x);
        }
        if (!EXPAND) {
#pragma unroll
          for (int q = 0; q < K; q++) amask[q] = am[q];
          P = x;
          fvarF = lovk_F<K>(x, nv);
          var = nv; flo = nb.x; fhi = nb.y; iter = 0; last = nlast;
          flevel = lev; fhash = hsh;
          level++;
          __syncwarp();
        }
      }
    }

    if (!EXPAND && (++poll & (*reinterpret_cast<volatile int *>(&s_blk_hungry) > 0 ? 1u : 7u)) == 0) {
      if (a.n_peers > 0) comm_poll(a, lane);
      if (a.sink_headroom > 0 && *reinterpret_cast<volatile int *>(&ctl->n_stored) > a.max_solutions - a.sink_headroom) {
        if (lane == 0) atomicMax(&ctl->signal, SIG_SLICE_END);      // the solution buffer is nearly full: let the host drain it
        break;
      }
      if (*reinterpret_cast<volatile int *>(&ctl->signal) != SIG_RUN) break;
      if (clock64() - t0 > a.slice_cycles) {
        if (lane == 0) atomicMax(&ctl->signal, SIG_SLICE_END);
        break;
      }
      int tgt = -1;
      if (level >= base && (tgt = donation_target(a, lane)) >= 0) {
        int L = -1;
        unsigned d_iter = 0; int d_lo = 0, d_hi = 0;
        if (lane == 0) {
          for (int q = base; q <= level; ++q) {
            unsigned it2, la2; int lo2, hi2;
            if (q == level) { it2 = iter; la2 = last; lo2 = flo; hi2 = fhi; }
            else {
              const int4 g0 = __ldcg(reinterpret_cast<const int4 *>(stack + (size_t)q * fw));
              it2 = (unsigned)g0.y; la2 = (unsigned)g0.z; lo2 = g0.w; hi2 = __ldcg(&stack[(size_t)q * fw + FR_HI]);
            }
            // a frame below the top is given away whole (the warp still has the deeper levels); the top frame is halved
            if (it2 <= la2 && (q < level || la2 - it2 >= 1)) { L = q; d_iter = it2; d_lo = lo2; d_hi = hi2; break; }
          }
        }
        L = __shfl_sync(FULL, L, 0);
        if (L >= 0) {
          d_iter = __shfl_sync(FULL, d_iter, 0); d_lo = __shfl_sync(FULL, d_lo, 0); d_hi = __shfl_sync(FULL, d_hi, 0);
          const long long ua = (long long)d_lo + ((d_iter + 1) >> 1), ub = (long long)d_hi - (d_iter >> 1);
          const long long mid = L < level ? ua - 1 : ua + (ub - ua) / 2;
          int *own = stack + (size_t)L * fw;
          for (;;) {
            const int slot = reserve_slot(a, lane, tgt);
            int *g = ring_frame(a, tgt, slot);
            for (int w = lane; w < fw; w += 32) __stcg(&g[w], __ldcg(&own[w]));
            __syncwarp();
            if (lane == 0) {
              __stcg(&g[FR_ITER], 0); __stcg(&g[FR_LO], (int)(mid + 1)); __stcg(&g[FR_HI], (int)ub);
              __stcg(&g[FR_LAST], (int)(unsigned)(ub - mid - 1));
            }
            if (publish_slot(a, lane, tgt, slot)) break;
          }
          if (lane == 0) {
            if (L < level) { __stcg(&own[FR_ITER], 1); __stcg(&own[FR_LAST], 0); }     // exhausted: iter > last
            else {
              __stcg(&own[FR_ITER], 0); __stcg(&own[FR_LO], (int)ua); __stcg(&own[FR_HI], (int)mid);
              __stcg(&own[FR_LAST], (int)(unsigned)(mid - ua));
            }
          }
          if (L == level) { iter = 0; flo = (int)ua; fhi = (int)mid; last = (unsigned)(mid - ua); }
          __syncwarp();
        }
        donation_done(a, lane, tgt);
      }
    }
  }

#pragma unroll
  for (int o = 16; o > 0; o >>= 1) props += __shfl_xor_sync(FULL, props, o);      // counted per lane in lovk_fixpoint
  if (lane == 0) {
    if (have && level >= base) __stcg(&stack[(size_t)level * fw + FR_ITER], (int)iter);
    a.wstate[gw].level = level;
    a.wstate[gw].base = base;
    a.wstate[gw].claim_base = cl.base; a.wstate[gw].claim_mask = cl.mask;
    unsigned long long *c = a.wcount + (size_t)gw * CNT_WIDTH;
    c[CNT_NODES] += nodes; c[CNT_CUTS] += cuts; c[CNT_PROPS] += props;
    c[CNT_VISITS] += visits; c[CNT_SOLUTIONS] += sols;
  }
}

// node transitions (parity hook) and batched root frames for the K-variables-per-lane form
template <int K>
__device__ __forceinline__ void lovk_from_domains(const DevModel &m, const int32_t *dom, int lane, LovK<K> &x, uint32_t (&singles)[K]) {
  const int V = m.n_vars;
#pragma unroll
  for (int q = 0; q < K; q++) {
    const int v = lane + 32 * q;
    int2 d = make_int2(m.lov_vbase, m.lov_vbase);
    if (v < V) d = __ldg(reinterpret_cast<const int2 *>(dom) + v);
    x.lo[q] = d.x; x.hi[q] = d.y;
    x.F[q] = v < V ? __ldg(&m.lov_fconst[v]) : 0u;
    singles[q] = __ballot_sync(FULL, v < V && d.x == d.y);
  }
}

template <int K>
__global__ void __launch_bounds__(THREADS_PER_BLOCK)
k_propagate_batch_lovk(const DevModel m, int n_nodes, const int32_t *dom_in, const int32_t *var, const int32_t *val,
                       int32_t *dom_out, uint8_t *failed) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int V = m.n_vars;
  const int n_warps = gridDim.x * WARPS_PER_BLOCK;
  for (int b = blockIdx.x * WARPS_PER_BLOCK + wib; b < n_nodes; b += n_warps) {
    LovK<K> x; uint32_t pend[K];
    lovk_from_domains<K>(m, dom_in + (size_t)b * 2 * V, lane, x, pend);
    unsigned props = 0, visits = 0;
    // value sets of the incoming state (which the search always leaves at a fixpoint): the values of the
    // variables that are a value are distributed WITHOUT trimming; then the decision is propagated
    for (int i = 0; i < V; i++) {
      bool single_i = false;
#pragma unroll
      for (int q = 0; q < K; q++) if ((i >> 5) == q) single_i = (pend[q] >> (i & 31)) & 1u;
      if (!single_i) continue;
      const int2 bi = lovk_bounds<K>(x, i);
#pragma unroll
      for (int q = 0; q < K; q++) {
        const int v = lane + 32 * q;
        if (v < V) x.F[q] |= lov_forbid(__ldg(&m.lov_pair[(size_t)i * (32 * K) + v]), bi.x, m.lov_vbase);
      }
    }
    bool ok = true;
    const int xv = var[b];
#pragma unroll
    for (int q = 0; q < K; q++) {
      pend[q] = 0;
      if ((xv >> 5) == q) {
        pend[q] = 1u << (xv & 31);
        if (lane == (xv & 31) && x.lo[q] != x.hi[q]) { x.lo[q] = val[b]; x.hi[q] = val[b]; }
      }
    }
    if (ok) ok = lovk_fixpoint<K>(m, lane, x, pend, props, visits);
#pragma unroll
    for (int q = 0; q < K; q++)
      if (lane + 32 * q < V) reinterpret_cast<int2 *>(dom_out + (size_t)b * 2 * V)[lane + 32 * q] = make_int2(x.lo[q], x.hi[q]);
    if (lane == 0) failed[b] = ok ? 0 : 1;
  }
}

template <int K>
__global__ void __launch_bounds__(THREADS_PER_BLOCK)
k_root_frames_lovk(const DevModel m, int n_roots, const int32_t *root_dom, int order, int32_t *frames_out, int out_cap,
                   int32_t *n_out, unsigned char *root_failed) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int V = m.n_vars, fw = m.frame_words;
  const int n_warps = gridDim.x * WARPS_PER_BLOCK;
  for (int r = blockIdx.x * WARPS_PER_BLOCK + wib; r < n_roots; r += n_warps) {
    LovK<K> x; uint32_t pend[K];
    lovk_from_domains<K>(m, root_dom + (size_t)r * 2 * V, lane, x, pend);
    unsigned props = 0, visits = 0;
    const bool ok = lovk_fixpoint<K>(m, lane, x, pend, props, visits);   // root phase: every value is distributed
    if (lane == 0) root_failed[r] = ok ? 0 : 1;
    if (!ok) continue;
    uint32_t am[K];
#pragma unroll
    for (int q = 0; q < K; q++) am[q] = 0;
    // the first level (levels of variables that already are a value are skipped by the search kernel, not here:
    // the root frame must exist even when everything is fixed)
    const int nv = lovk_select<K>(m, lane, x, am, order, 0);
    const int2 nb = lovk_bounds<K>(x, nv);
    int slot = 0;
    if (lane == 0) slot = atomicAdd(n_out, 1);
    slot = __shfl_sync(FULL, slot, 0);
    if (slot >= out_cap) continue;             // the host checks the capacity before the launch; never write past the pool
    int *g = frames_out + (size_t)slot * fw;
    if (lane == 0) {
      __stcg(reinterpret_cast<int4 *>(g), make_int4(nv, 0, (int)((unsigned)nb.y - (unsigned)nb.x), nb.x));
      __stcg(reinterpret_cast<int4 *>(g) + 1, make_int4(nb.y, 0, r, (int)mix_hash(0x1234567u, (unsigned)r, 0u)));
    }
    if (lane < K) __stcg(&g[FR_MASK + lane], 0);
    lovk_store<K>(m, g, lane, x);
  }
}

// =====================================================================================================
// "Bit state" search kernel for pure SAT models (DevModel::sat: every variable 0/1, every clause a disjunction of at
// most three literals -- what scripts/cnf2csolve produces, BASELINE config 5).
//
// On such a model an interval is one of [0,0], [1,1], [0,1], and the reference's contractors are unit propagation
// (contract_lits above; src/propagate.c:320-376). A node's whole domain vector is two bit vectors -- assigned, value --
// that live in REGISTERS: lane w holds word w of each (V <= 1024). Nothing is copied per node:
//   * a decision and everything it implies is appended to the warp's trail (shared memory, 16-bit literal codes) and
//     the trail itself is the propagation queue: four pending assignments are expanded per step, eight lanes each,
//     every lane looking at one clause in which the assignment falsified a literal (the clause's two OTHER literals
//     come with the occurrence record: two shuffles each tell their state); unit literals found by the 32 lanes are
//     gathered with a ballot and applied one after the other (two lanes may force the same variable, also to
//     different values: that is the reference's empty intersection, src/propagate.c:57-66);
//   * a failed node restores the two words from the copy taken before the decision (registers); only when a decision
//     has no value left is the trail undone entry by entry down to the previous decision;
//   * the DFS stack is a list of decision records (variable, trail position, level, values left): 8 bytes a level.
// Levels, node counts and the order of decisions are exactly those of k_search (static order: a level per variable in
// DevModel::order, levels of variables that already are a value are one never-failing node each -- counted here by
// position, without executing them; with failure-driven priorities: highest priority among the open variables).
// The HBM frame format is spoken at the edges only: frames claimed from the frontier / donation ring are turned into
// bit vectors, a donated or parked decision is turned into a frame (the state before it is replayed from the trail).
// Parked frames are a private LIFO in the warp's HBM stack: a resumed warp pops them like claimed frames.
static const int SAT_FRAME_BITS = 0x5A7B175;      // header word 6 of a frame whose domain area holds the two bit vectors
#ifndef SAT_DONATE_MIN
#define SAT_DONATE_MIN 32
#endif
struct SatRec { unsigned short var, trail, pos; unsigned char nxt, cnt; };     // 8 bytes: one decision level

__host__ __device__ __forceinline__ int sat_table_words(const DevModel &m) {
  int o = (m.n_vars * 2 + 3) / 4;            // order16
  o += 2 * m.mask_words;                     // root_asg, root_val
  o = (o + 1) & ~1;
  if (m.sat_smem_bytes > 0) {
    o += 2 * m.n_vars + 1; o = (o + 1) & ~1;
    o += 2 * m.n_sat_occ;
  }
  return (o + 3) & ~3;
}
__host__ __device__ __forceinline__ int sat_warp_words(const DevModel &m) {
  return (((m.n_vars * 2 + 3) / 4 + 1) & ~1) + 3 * (m.n_vars + 1);      // trail (16-bit), decision records, entry counters
}

template <bool SAMPLE>
__global__ void __launch_bounds__(THREADS_PER_BLOCK, 4)
k_search_sat(const SearchArgs a) {
  extern __shared__ __align__(16) int smem[];
  const DevModel &m = a.m;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int gw = blockIdx.x * WARPS_PER_BLOCK + wib;
  __shared__ int s_blk_hungry;
  const int V = m.n_vars, W = m.mask_words, fw = m.frame_words;
  const int dofs = frame_dom_offset(W);

  // ---- block: tables into shared memory -------------------------------------------------------------------------
  unsigned short *order16 = reinterpret_cast<unsigned short *>(smem);
  int o = (V * 2 + 3) / 4;
  unsigned *root_asg = reinterpret_cast<unsigned *>(smem + o), *root_val = root_asg + W;
  o += 2 * W; o = (o + 1) & ~1;
  const int *optr = m.sat_occ_ptr;
  const int2 *occ = reinterpret_cast<const int2 *>(m.sat_occ);
  if (threadIdx.x == 0) s_blk_hungry = 0;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    const int rem = V - (i << 5);
    root_asg[i] = rem >= 32 ? 0u : ~((1u << rem) - 1u);     // bits beyond V count as assigned
    root_val[i] = 0u;
  }
  for (int i = threadIdx.x; i < V; i += blockDim.x) order16[i] = (unsigned short)__ldg(&m.order[i]);
  if (m.sat_smem_bytes > 0) {
    int *sp = smem + o;
    for (int i = threadIdx.x; i <= 2 * V; i += blockDim.x) sp[i] = __ldg(&m.sat_occ_ptr[i]);
    o += 2 * V + 1; o = (o + 1) & ~1;
    int2 *so = reinterpret_cast<int2 *>(smem + o);
    for (int i = threadIdx.x; i < m.n_sat_occ; i += blockDim.x) so[i] = __ldg(reinterpret_cast<const int2 *>(m.sat_occ) + i);
    optr = sp; occ = so;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const int lo = __ldg(&m.root_dom[2 * i]), hi = __ldg(&m.root_dom[2 * i + 1]);
    if (lo == hi) { atomicOr(&root_asg[i >> 5], 1u << (i & 31)); if (lo == 1) atomicOr(&root_val[i >> 5], 1u << (i & 31)); }
  }
  __syncthreads();
  if (gw >= a.n_warps) return;

  // ---- warp --------------------------------------------------------------------------------------------------------
  int *wbase = smem + sat_table_words(m) + wib * sat_warp_words(m);
  unsigned short *trail = reinterpret_cast<unsigned short *>(wbase);
  SatRec *drec = reinterpret_cast<SatRec *>(wbase + (((V * 2 + 3) / 4 + 1) & ~1));
  unsigned *dentry = reinterpret_cast<unsigned *>(drec + (V + 1));     // node counter when a level's current value was entered
  int *stack = a.stacks + (size_t)gw * (V + 1) * fw;
  SearchCtl *ctl = a.ctl;
  const bool dynamic = a.gprio != nullptr;

  unsigned asg = lane < W ? root_asg[lane] : 0xffffffffu, val = lane < W ? root_val[lane] : 0u;
  unsigned pa = asg, pv = val;             // state before the top decision (valid while `snap`)
  unsigned ba = asg, bv = val;             // state the current frame was loaded with (trail position 0)
  bool snap = false;
  int depth = -1, tlen = 0;
  int tvar = 0, ttrail = 0, tpos = 0, tnxt = 0, tcnt = 0;      // top decision record
  unsigned n32 = 0;                        // nodes since the kernel started (32 bits are plenty inside one slice)
  // private LIFO of parked frames: stack[pbase .. parked] (empty: parked < pbase). k_rebalance hands an idle warp ONE frame,
  // at the index it had in the donor's stack (level = base = that index): what lies below `base` in this warp's stack is
  // left over from earlier slices and must not be searched again (found by the CPU emulation of this kernel: ALL-mode
  // counts of a time-sliced search came out too high)
  int parked = a.wstate[gw].level, pbase = a.wstate[gw].base;
  Claim cl; cl.base = a.wstate[gw].claim_base; cl.mask = a.wstate[gw].claim_mask; cl.drained = false;
  unsigned long long nodes = 0, cuts = 0, sols = 0, cuts_reported = 0;
  unsigned props = 0, visits = 0, poll = 0;
  const long long t0 = clock64();
  long long waited = 0, lastwork = -1;
  unsigned claims = 0, dbg_donated = 0;
  bool hungry = false;

  // state of literal code c (var << 1 | negated): 0 open, 1 true, 2 false (two shuffles; c < 0: "false")
  auto lit_state = [&](int c) {
    const int v = c >= 0 ? c >> 1 : 0;
    const unsigned aw = __shfl_sync(FULL, asg, v >> 5), vw = __shfl_sync(FULL, val, v >> 5);
    if (c < 0) return 2;
    if (!((aw >> (v & 31)) & 1u)) return 0;
    return (int)((vw >> (v & 31)) & 1u) != (c & 1) ? 1 : 2;
  };
  // bit vectors -> lo,hi pairs of variables lane, lane + 32, ... written to dst (2 * V words)
  auto write_domains = [&](unsigned xa, unsigned xv, int *dst, bool global) {
    for (int k = 0; k < W; k++) {
      const unsigned aw = __shfl_sync(FULL, xa, k), vw = __shfl_sync(FULL, xv, k);
      const int v = (k << 5) + lane;
      if (v < V) {
        const int b = (int)((vw >> lane) & 1u);
        const int2 d = ((aw >> lane) & 1u) ? make_int2(b, b) : make_int2(0, 1);
        if (global) __stcg(reinterpret_cast<int2 *>(dst) + v, d); else reinterpret_cast<int2 *>(dst)[v] = d;
      }
    }
  };
  // state after the first `upto` trail entries of the current frame
  auto prefix_state = [&](int upto, unsigned &xa, unsigned &xv) {
    xa = ba; xv = bv;
    for (int i = 0; i < upto; i++) {
      const int l = trail[i];
      if (lane == (l >> 6)) { xa |= 1u << ((l >> 1) & 31); if (l & 1) xv |= 1u << ((l >> 1) & 31); }
    }
  };
  // a decision level with values left -> HBM frame (device_model.h): the state before the decision, its open values
  auto frame_out = [&](int var, int trail_pos, int pos, int nxt, int cnt, int *g) {
    unsigned xa, xv;
    prefix_state(trail_pos, xa, xv);
    if (lane == 0) {
      __stcg(reinterpret_cast<int4 *>(g), make_int4(var, 0, cnt - 1, nxt));
      __stcg(reinterpret_cast<int4 *>(g) + 1, make_int4(nxt + cnt - 1, pos, 0, 0));
    }
    for (int w = lane; w < W; w += 32) __stcg(&g[FR_MASK + w], 0);
    write_domains(xa, xv, g + dofs, true);
    __syncwarp();
  };
  // the same for a frame that only travels through a donation ring to another warp of this kernel: the two bit vectors
  // instead of 2 * V domain words (header word 6 says so) -- 56 bytes instead of 1.6 KB for 200 variables
  auto frame_out_bits = [&](int var, int trail_pos, int pos, int nxt, int cnt, int *g) {
    unsigned xa, xv;
    prefix_state(trail_pos, xa, xv);
    if (lane == 0) {
      __stcg(reinterpret_cast<int4 *>(g), make_int4(var, 0, cnt - 1, nxt));
      __stcg(reinterpret_cast<int4 *>(g) + 1, make_int4(nxt + cnt - 1, pos, SAT_FRAME_BITS, 0));
    }
    if (lane < W) { __stcg(&g[dofs + lane], (int)xa); __stcg(&g[dofs + W + lane], (int)xv); }
    __syncwarp();
  };
  // identity hash of a state (SAMPLE): same function of the domains as the other kernels'
  auto state_hash = [&](unsigned xa, unsigned xv) {
    unsigned h = 0;
    for (int k = 0; k < W; k++) {
      const unsigned aw = __shfl_sync(FULL, xa, k), vw = __shfl_sync(FULL, xv, k);
      const int v = (k << 5) + lane;
      if (v < V) {
        const int b = (int)((vw >> lane) & 1u);
        const bool as = (aw >> lane) & 1u;
        h ^= mix_hash(0x9E3779B9u + (unsigned)v, (unsigned)(as ? b : 0), (unsigned)(as ? b : 1));
      }
    }
    return __reduce_xor_sync(FULL, h);
  };
  auto s_record = [&](int flags, int svar, int sval, unsigned xa0, unsigned xv0, unsigned xa1, unsigned xv1) {
    int *r = sample_begin(a, lane, flags, svar, sval, 0);
    if (r == nullptr) return;
    write_domains(xa0, xv0, r + 4, false);
    write_domains(xa1, xv1, r + 4 + 2 * V, false);
    __syncwarp();
  };

  for (;;) {
    // ---- poll: stop / slice end / somebody waiting for work ---------------------------------------------------------
    // (every 16 nodes; every 4 while a warp of this block waits for work: an L2 round trip per node would halve the node rate)
    if ((++poll & 3u) == 0 && ((poll & 15u) == 0 || *reinterpret_cast<volatile int *>(&s_blk_hungry) > 0)) {
      if (a.n_peers > 0) comm_poll(a, lane);
      if (a.fail_limit > 0 && restart_due(a, lane, cuts, cuts_reported)) break;
      if (a.sink_headroom > 0 && *reinterpret_cast<volatile int *>(&ctl->n_stored) > a.max_solutions - a.sink_headroom) {
        if (lane == 0) atomicMax(&ctl->signal, SIG_SLICE_END);      // the solution buffer is nearly full: let the host drain it
        break;
      }
      if (*reinterpret_cast<volatile int *>(&ctl->signal) != SIG_RUN) break;
      if (clock64() - t0 > a.slice_cycles) {
        if (lane == 0) atomicMax(&ctl->signal, SIG_SLICE_END);
        break;
      }
      int tgt = -1;
      if (depth >= 0 && (tgt = donation_target(a, lane)) >= 0) {
        // Shallowest decision level with a value left: given away whole (a 0/1 level has at most one value besides
        // the one being searched). Only if this warp has already spent SAT_DONATE_MIN nodes below that level's current
        // value: sibling sub-trees are of similar size, and handing over one that is finished in a handful of nodes
        // costs both warps more than it saves (measured on the unsatisfiable seed 1: 61 M hand-offs of 7 nodes each,
        // three quarters of all warp time spent waiting).
        int L = -1, dv = 0, dt = 0, dp = 0, dn = 0, dc = 0;
        for (int q = 0; q < depth && L < 0; q++) {
          const SatRec r = drec[q];
          if (r.cnt >= 1) {
            if (n32 - dentry[q] >= (unsigned)SAT_DONATE_MIN) { L = q; dv = r.var; dt = r.trail; dp = r.pos; dn = r.nxt; dc = r.cnt; }
            break;
          }
        }
        if (L >= 0) {
          for (;;) {
            const int slot = reserve_slot(a, lane, tgt);
            frame_out_bits(dv, dt, dp, dn, dc, ring_frame(a, tgt, slot));
            if (publish_slot(a, lane, tgt, slot)) break;
          }
          if (lane == 0) drec[L].cnt = 0;
          dbg_donated++;
          __syncwarp();
        }
        donation_done(a, lane, tgt);
      }
    }

    // ---- no decision level: take a frame (parked ones first) --------------------------------------------------------
    if (depth < 0) {
      const int *src;
      int ring_slot = -1;
      if (parked >= pbase) {
        src = stack + (size_t)parked * fw;
        parked--;
      } else {
        const long long w0 = clock64();
        lastwork = w0 - t0;
        const int slot = claim_frame<true>(a, lane, hungry, cl, &s_blk_hungry, t0);
        waited += clock64() - w0;
        if (slot < 0) break;
        claims++;
        src = claimed_frame(a, slot);
        ring_slot = slot;
      }
      const int4 h0 = __ldcg(reinterpret_cast<const int4 *>(src));
      const int4 h1 = __ldcg(reinterpret_cast<const int4 *>(src) + 1);
      if (h1.z == SAT_FRAME_BITS) {
        if (lane < W) { asg = (unsigned)__ldcg(&src[dofs + lane]); val = (unsigned)__ldcg(&src[dofs + W + lane]); }
      } else {
        for (int k = 0; k < W; k++) {
          const int v = (k << 5) + lane;
          int2 d = make_int2(0, 0);
          if (v < V) d = __ldcg(reinterpret_cast<const int2 *>(src + dofs) + v);
          const unsigned xa = __ballot_sync(FULL, d.x == d.y), xv = __ballot_sync(FULL, d.x == d.y && d.x == 1);
          if (lane == k) { asg = xa; val = xv; }
        }
      }
      if (lane >= W) { asg = 0xffffffffu; val = 0u; }
      // The level's variable is taken as open whatever the frame's domains say: k_rebalance writes the interval a
      // half owns into them, and re-propagating a variable that already was a value changes nothing (the state is a
      // fixpoint that contains it).
      if (lane == (h0.x >> 5)) { asg &= ~(1u << (h0.x & 31)); val &= ~(1u << (h0.x & 31)); }
      if (ring_slot >= a.n_initial) {
        __syncwarp();
        if (lane == 0) { __threadfence(); __stcg(&a.ready[ring_slot], 0); }
      }
      __syncwarp();
      ba = asg; bv = val; pa = asg; pv = val; snap = true;
      tlen = 0; depth = 0;
      const unsigned it = (unsigned)h0.y, la = (unsigned)h0.z;
      tvar = h0.x; ttrail = 0; tpos = h1.y;
      tnxt = step_value(h0.w, h1.x, it);
      tcnt = it <= la ? (int)(la - it + 1u) : 0;
      continue;
    }

    // ---- top level has no value left: back to the level below ----------------------------------------------------
    if (tcnt == 0) {
      depth--;
      if (depth >= 0) {
        const SatRec r = drec[depth];
        tvar = r.var; ttrail = r.trail; tpos = r.pos; tnxt = r.nxt; tcnt = r.cnt;
      }
      snap = false;
      continue;
    }
    if (!snap) {
      // undo the trail down to the state before this level's decision
      for (int i = tlen - 1; i >= ttrail; i--) {
        const int l = trail[i];
        if (lane == (l >> 6)) { asg &= ~(1u << ((l >> 1) & 31)); val &= ~(1u << ((l >> 1) & 31)); }
      }
      tlen = ttrail;
      pa = asg; pv = val; snap = true;
    }

    // ---- one search node: tvar := b, unit propagation to fixpoint (src/csolve.c:444-457) ---------------------------
    const int b = tnxt;
    tnxt = b + 1; tcnt--;
    nodes++; n32++;
    bool ok = true;
    int fail_var = -1;
    {
      const unsigned aw = __shfl_sync(FULL, asg, tvar >> 5);
      const unsigned bit = 1u << (tvar & 31);
      if (aw & bit) {
        // the level of a variable that already is a value (a frame of the expanded frontier): nothing to propagate
        const unsigned vw = __shfl_sync(FULL, val, tvar >> 5);
        ok = (int)((vw >> (tvar & 31)) & 1u) == b;
      } else {
        if (lane == (tvar >> 5)) { asg |= bit; if (b) val |= bit; }
        if (lane == 0) trail[tlen] = (unsigned short)((tvar << 1) | b);
        tlen++;
        __syncwarp();
        int qh = tlen - 1;
        while (ok && qh < tlen) {
          const int li = qh + (lane >> 3);
          const int ev = li < tlen ? (int)trail[li] : -1;       // assignment event var << 1 | value
          int i = 0, e = 0;
          if (ev >= 0) { i = optr[ev] + (lane & 7); e = optr[ev + 1]; }
          qh = min(qh + 4, tlen);
          while (__any_sync(FULL, i < e)) {
            const bool act = i < e;
            int2 r = make_int2(-1, -1);
            if (act) r = occ[i];
            i += 8;
            const int s0 = lit_state(r.x), s1 = lit_state(r.y);
            int unit = -1;
            bool conf = false;
            if (act && s0 != 1 && s1 != 1) {
              if (s0 == 2 && s1 == 2) conf = true;
              else if (s0 == 0 && s1 == 2) unit = r.x;
              else if (s0 == 2 && s1 == 0) unit = r.y;
            }
            visits += act ? 1u : 0u;
            const unsigned cm = __ballot_sync(FULL, conf);
            if (cm) { ok = false; fail_var = __shfl_sync(FULL, ev, __ffs((int)cm) - 1) >> 1; break; }
            unsigned um = __ballot_sync(FULL, unit >= 0);
            while (um) {
              const int src = __ffs((int)um) - 1;
              um &= um - 1u;
              const int u = __shfl_sync(FULL, unit, src);
              const int v = u >> 1, want = (u & 1) ^ 1;
              const unsigned aw2 = __shfl_sync(FULL, asg, v >> 5), vw2 = __shfl_sync(FULL, val, v >> 5);
              const unsigned bit2 = 1u << (v & 31);
              if (aw2 & bit2) {
                if ((int)((vw2 >> (v & 31)) & 1u) != want) { ok = false; fail_var = v; break; }      // forced to 0 and to 1
                continue;
              }
              if (lane == (v >> 5)) { asg |= bit2; if (want) val |= bit2; }
              if (lane == 0) trail[tlen] = (unsigned short)((v << 1) | want);
              tlen++;
              props++;
            }
            if (!ok) break;
          }
          __syncwarp();
        }
      }
    }
    if (dynamic && lane == 0) {
      // prio-- on success, prio++ on failure, and ++ for the variable whose clauses failed (src/csolve.c:459-462, src/propagate.c:44-54)
      atomicAdd(&a.gprio[tvar], ok ? -1 : 1);
      if (!ok && fail_var >= 0) atomicAdd(&a.gprio[fail_var], 1);
    }
    unsigned s_hp = 0;
    bool s_hit = false;
    if (SAMPLE) {
      s_hp = state_hash(pa, pv);
      s_hit = sample_hit(a, s_hp, tvar, b, !ok);
    }
    if (!ok) {
      if (SAMPLE && s_hit) s_record(SAMPLE_FAILED, tvar, b, pa, pv, asg, val);
      cuts++;
      asg = pa; val = pv; tlen = ttrail;
      continue;
    }

    // ---- the next level ------------------------------------------------------------------------------------------------
    int nv = -1, npos = V;
    if (!dynamic) {
      // first position after tpos whose variable is still open; the positions passed are levels of variables that
      // are a value: one never-failing node each
      for (int p0 = tpos + 1; p0 < V; p0 += 32) {
        const int p = p0 + lane;
        const int v = p < V ? (int)order16[p] : 0;
        const unsigned aw = __shfl_sync(FULL, asg, v >> 5);
        const unsigned open = __ballot_sync(FULL, p < V && !((aw >> (v & 31)) & 1u));
        if (open) { npos = p0 + __ffs((int)open) - 1; break; }
      }
      if (npos < V) nv = order16[npos];
      const int passed = npos - tpos - 1;
      if (SAMPLE) {
        // the node just executed, then the levels counted in bulk; the last level of a complete assignment is the leaf
        const bool leaf_here = npos >= V && passed == 0;
        if (s_hit) s_record(leaf_here ? SAMPLE_LEAF : 0, tvar, b, pa, pv, asg, val);
        if (passed > 0) {
          const unsigned hs = state_hash(asg, val);
          for (int p = tpos + 1; p < npos; p++) {
            const int v = order16[p];
            const int fv = (int)((__shfl_sync(FULL, val, v >> 5) >> (v & 31)) & 1u);
            if (sample_hit(a, hs, v, fv, false)) s_record(SAMPLE_COUNTED | ((npos >= V && p == V - 1) ? SAMPLE_LEAF : 0), v, fv, asg, val, asg, val);
          }
        }
      }
      nodes += (unsigned)passed;
    } else {
      // highest failure-driven priority among the open variables, ties: lower index (the levels of the variables this
      // node made a value are counted here: one node each)
      unsigned long long bestk = ~0ull;
      for (int k = 0; k < W; k++) {
        const unsigned aw = __shfl_sync(FULL, asg, k);
        const int v = (k << 5) + lane;
        if (v < V && !((aw >> lane) & 1u)) {
          const unsigned key = ~((unsigned)__ldcg(&a.gprio[v]) ^ 0x80000000u);
          const unsigned long long kk = ((unsigned long long)key << 32) | (unsigned)v;
          if (kk < bestk) bestk = kk;
        }
      }
#pragma unroll
      for (int q = 16; q > 0; q >>= 1) {
        const unsigned long long ok2 = __shfl_xor_sync(FULL, bestk, q);
        if (ok2 < bestk) bestk = ok2;
      }
      if (bestk != ~0ull) { nv = (int)(unsigned)bestk; npos = 0; }
      if (SAMPLE && s_hit) s_record(nv < 0 ? SAMPLE_LEAF : 0, tvar, b, pa, pv, asg, val);
      nodes += (unsigned)(tlen - ttrail > 0 ? tlen - ttrail - 1 : 0);
    }

    if (nv < 0) {
      // every variable is a value and no clause failed: an accepted leaf (src/csolve.c:222-244)
      bool accepted = true;
      if (m.objective == CSOLVE_OBJ_ANY) {
        int old = 0;
        if (lane == 0) old = atomicMax(&ctl->signal, SIG_STOP);
        old = __shfl_sync(FULL, old, 0);
        accepted = old != SIG_STOP;
        if (accepted && a.n_peers > 0) comm_push_stop(a, lane);
      }
      if (accepted) {
        sols++;
        int slot = 0;
        if (lane == 0) slot = atomicAdd(&ctl->n_stored, 1);
        slot = __shfl_sync(FULL, slot, 0);
        if (slot < a.max_solutions) {
          int *dst = a.solbuf + (size_t)slot * (V + 1);
          for (int k = 0; k < W; k++) {
            const unsigned vw = __shfl_sync(FULL, val, k);
            const int v = (k << 5) + lane;
            if (v < V) dst[v] = (int)((vw >> lane) & 1u);
          }
          if (lane == 0) dst[V] = 0;
        }
        __syncwarp();
      }
      asg = pa; val = pv; tlen = ttrail;          // the leaf's level is done: on with the next value of this one
      continue;
    }
    // push
    if (lane == 0) {
      SatRec r; r.var = (unsigned short)tvar; r.trail = (unsigned short)ttrail; r.pos = (unsigned short)tpos;
      r.nxt = (unsigned char)tnxt; r.cnt = (unsigned char)tcnt;
      drec[depth] = r;
      dentry[depth] = n32;
    }
    depth++;
    tvar = nv; ttrail = tlen; tpos = npos; tnxt = 0; tcnt = 2;
    pa = asg; pv = val; snap = true;
    __syncwarp();
  }

  // ---- park: every level with values left becomes a frame of the private LIFO (shallowest first) -----------------
  if (parked < pbase) { parked = -1; pbase = 0; }      // empty: start again at the bottom of the stack
  if (depth >= 0) {
    if (lane == 0) {
      SatRec r; r.var = (unsigned short)tvar; r.trail = (unsigned short)ttrail; r.pos = (unsigned short)tpos;
      r.nxt = (unsigned char)tnxt; r.cnt = (unsigned char)tcnt;
      drec[depth] = r;
    }
    __syncwarp();
    for (int q = 0; q <= depth; q++) {
      const SatRec r = drec[q];
      if (r.cnt == 0) continue;
      parked++;
      frame_out(r.var, r.trail, r.pos, r.nxt, r.cnt, stack + (size_t)parked * fw);
    }
  }
#pragma unroll
  for (int q = 16; q > 0; q >>= 1) visits += __shfl_xor_sync(FULL, visits, q);
  if (lane == 0) {
    a.wstate[gw].level = parked;
    a.wstate[gw].base = pbase;
    a.wstate[gw].claim_base = cl.base; a.wstate[gw].claim_mask = cl.mask;
    unsigned long long *c = a.wcount + (size_t)gw * CNT_WIDTH;
    c[CNT_NODES] += nodes; c[CNT_CUTS] += cuts; c[CNT_PROPS] += props;
    c[CNT_VISITS] += visits; c[CNT_SOLUTIONS] += sols;
    c[CNT_WAIT] += (unsigned long long)waited; c[CNT_CLAIMS] += claims; c[CNT_DONATED] += dbg_donated;
    c[CNT_LASTWORK] = (unsigned long long)(lastwork >= 0 ? lastwork : clock64() - t0);
  }
}

// ---- rebalance -----------------------------------------------------------------------------------
// One block. Idle warps (level < base) are paired with busy warps that own a frame with at
// least two untried values; the donor keeps the lower half of the untried interval, the idle
// warp gets the upper half (both restart their value iteration over the new bounds).
// scratch: [0] n_idle, [1] n_donor, then idle list [n_warps], donor list [2 * n_warps].
__global__ void __launch_bounds__(1024)
k_rebalance(const SearchArgs a, int32_t *scratch) {
  const DevModel &m = a.m;
  const int V = m.n_vars, fw = m.frame_words, nw = a.n_warps;
  int *idle_list = scratch + 4;
  int *donor_list = idle_list + nw;
  __shared__ int n_idle, n_donor, idle_done, busy, moved;
  if (threadIdx.x == 0) {
    n_idle = 0; n_donor = 0; idle_done = 0; busy = 0; moved = 0;
    if (a.n_peers > 0 && a.peer_ready[a.rank] != nullptr) {
      // ranks of a comm: the peers serve this rank's ring while its search kernel runs; the ring is closed (and the
      // visitors already inside waited for) before its counters are looked at and rebased below
      atomicExch_system(&a.comm->ring_open, 0);
      __threadfence_system();
      while (*reinterpret_cast<volatile int *>(&a.comm->inflight) > 0) __nanosleep(200);
      __threadfence_system();
    }
  }
  __syncthreads();
  for (int w = threadIdx.x; w < nw; w += blockDim.x) {
    if (a.wstate[w].level >= a.wstate[w].base || a.wstate[w].claim_mask != 0u) atomicAdd(&busy, 1);   // claimed root-frontier frames are work too
    else idle_list[atomicAdd(&n_idle, 1)] = w;
  }
  __syncthreads();
  // every waiting warp has left the search kernel: tickets no donor served are void (their slots carry the -1 marker)
  {
    const int served = a.ctl->item_count, taken = a.ctl->item_next;
    const unsigned ring = (unsigned)(a.pool_cap - a.n_initial);
    for (int i = served + (int)threadIdx.x; i - taken < 0; i += blockDim.x) a.ready[a.n_initial + (int)((unsigned)i % ring)] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
      // unserved tickets are void; both counters are rebased to the ring size so that the ticket -> slot mapping
      // (ticket % ring) survives any number of slices without the 32-bit counters wrapping
      const int cnt = taken - served > 0 ? taken : served;
      const int shift = (int)((unsigned)taken / ring * ring);
      a.ctl->item_count = cnt - shift;
      a.ctl->item_next = taken - shift;
    }
    __syncthreads();
  }
  // frames still in the frontier pool are work too: nothing to move while they last
  const bool pool_left = a.ctl->item_count - a.ctl->item_next > 0 || *reinterpret_cast<volatile int *>(&a.front_ctl->init_next) < a.n_initial;
  for (int round = 0; round < 4 && !pool_left; ++round) {
    if (idle_done >= n_idle) break;
    if (threadIdx.x == 0) n_donor = 0;
    __syncthreads();
    for (int w = threadIdx.x; w < nw; w += blockDim.x) {
      const int lv = a.wstate[w].level, bs = a.wstate[w].base;
      if (lv < bs) continue;
      const int *stack = a.stacks + (size_t)w * (V + 1) * fw;
      for (int L = bs; L <= lv; ++L) {
        const int *f = stack + (size_t)L * fw;
        const unsigned iter = (unsigned)f[FR_ITER], last = (unsigned)f[FR_LAST];
        if (iter <= last && last - iter >= 1) {   // at least two untried values
          const int k = atomicAdd(&n_donor, 1);
          donor_list[2 * k] = w; donor_list[2 * k + 1] = L;
          break;
        }
      }
    }
    __syncthreads();
    const int pairs = min(n_idle - idle_done, n_donor);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int k = wid; k < pairs; k += nwarp) {
      const int thief = idle_list[idle_done + k];
      const int donor = donor_list[2 * k], L = donor_list[2 * k + 1];
      int *df = a.stacks + (size_t)donor * (V + 1) * fw + (size_t)L * fw;
      int *tf = a.stacks + (size_t)thief * (V + 1) * fw + (size_t)L * fw;
      const unsigned iter = (unsigned)df[FR_ITER];
      const long long lo = df[FR_LO], hi = df[FR_HI];
      const long long ua = lo + ((iter + 1) >> 1), ub = hi - (iter >> 1);   // untried interval
      const long long mid = ua + (ub - ua) / 2;
      __syncwarp();
      for (int w = lane; w < fw; w += 32) tf[w] = df[w];
      __syncwarp();
      if (lane == 0) {
        const int var = df[FR_VAR];
        const int dofs = frame_dom_offset(m.mask_words);
        df[FR_ITER] = 0; df[FR_LO] = (int)ua; df[FR_HI] = (int)mid; df[FR_LAST] = (int)(unsigned)(mid - ua);
        df[dofs + 2 * var] = (int)ua; df[dofs + 2 * var + 1] = (int)mid;
        tf[FR_ITER] = 0; tf[FR_LO] = (int)(mid + 1); tf[FR_HI] = (int)ub; tf[FR_LAST] = (int)(unsigned)(ub - mid - 1);
        tf[dofs + 2 * var] = (int)(mid + 1); tf[dofs + 2 * var + 1] = (int)ub;
        a.wstate[thief].level = L; a.wstate[thief].base = L;
        atomicAdd(&moved, 1); atomicAdd(&busy, 1);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) idle_done += pairs;
    __syncthreads();
    if (pairs == 0) break;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.ctl->busy = busy + (pool_left ? 1 : 0);
    a.ctl->moved = moved;
    a.ctl->hungry = 0;
    a.ctl->signal = a.ctl->signal == SIG_STOP ? SIG_STOP : SIG_RUN;
  }
}

// ---- frontier rebalancing between ranks ---------------------------------------------------------------
// Between two time slices (every warp parked): split up to max_frames frames off the parked stacks into `out`
// (frames of frame_words words) so that the host can ship them to a rank that ran dry. Same rule as the in-kernel
// donation: the shallowest frame of a busy warp that still has untried values -- whole if it lies below the warp's
// top frame, its upper half otherwise. One block.
__global__ void __launch_bounds__(1024)
k_export_frames(const SearchArgs a, int32_t *out, int max_frames, int32_t *n_out) {
  const DevModel &m = a.m;
  const int V = m.n_vars, fw = m.frame_words, nw = a.n_warps;
  __shared__ int n_taken;
  if (threadIdx.x == 0) n_taken = 0;
  __syncthreads();
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int w = wid; w < nw; w += nwarp) {
    const int lv = a.wstate[w].level, bs = a.wstate[w].base;
    if (lv < bs) continue;
    int *stack = a.stacks + (size_t)w * (V + 1) * fw;
    int L = -1;
    for (int q = bs; q <= lv && L < 0; ++q) {
      const unsigned iter = (unsigned)stack[(size_t)q * fw + FR_ITER], last = (unsigned)stack[(size_t)q * fw + FR_LAST];
      if (iter <= last && (q < lv || last - iter >= 1)) L = q;
    }
    if (L < 0) continue;                         // warp-uniform
    int k = 0;
    if (lane == 0) k = atomicAdd(&n_taken, 1);
    k = __shfl_sync(FULL, k, 0);
    if (k >= max_frames) break;
    int *df = stack + (size_t)L * fw;
    int *tf = out + (size_t)k * fw;
    const unsigned iter = (unsigned)df[FR_ITER];
    const long long lo = df[FR_LO], hi = df[FR_HI];
    const long long ua = lo + ((iter + 1) >> 1), ub = hi - (iter >> 1);   // untried interval
    const long long mid = L < lv ? ua - 1 : ua + (ub - ua) / 2;
    __syncwarp();
    for (int x = lane; x < fw; x += 32) tf[x] = df[x];
    __syncwarp();
    if (lane == 0) {
      const int var = df[FR_VAR];
      const int dofs = frame_dom_offset(m.mask_words);
      tf[FR_ITER] = 0; tf[FR_LO] = (int)(mid + 1); tf[FR_HI] = (int)ub; tf[FR_LAST] = (int)(unsigned)(ub - mid - 1);
      tf[dofs + 2 * var] = (int)(mid + 1); tf[dofs + 2 * var + 1] = (int)ub;
      if (L < lv) { df[FR_ITER] = 1; df[FR_LAST] = 0; }      // exhausted: iter > last
      else {
        df[FR_ITER] = 0; df[FR_LO] = (int)ua; df[FR_HI] = (int)mid; df[FR_LAST] = (int)(unsigned)(mid - ua);
        df[dofs + 2 * var] = (int)ua; df[dofs + 2 * var + 1] = (int)mid;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) *n_out = min(n_taken, max_frames);
}

// frames received from another rank go into the ring as tickets served ahead of their holders: the first warps that
// run dry in the next slice find them ready. One block.
__global__ void __launch_bounds__(1024)
k_import_frames(const SearchArgs a, const int32_t *in, int n_frames) {
  const int fw = a.m.frame_words;
  const unsigned ring = (unsigned)(a.pool_cap - a.n_initial);
  __shared__ int first;
  if (threadIdx.x == 0) first = atomicAdd(&a.ctl->item_count, n_frames);
  __syncthreads();
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int k = wid; k < n_frames; k += nwarp) {
    const int slot = a.n_initial + (int)((unsigned)(first + k) % ring);
    int *g = a.pool + (size_t)slot * fw;
    for (int x = lane; x < fw; x += 32) g[x] = in[(size_t)k * fw + x];
    __syncwarp();
    if (lane == 0) a.ready[slot] = 1;
  }
}

// ---- comm: a rank's state between two launches of its search kernel (one thread) ------------------------------------
// See CommBlock. A rank that has nothing left clears its busy mark (and takes itself out of rank 0's count of active
// ranks) -- and looks at its ring once more afterwards: a peer that reserved a ticket in it before the mark was cleared
// is seen then, one that comes later finds the mark cleared and counts the rank in again itself (publish_slot).
__global__ void k_comm_state(const SearchArgs a, int want_frames, int force_idle, int32_t *out) {
  CommBlock *root = a.peer_comm[0];
  auto has_work = [&]() {
    return *reinterpret_cast<volatile int *>(&a.ctl->item_count) - *reinterpret_cast<volatile int *>(&a.ctl->item_next) > 0 ||
           *reinterpret_cast<volatile int *>(&a.front_ctl->init_next) < a.n_initial;
  };
  auto close_ring = [&]() {
    atomicExch_system(&a.comm->ring_open, 0);
    __threadfence_system();
    while (*reinterpret_cast<volatile int *>(&a.comm->inflight) > 0) __nanosleep(200);
  };
  bool work = false;
  for (;;) {
    work = !force_idle && has_work();
    if (work || force_idle) {
      close_ring();                          // nobody is writing into the ring after this: the look below is exact
      work = !force_idle && has_work();
    }
    if (work) {
      if (atomicExch_system(&a.comm->busy_epoch, a.epoch) != a.epoch) atomicAdd_system(&root->active64, 1ull);
      break;
    }
    if (!force_idle) atomicExch_system(&a.comm->ring_open, a.epoch);
    if (atomicExch_system(&a.comm->busy_epoch, 0) == a.epoch) atomicAdd_system(&root->active64, ~0ull);
    __threadfence_system();
    if (force_idle || !has_work()) break;    // a ticket reserved before the mark was cleared is seen here
  }
  for (int r = 0; r < a.world; r++) {
    if (r != a.rank) *reinterpret_cast<volatile int *>(&a.peer_comm[r]->demand[a.rank]) = work ? 0 : want_frames;
  }
  const unsigned long long act = *reinterpret_cast<volatile unsigned long long *>(&root->active64);
  out[0] = work ? 1 : 0;
  out[1] = (int)(act >> 32) == a.epoch ? (int)(unsigned)act : -1;
  out[2] = *reinterpret_cast<volatile int *>(&a.comm->stop_epoch) == a.epoch ? 1 : 0;
}

// comm: a rank waits ON THE DEVICE for rank 0's frontier of this epoch (one thread, polling rank 0's block over NVLink
// about once a microsecond). The host then needs one copy of the block instead of a polling loop of copies: small
// synchronous copies out of a peer's memory were measured to serialise with whatever that peer's streams are doing --
// a rank polling that way saw the frontier only after rank 0 had finished the whole search.
__global__ void k_comm_wait_front(const CommBlock *root, int epoch, unsigned long long timeout_ns, CommBlock *copy) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (*reinterpret_cast<volatile const int *>(&root->front_epoch) - epoch < 0) {
    __nanosleep(500);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) break;
  }
  __threadfence_system();
  // everything rank 0 wrote before front_epoch (front_n, front_fw) is visible now: hand the host a consistent copy
  const volatile int *src = reinterpret_cast<const volatile int *>(root);
  int *dst = reinterpret_cast<int *>(copy);
  for (int i = 0; i < (int)(sizeof(CommBlock) / sizeof(int)); i++) dst[i] = src[i];
}

// sums the per-warp counters: thread t of a block always adds the same counter (the stride is a multiple of CNT_WIDTH),
// consecutive threads read consecutive words. `out` is zeroed by the caller. (One block of 256 threads walking the
// rows one after the other took 55-60 us for 4 736 warps -- of a search that takes 7 ms on eight GPUs.)
static const int REDUCE_THREADS = 85 * CNT_WIDTH;      // 1020
__global__ void __launch_bounds__(1024)
k_reduce_counters(const unsigned long long *wcount, int n_warps, unsigned long long *out) {
  __shared__ unsigned long long acc[CNT_WIDTH];
  if (threadIdx.x < CNT_WIDTH) acc[threadIdx.x] = 0;
  __syncthreads();
  const size_t total = (size_t)n_warps * CNT_WIDTH;
  unsigned long long sum = 0;
  if (threadIdx.x < REDUCE_THREADS)
    for (size_t i = (size_t)blockIdx.x * REDUCE_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * REDUCE_THREADS) sum += wcount[i];
  if (sum) atomicAdd(&acc[threadIdx.x % CNT_WIDTH], sum);
  __syncthreads();
  if (threadIdx.x < CNT_WIDTH && acc[threadIdx.x]) atomicAdd(&out[threadIdx.x], acc[threadIdx.x]);
}

// ---- parity hook: independent node transitions ------------------------------------------------------
__global__ void __launch_bounds__(THREADS_PER_BLOCK)
k_propagate_batch(const DevModel m, int n_nodes, const int32_t *dom_in, const int32_t *var, const int32_t *val,
                  const int32_t *best, int32_t *dom_out, uint8_t *failed) {
  extern __shared__ __align__(16) int smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int4 *wrec; const int *wptr;
  LinTables ltab;
  stage_table(m, smem, wrec, wptr, ltab);
  const int wwords = (warp_smem_words(m) + 3) & ~3;
  WarpSmem s = carve(m, smem + (m.table_smem_bytes >> 2) + wib * wwords);
  s.lt = ltab;
  const int V = m.n_vars;
  const int n_warps = gridDim.x * WARPS_PER_BLOCK;
  for (int b = blockIdx.x * WARPS_PER_BLOCK + wib; b < n_nodes; b += n_warps) {
    const int2 *src = reinterpret_cast<const int2 *>(dom_in + (size_t)b * 2 * V);
    for (int v = lane; v < V; v += 32) reinterpret_cast<int2 *>(s.d)[v] = __ldg(&src[v]);
    for (int w = lane; w < m.mask_words; w += 32) { s.cur[w] = 0; s.nxt[w] = 0; }
    __syncwarp();
    bool ok = true;
    if (lane == 0) {
      const int x = var[b];
      // a variable that already is a value is not re-bound (src/csolve.c:301-303)
      if (s.d[2 * x] != s.d[2 * x + 1]) { s.d[2 * x] = val[b]; s.d[2 * x + 1] = val[b]; }
      s.cur[x >> 5] |= 1u << (x & 31);
      if (m.obj_var >= 0) {
        Dom o; o.lo = s.d[2 * m.obj_var]; o.hi = s.d[2 * m.obj_var + 1];
        o = objective_tighten(m.objective, o, best[b]);
        s.d[2 * m.obj_var] = o.lo; s.d[2 * m.obj_var + 1] = o.hi;
        s.cur[m.obj_var >> 5] |= 1u << (m.obj_var & 31);
        ok = o.lo <= o.hi;
      }
    }
    ok = __shfl_sync(FULL, ok, 0);
    __syncwarp();
    unsigned props = 0, visits = 0;
    if (ok) ok = m.n_lin > 0 ? warp_fixpoint<false, true>(m, s, wrec, wptr, lane, props, visits)
                             : warp_fixpoint<false, false>(m, s, wrec, wptr, lane, props, visits);
    int2 *dst = reinterpret_cast<int2 *>(dom_out + (size_t)b * 2 * V);
    for (int v = lane; v < V; v += 32) dst[v] = reinterpret_cast<int2 *>(s.d)[v];
    if (lane == 0) failed[b] = ok ? 0 : 1;
    __syncwarp();
  }
}

// ---- batched roots: root phase on the device -----------------------------------------------------------
// Root r = a full domain vector over the shared network (e.g. one sudoku: the clue cells are single values).
// The warp propagates EVERY variable's watchers to fixpoint (what the reference's root sweeps do,
// src/propagate.c:474-485) and, if the root is consistent, emits its level-0 frame tagged with r.
__global__ void __launch_bounds__(THREADS_PER_BLOCK)
k_root_frames(const DevModel m, int n_roots, const int32_t *root_dom, int order, int32_t *frames_out, int out_cap,
              int32_t *n_out, unsigned char *root_failed) {
  extern __shared__ __align__(16) int smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int4 *wrec; const int *wptr;
  LinTables ltab;
  stage_table(m, smem, wrec, wptr, ltab);
  const int wwords = (warp_smem_words(m) + 3) & ~3;
  WarpSmem s = carve(m, smem + (m.table_smem_bytes >> 2) + wib * wwords);
  s.lt = ltab;
  const int V = m.n_vars, fw = m.frame_words;
  const int n_warps = gridDim.x * WARPS_PER_BLOCK;
  for (int r = blockIdx.x * WARPS_PER_BLOCK + wib; r < n_roots; r += n_warps) {
    const int2 *src = reinterpret_cast<const int2 *>(root_dom + (size_t)r * 2 * V);
    for (int v = lane; v < V; v += 32) reinterpret_cast<int2 *>(s.d)[v] = __ldg(&src[v]);
    for (int w = lane; w < m.mask_words; w += 32) {
      const int rem = V - (w << 5);
      s.cur[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
      s.nxt[w] = 0; s.amask[w] = 0;
    }
    __syncwarp();
    unsigned props = 0, visits = 0;
    const bool ok = m.n_lin > 0 ? warp_fixpoint<false, true>(m, s, wrec, wptr, lane, props, visits)
                                : warp_fixpoint<false, false>(m, s, wrec, wptr, lane, props, visits);
    if (lane == 0) root_failed[r] = ok ? 0 : 1;
    if (ok) {
      int nv = warp_select_var(m, s, lane, order, 0, -1);
      int slot = 0;
      if (lane == 0) slot = atomicAdd(n_out, 1);
      slot = __shfl_sync(FULL, slot, 0);
      // header word 6 = root id; the path hash starts from the root id so partitions spread the roots
      if (slot < out_cap)
        write_child_frame(m, s, frames_out + (size_t)slot * fw, lane, nv, 0, r, mix_hash(0x1234567u, (unsigned)r, 0u), -1);
    }
    __syncwarp();
  }
}

// ---- host-side launch wrappers -----------------------------------------------------------------------
// does the depth-first phase of this search run on the bit-state kernel?
bool search_uses_sat(const DevModel &m, bool learn, int order) {
  return m.sat && !learn && order == CSOLVE_ORDER_NONE && !m.lov && !m.lovk;
}

size_t search_smem_bytes(const DevModel &m, bool learn, bool sat) {
  if (sat) return (size_t)(sat_table_words(m) + WARPS_PER_BLOCK * sat_warp_words(m)) * sizeof(int);
  if (learn) {
    const int wwords = (warp_smem_words(m, true) + 3) & ~3;
    return (size_t)m.table_smem_bytes + (size_t)wwords * sizeof(int) * WARPS_PER_BLOCK;
  }
  if (m.lovk) return 0;
  if (m.lov) return (size_t)m.lov_smem_bytes + (size_t)WARPS_PER_BLOCK * m.n_vars * ((8 + 3 * m.n_vars + 3) & ~3) * sizeof(int);
  const int wwords = (4 * m.n_vars + 3 * m.mask_words + 3) & ~3;
  return (size_t)m.table_smem_bytes + (size_t)wwords * sizeof(int) * WARPS_PER_BLOCK;
}

template <bool SAMPLE>
static const void *lovk_kernel(bool expand, int K) {
  switch (K) {
  case 2: return expand ? (const void *)k_search_lovk<true, 2, SAMPLE> : (const void *)k_search_lovk<false, 2, SAMPLE>;
  case 3: return expand ? (const void *)k_search_lovk<true, 3, SAMPLE> : (const void *)k_search_lovk<false, 3, SAMPLE>;
  default: return expand ? (const void *)k_search_lovk<true, 4, SAMPLE> : (const void *)k_search_lovk<false, 4, SAMPLE>;
  }
}

template <bool SAMPLE>
static const void *lov_kernel(bool expand, bool bits) {
  if (expand) return bits ? (const void *)k_search_lov<true, true, SAMPLE> : (const void *)k_search_lov<true, false, SAMPLE>;
  return bits ? (const void *)k_search_lov<false, true, SAMPLE> : (const void *)k_search_lov<false, false, SAMPLE>;
}

template <bool SAMPLE>
static const void *general_kernel(bool expand, bool lin) {
  if (lin) return expand ? (const void *)k_search<true, false, true, SAMPLE> : (const void *)k_search<false, false, true, SAMPLE>;
  return expand ? (const void *)k_search<true, false, false, SAMPLE> : (const void *)k_search<false, false, false, SAMPLE>;
}

// the kernel instance a search runs on (sample: the parity-instrumented instances, never with learning)
static const void *search_kernel(const DevModel &m, bool expand, bool learn, bool sample, bool sat = false, bool backjump = false) {
  if (sat && !expand) return sample ? (const void *)k_search_sat<true> : (const void *)k_search_sat<false>;
#ifndef CSOLVE_BJ_UNIT
  if (learn && backjump && !expand) return csolve_bj_search_kernel();    // kernels_bj.cu
#endif
  if (learn) return expand ? (const void *)k_search<true, true> : (const void *)k_search<false, true>;
  if (m.lovk) return sample ? lovk_kernel<true>(expand, m.lovk) : lovk_kernel<false>(expand, m.lovk);
  if (m.lov) return sample ? lov_kernel<true>(expand, m.lov_bits != 0) : lov_kernel<false>(expand, m.lov_bits != 0);
  return sample ? general_kernel<true>(expand, m.n_lin > 0) : general_kernel<false>(expand, m.n_lin > 0);
}

static cudaError_t ensure_smem(const void *fn, size_t bytes) {
  if (bytes > 48 * 1024) return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return cudaSuccess;
}

bool search_learns(const SearchArgs &a) { return a.ng.lits != nullptr; }

int search_blocks_per_sm(const DevModel &m, bool expand, bool learn, bool sample, bool sat, bool backjump) {
  int n = 0;
  const size_t smem = search_smem_bytes(m, learn, sat && !expand);
  const void *fn = search_kernel(m, expand, learn, sample, sat, backjump);
  if (ensure_smem(fn, smem) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, THREADS_PER_BLOCK, smem) != cudaSuccess) return 0;
  return n;
}

cudaError_t launch_search(const SearchArgs &a, int grid, bool expand, cudaStream_t st, bool backjump) {
  const bool learn = search_learns(a);
  const bool sat = a.use_sat != 0 && !expand;
  const size_t smem = search_smem_bytes(a.m, learn, sat);
  const void *fn = search_kernel(a.m, expand, learn, a.sample_mod != 0u, sat, learn && backjump);
  cudaError_t e = ensure_smem(fn, smem);
  if (e != cudaSuccess) return e;
  void *args[] = {(void *)&a};
  return cudaLaunchKernel(fn, dim3(grid), dim3(THREADS_PER_BLOCK), args, smem, st);
}

cudaError_t launch_root_frames(const DevModel &m_in, int n_roots, const int32_t *root_dom, int order, int32_t *frames_out,
                               int out_cap, int32_t *n_out, unsigned char *root_failed, int grid, cudaStream_t st) {
  if (m_in.lovk) {
    switch (m_in.lovk) {
    case 2: k_root_frames_lovk<2><<<grid, THREADS_PER_BLOCK, 0, st>>>(m_in, n_roots, root_dom, order, frames_out, out_cap, n_out, root_failed); break;
    case 3: k_root_frames_lovk<3><<<grid, THREADS_PER_BLOCK, 0, st>>>(m_in, n_roots, root_dom, order, frames_out, out_cap, n_out, root_failed); break;
    default: k_root_frames_lovk<4><<<grid, THREADS_PER_BLOCK, 0, st>>>(m_in, n_roots, root_dom, order, frames_out, out_cap, n_out, root_failed); break;
    }
    return cudaGetLastError();
  }
  DevModel m = m_in;
  m.lov = 0;                                  // otherwise the batched path runs on the general kernels
  const size_t smem = search_smem_bytes(m, false);
  cudaError_t e = ensure_smem((const void *)k_root_frames, smem);
  if (e != cudaSuccess) return e;
  k_root_frames<<<grid, THREADS_PER_BLOCK, smem, st>>>(m, n_roots, root_dom, order, frames_out, out_cap, n_out, root_failed);
  return cudaGetLastError();
}

cudaError_t launch_comm_state(const SearchArgs &a, int want_frames, int force_idle, int32_t *out, cudaStream_t st) {
  k_comm_state<<<1, 1, 0, st>>>(a, want_frames, force_idle, out);
  return cudaGetLastError();
}

cudaError_t launch_comm_wait_front(const CommBlock *root, int epoch, double timeout_s, CommBlock *copy, cudaStream_t st) {
  k_comm_wait_front<<<1, 1, 0, st>>>(root, epoch, (unsigned long long)(timeout_s * 1e9), copy);
  return cudaGetLastError();
}

cudaError_t launch_rebalance(const SearchArgs &a, int32_t *scratch, cudaStream_t st) {
  k_rebalance<<<1, 1024, 0, st>>>(a, scratch);
  return cudaGetLastError();
}

cudaError_t launch_export_frames(const SearchArgs &a, int32_t *out, int max_frames, int32_t *n_out, cudaStream_t st) {
  k_export_frames<<<1, 1024, 0, st>>>(a, out, max_frames, n_out);
  return cudaGetLastError();
}

cudaError_t launch_import_frames(const SearchArgs &a, const int32_t *in, int n_frames, cudaStream_t st) {
  k_import_frames<<<1, 1024, 0, st>>>(a, in, n_frames);
  return cudaGetLastError();
}

cudaError_t launch_reduce_counters(const unsigned long long *wcount, int n_warps, unsigned long long *out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, CNT_WIDTH * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  const int grid = (int)std::min<size_t>(32, ((size_t)n_warps * CNT_WIDTH + REDUCE_THREADS - 1) / REDUCE_THREADS);
  k_reduce_counters<<<std::max(grid, 1), 1024, 0, st>>>(wcount, n_warps, out);
  return cudaGetLastError();
}

cudaError_t launch_propagate_batch(const DevModel &m, int n_nodes, const int32_t *dom_in, const int32_t *var,
                                   const int32_t *val, const int32_t *best, int32_t *dom_out, uint8_t *failed,
                                   int grid, cudaStream_t st) {
  const size_t smem = search_smem_bytes(m, false);
  if (m.lovk) {
    switch (m.lovk) {
    case 2: k_propagate_batch_lovk<2><<<grid, THREADS_PER_BLOCK, 0, st>>>(m, n_nodes, dom_in, var, val, dom_out, failed); break;
    case 3: k_propagate_batch_lovk<3><<<grid, THREADS_PER_BLOCK, 0, st>>>(m, n_nodes, dom_in, var, val, dom_out, failed); break;
    default: k_propagate_batch_lovk<4><<<grid, THREADS_PER_BLOCK, 0, st>>>(m, n_nodes, dom_in, var, val, dom_out, failed); break;
    }
    return cudaGetLastError();
  }
  if (m.lov) {
    cudaError_t e = ensure_smem((const void *)k_propagate_batch_lov, smem);
    if (e != cudaSuccess) return e;
    k_propagate_batch_lov<<<grid, THREADS_PER_BLOCK, smem, st>>>(m, n_nodes, dom_in, var, val, dom_out, failed);
    return cudaGetLastError();
  }
  cudaError_t e = ensure_smem((const void *)k_propagate_batch, smem);
  if (e != cudaSuccess) return e;
  k_propagate_batch<<<grid, THREADS_PER_BLOCK, smem, st>>>(m, n_nodes, dom_in, var, val, best, dom_out, failed);
  return cudaGetLastError();
}

}  // namespace csolve_dev
#endif  // CSOLVE_BJ_UNIT
