// device_model.h -- the GPU-resident form of a csolve_flat_model ("compiled"
// model) and the layout of search frames in HBM. Plain structs shared by the
// host compiler (compile.cpp), the contractors (contract.cuh) and the kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CSOLVE_HOSTDEV __host__ __device__
#else
#define CSOLVE_HOSTDEV
#endif

namespace csolve_dev {

// Compiled clause record: one 16-byte word per clause (a single LDG.128).
//   kind == CK_GENERIC : a = first node, b = root node (= last node), interpreted
//   kind == CK_NE_VV   : NOT(EQ(x + ka, y + kb))  ->  a = x, b = y, c = ka - kb   (x + c != y)
//   kind == CK_NE_VC   : NOT(EQ(x + ka, const))   ->  a = x, c = const - ka       (x != c)
//   kind == CK_LITS | n << 8 : a disjunction of n <= 3 literals over distinct 0/1 variables (a SAT clause in
//                        any OR / NOT(AND) nesting); a, b, c = literal codes var << 1 | negated
// The NE kinds are only emitted when no intermediate value can reach the
// saturation sentinels (see compile.cpp), so plain int32 arithmetic is bit-exact
// with the reference's saturating operators (src/arith.c).
enum ClauseKind : int32_t { CK_GENERIC = 0, CK_NE_VV = 1, CK_NE_VC = 2, CK_LITS = 3 };

struct ClauseRec { int32_t kind, a, b, c; };

// Watch record: what a lane executes when variable `self` is on the node's worklist. One 16-byte
// word; the records of a variable are contiguous (wrec_ptr[self] .. wrec_ptr[self + 1]).
//   w0 = kind << 30 | n << 28 | arg        (n = number of offsets/constants used, 1..3)
//   WK_GENERIC : n == 1: arg = clause index; the whole clause is contracted by the interpreter
//                n == 2: arg = index of a linear clause (LinClause below), contracted by the whole warp
//                n == 3: arg = index of a small linear relation (LinRel below), contracted by this lane, no interpreter
//   WK_NE_VV   : arg = partner variable p; for k < n:  self + w[1 + k] != p
//                (all NOT(EQ) clauses between the two variables, oriented from `self`, duplicates
//                 merged: contracting the same clause twice cannot change the fixpoint)
//   WK_NE_VC   : for k < n:  self != w[1 + k]
//   WK_LITS    : w[1 + k] = literal codes (var << 1 | negated) of a disjunction of n literals; unit propagation
enum WatchKind : uint32_t { WK_GENERIC = 0, WK_NE_VV = 1, WK_NE_VC = 2, WK_LITS = 3 };
struct WatchRec { uint32_t w0; int32_t c[3]; };
CSOLVE_HOSTDEV static inline uint32_t wrec_kind(uint32_t w0) { return w0 >> 30; }
CSOLVE_HOSTDEV static inline int wrec_n(uint32_t w0) { return (int)((w0 >> 28) & 3u); }
CSOLVE_HOSTDEV static inline int wrec_arg(uint32_t w0) { return (int)(w0 & 0x0fffffffu); }

// Linear clause  x_obj == konst + SUM_i (+-) k_i * x_i  (the objective clause of a MIN / MAX model with a linear
// objective; EQ(<obj>, expr) or EQ(expr, <obj>), src/parser.y): contracted by the whole warp, one term per lane,
// instead of one lane interpreting the whole tree. term.var: variable index | LIN_NEG (the term sits under a NEG
// node) | LIN_MUL (the term is a MUL(var, const) / MUL(const, var) node: propagate_mul's rules apply; without it the
// term is the bare variable, k == 1).
struct LinClause { int32_t obj, n_terms, first, konst, clause, pad0, pad1, pad2; };
struct LinTerm { int32_t var, k; };
// Linear relation between at most four variables with unit coefficients:  SUM(+-x_k) + konst  REL  0, where
// REL is == (EQ(l, r)), < (LT(l, r)) or >= (NOT(LT(l, r))) and the terms of r count negative. One lane contracts it in
// straight-line code (contract_linrel): a watch record WK_GENERIC with n == 3, arg = index into linrel[].
// v[k]: variable | LR_NEG (the term counts negative); unused slots are -1.
struct LinRel { int32_t rel, n, konst, clause; int32_t v[4]; };
static const int32_t LR_EQ = 0, LR_LT = 1, LR_GE = 2;
static const int32_t DF_CLAUSE = 0, DF_LINREL = 1, DF_LIN = 2;
static const int32_t LR_NEG = 1 << 30, LR_VAR = (1 << 28) - 1;
static const int32_t LIN_NEG = 1 << 30, LIN_MUL = 1 << 29, LIN_VAR = (1 << 28) - 1;
static const int MAX_LIN = 32;          // linear clauses per model (a dirty bit each), terms per clause (a lane each)

// Maximum expression depth the per-lane interpreter stacks are sized for.
// csolve_gpu_load() rejects deeper models (CSOLVE_ERR_UNSUPPORTED).
static const int MAX_DEPTH = 48;

struct int2_t { int32_t x, y; };

struct DevModel {
  int32_t n_vars, n_clauses, n_nodes, n_watch;
  int32_t objective, obj_var;
  int32_t mask_words;        // ceil(n_vars / 32)
  int32_t frame_words;       // 32-bit words per search frame (header + mask + domains), multiple of 4
  int32_t max_depth;         // deepest clause tree
  int32_t n_generic;         // number of CK_GENERIC clauses
  const ClauseRec *clause;   // [n_clauses]
  const int32_t *watch_ptr;  // [n_vars + 1]
  const int32_t *watch_idx;  // [n_watch]
  const WatchRec *wrec;      // [n_wrec] compiled watch records
  const int32_t *wrec_ptr;   // [n_vars + 1]
  int32_t n_wrec;
  int32_t table_smem_bytes;  // bytes needed to stage wrec + wrec_ptr in shared memory (0 = too large)
  // "lane owns variable" form (pure NOT(EQ) networks with at most 32 variables, e.g. N-queens):
  //   lov_pair[i * 32 + j] = 64-bit set of forbidden differences: bit (c + 32) is set iff the network
  //                           holds the clause x_i + c != x_j  (-32 <= c < 32)
  //   lov_cptr / lov_cval   = per variable CSR of forbidden constants (x_i != c)
  int32_t lov;               // 1 when the model is eligible
  int32_t lov_bits;          // 1 when additionally all root domains fit a 32-value window (forbidden-value sets)
  int32_t lov_vbase;         // smallest root lower bound (bit 0 of the value sets)
  int32_t lovk;              // K = 2..4 when the model (33..128 variables, value window <= 32) runs on the K-variables-per-lane
                             // kernel; its frames carry n_vars extra words (the forbidden-value sets) after the domains
  int32_t lov_smem_bytes;
  int32_t n_lov_cval;
  const unsigned long long *lov_pair;  // [n_vars * 32]
  const int32_t *lov_cptr;   // [n_vars + 1]
  const int32_t *lov_cval;   // [n_lov_cval]
  const uint32_t *lov_fconst; // [n_vars] forbidden-value set of the constants (lov_bits)
  // K-per-lane form of an all-different style network (every pair clause is x_i != x_j, offset 0: sudoku): the pair
  // table is an adjacency bit matrix, lov_adj[i * K + q] bit l = variable l + 32 q shares a clause with variable i
  // (972 bytes for sudoku instead of 62 KB of 64-bit offset sets). lov_adj_only: 1 when the model qualifies.
  const uint32_t *lov_adj;    // [n_vars * K]
  int32_t lov_adj_only;
  // "bit state" form (pure SAT: every variable's root domain lies in [0,1], every clause is a disjunction of at most
  // three literals): sat_occ_ptr[e] .. sat_occ_ptr[e + 1], e = var << 1 | value, are the clauses in which the
  // assignment var := value falsifies a literal; each record holds the clause's OTHER literals (code = var << 1 |
  // negated, -1 = none)
  int32_t sat;               // 1 when the model is eligible
  int32_t n_sat_occ;
  int32_t sat_smem_bytes;    // bytes needed to stage the occurrence table, the order and the root state (0 = table stays in global memory)
  const int32_t *sat_occ_ptr;   // [2 * n_vars + 1]
  const int2_t *sat_occ;        // [n_sat_occ]
  int32_t n_linrel;          // small linear relations (watch records WK_GENERIC with n == 3)
  const LinRel *linrel;      // [n_linrel]
  // "dense" propagation (at most 32 clauses): a fixpoint round contracts EVERY clause, lane c clause c, instead of
  // walking worklists -- dense_form[c] = {form, arg}: DF_CLAUSE (ClauseRec c), DF_LINREL (linrel[arg]), DF_LIN (lin[arg],
  // contracted by the whole warp after the lanes' clauses)
  int32_t dense;
  const int2_t *dense_form;  // [n_clauses]
  int32_t n_lin_term;        // terms of all linear clauses
  int32_t n_lin;             // linear clauses (watch records WK_GENERIC with n == 2, arg = index into lin[])
  const LinClause *lin;      // [n_lin]
  const LinTerm *lin_term;   // terms of all linear clauses
  const uint8_t *node_op;    // [n_nodes]
  const int32_t *node_l;     // [n_nodes]
  const int32_t *node_r;     // [n_nodes]
  const int32_t *node_first; // [n_nodes] first node of the subtree rooted at a node (post-order => contiguous)
  const int32_t *order;      // [n_vars] static branching order: priority descending, index ascending
  const int32_t *prio;       // [n_vars] parse-time priority clamped to int32 (tie-break of the dynamic orders)
  const int32_t *root_dom;   // [2 * n_vars] lo,hi pairs
};

// ---- search frame (one per DFS level per warp, and one per frontier item) -----
// word 0  var        branching variable of this level
// word 1  iter       next iteration index to try (src/csolve.c:331-338 value order)
// word 2  iter_last  last iteration index this frame owns (inclusive)
// word 3  lo         bounds captured when the level was activated (src/csolve.c:282)
// word 4  hi
// word 5  level      number of variables assigned before this frame
// word 6  best_seen  incumbent the domains were propagated against (MIN/MAX)
// word 7  reserved
// words 8 .. 8+mask_words-1                 bitmask of variables assigned before this frame
// then (16-byte aligned) 2*n_vars words     domains BEFORE the assignment of this level (lo,hi pairs)
static const int FR_VAR = 0, FR_ITER = 1, FR_LAST = 2, FR_LO = 3, FR_HI = 4, FR_LEVEL = 5, FR_BEST = 6;
static const int FR_MASK = 8;

CSOLVE_HOSTDEV static inline int frame_dom_offset(int mask_words) { return (FR_MASK + mask_words + 3) & ~3; }
CSOLVE_HOSTDEV static inline int frame_words(int n_vars, int mask_words) {
  return (frame_dom_offset(mask_words) + 2 * n_vars + 3) & ~3;
}

}  // namespace csolve_dev
