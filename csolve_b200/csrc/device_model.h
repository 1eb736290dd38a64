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
// The NE kinds are only emitted when no intermediate value can reach the
// saturation sentinels (see compile.cpp), so plain int32 arithmetic is bit-exact
// with the reference's saturating operators (src/arith.c).
enum ClauseKind : int32_t { CK_GENERIC = 0, CK_NE_VV = 1, CK_NE_VC = 2 };

struct ClauseRec { int32_t kind, a, b, c; };

// Maximum expression depth the per-lane interpreter stacks are sized for.
// csolve_gpu_load() rejects deeper models (CSOLVE_ERR_UNSUPPORTED).
static const int MAX_DEPTH = 48;

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
