// kernels.cuh -- declarations shared by kernels.cu (device code) and capi.cu (host
// orchestration): the control block, per-warp state and launch wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "device_model.h"

namespace csolve_dev {

static const int WARPS_PER_BLOCK = 8;
static const int THREADS_PER_BLOCK = WARPS_PER_BLOCK * 32;

// signal values polled by every warp once per search node
static const int SIG_RUN = 0;          // keep going
static const int SIG_SLICE_END = 1;    // park at the next node boundary (rebalance / incumbent exchange / time slice)
static const int SIG_STOP = 2;         // ANY: a solution was found

// per-warp counters (uint64 each), accumulated over all slices
// CNT_WAIT / CNT_CLAIMS / CNT_LASTWORK: load-balance diagnostics (cycles spent waiting for a frame, frames claimed,
// cycle offset of the warp's last node inside its last slice), printed under CSOLVE_DEBUG
enum { CNT_NODES = 0, CNT_CUTS, CNT_PROPS, CNT_VISITS, CNT_SOLUTIONS, CNT_REFRESH, CNT_WAIT, CNT_CLAIMS, CNT_LASTWORK, CNT_POLLS, CNT_WANTED, CNT_DONATED, CNT_WIDTH };

// The fields every warp hammers with atomics sit on cache lines of their own (an L2 atomic locks its line: with
// everything on one line the claims of the root frontier, the ring head/tail and the hungry counter serialised each
// other and the signal polls -- 8..16 us per claim were measured on a busy B200).
struct SearchCtl {
  // line 0: read-mostly (polled)
  int32_t signal;
  int32_t best;         // incumbent objective value (objective_best(), src/objective.c:133)
  int32_t busy;         // written by the rebalance kernel: warps that still own work
  int32_t moved;        // written by the rebalance kernel: frames handed to idle warps
  int32_t pad0[28];
  // line 1: claims of the expanded root frontier
  int32_t init_next;    // next frame of the expanded root frontier (static part of the pool, claimed with atomicAdd)
  int32_t pad1[31];
  // line 2: ring of donated frames (depth-first phase) / input cursor (expansion)
  int32_t item_next;    // expansion: next frontier item to hand out; depth-first phase: tickets taken by waiting warps
  int32_t pad2[31];
  int32_t item_count;   // expansion: frontier items available; depth-first phase: tickets served by donating warps
  int32_t pad2b[31];
  // line 3
  int32_t hungry;       // warps waiting for a frame of the shared pool
  int32_t pad3[31];
  // line 4: output side
  int32_t n_stored;     // assignments written to the solution buffer
  int32_t out_count;    // expand mode: frames appended to the output frontier
  int32_t out_dropped;  // expand mode: children that did not fit (capacity error)
  int32_t passed;       // expand mode: frames passed through unsplit (domain too large to enumerate)
  int32_t fails;        // failed nodes reported by the warps since the last restart (only counted when restarts are on)
  int32_t pad4[27];
};

// One process (or host thread) per GPU: what the ranks of a csolve_gpu_comm share. Every rank owns one CommBlock in
// device memory that its peers map (NVLink peer access inside one process, CUDA IPC between processes) and write with
// system-scope atomics; nothing in it is ever reset -- every entry carries the epoch (the comm's solve counter) it
// belongs to, so a late store of the previous search cannot leak into the next one.
//   rmin64 / rmax64  best incumbent pushed by a peer (MIN / MAX models): epoch-tagged key, see comm_key_min / _max
//   stop_epoch       ANY models: a peer found a solution in this epoch (found_any(), src/csolve.c:207-209)
//   front_*          rank 0 only: the expanded root frontier of this epoch is in rank 0's segment (front_n frames of
//                    front_fw words; front_n < 0: not shared -- every rank expands and partitions by path hash)
//   done_epoch[r]    rank 0 only: rank r has finished the search of this epoch (the frontier may be overwritten)
//   busy_epoch       == the current epoch while the rank has (or is being handed) work; cleared by the rank itself when
//                    it ran dry, set again by whoever hands it a frame (k_comm_state / publish_slot)
//   demand[r]        frames rank r asks THIS rank for (r is running dry: raised from its waiting loop, claim_frame, or by
//                    its host loop, k_comm_state); a busy warp that
//                    sees a positive entry takes one unit and serves a ticket of rank r's donation ring over NVLink
//   ring_open/inflight  a peer serves this rank's ring while ring_open == epoch: while the rank's search kernel runs and
//                    while the rank waits in its idle loop; the rank closes the ring and waits for the peers in flight
//                    before its counters are rebased (k_rebalance) or judged (k_comm_state)
//   active64         rank 0 only: epoch << 32 | number of ranks whose busy_epoch is set -- 0 ends the search everywhere
struct CommBlock {
  unsigned long long rmin64, rmax64;
  int32_t stop_epoch;
  int32_t front_epoch, front_n, front_fw;
  int32_t done_epoch[8];
  int32_t busy_epoch;
  int32_t ring_open;          // == epoch while the rank sits in its idle loop: only then may peers serve its ring
  unsigned long long active64;
  int32_t demand[8];
  int32_t inflight;           // peers that are serving a ticket of this rank's ring right now
  int32_t pad[3];
};
static const int COMM_MAX_RANKS = 8;
// epoch-tagged incumbent keys: a later epoch always wins the atomic, inside an epoch the better value wins
CSOLVE_HOSTDEV static inline unsigned long long comm_key_min(int epoch, int32_t v) {
  return ((unsigned long long)(0xffffffffu - (uint32_t)epoch) << 32) | (uint32_t)((uint32_t)v ^ 0x80000000u);
}
CSOLVE_HOSTDEV static inline unsigned long long comm_key_max(int epoch, int32_t v) {
  return ((unsigned long long)(uint32_t)epoch << 32) | (uint32_t)((uint32_t)v ^ 0x80000000u);
}
CSOLVE_HOSTDEV static inline int32_t comm_key_value(unsigned long long k) { return (int32_t)((uint32_t)k ^ 0x80000000u); }
CSOLVE_HOSTDEV static inline bool comm_key_is_epoch_min(unsigned long long k, int epoch) { return (uint32_t)(k >> 32) == 0xffffffffu - (uint32_t)epoch; }
CSOLVE_HOSTDEV static inline bool comm_key_is_epoch_max(unsigned long long k, int epoch) { return (uint32_t)(k >> 32) == (uint32_t)epoch; }

struct WarpState {
  int32_t level;   // index of the top frame of the warp's stack; < base when the warp is idle
  int32_t base;    // lowest level the warp owns
  int32_t claim_base;   // frames of the expanded root frontier this warp has claimed and not searched yet:
  uint32_t claim_mask;  // bit b = frame claim_base + b (kept across time slices)
};

// Learned nogoods (src/conflict.c): an append-only pool shared by all warps of one GPU.
//   nogood k = lits[start[k] .. start[k] + len[k]), literal code = var << 1 | value (value in {0,1})
//   watch[v * cap_w + i] = nogoods that mention variable v (-1 = slot reserved, not written yet)
//   counters: [0] nogoods, [1] literals, [2] conflicts analysed, [3] analyses abandoned (a non-0/1 value was involved,
//   src/conflict.c:173-179, or the nogood was too long), [4] pool full, [5] back-jumps taken
struct NogoodPool {
  int32_t *lits, *start, *len, *watch, *watch_n, *counters;
  int32_t cap_ng, cap_lits, cap_w;
};
static const int NG_MAX_LITS = 96;   // longest nogood kept

struct SearchArgs {
  DevModel m;
  SearchCtl *ctl;
  int32_t *stacks;            // [n_warps][n_vars + 1][frame_words]
  WarpState *wstate;          // [n_warps]
  unsigned long long *wcount; // [n_warps][CNT_WIDTH]
  const int32_t *items;       // frontier pool: [item_count][frame_words]
  int32_t *items_out;         // expand mode: output pool
  int32_t out_cap;            // expand mode: capacity of the output pool (frames)
  int32_t *solbuf;            // [max_solutions][n_vars + 1]  (values..., objective key)
  int32_t max_solutions;
  int32_t fail_limit;         // > 0: restarts are on (ANY models, src/csolve.c:264-276): the slice ends when the warps have
                              // reported more failed nodes than this
  int32_t use_sat;            // the depth-first phase runs on k_search_sat (pure SAT model, static or failure-driven order)
  int32_t sink_headroom;      // > 0: the host drains the solution buffer between slices (csolve_gpu_set_solution_sink): a
                              // slice ends as soon as fewer than this many entries are free
  int32_t n_warps;
  int32_t order;              // CSOLVE_ORDER_*
  int32_t frozen_best;        // expand mode: incumbent every node of this level is propagated against
  long long slice_cycles;     // clock64() budget of one slice
  int32_t expand_branch_max;  // expand mode: frames with more values than this are passed through unsplit
  NogoodPool ng;              // ng.lits == nullptr: conflict-clause learning off
  int32_t *gprio;             // prefer-failing: device-wide dynamic priorities [n_vars] (else nullptr)
  unsigned int *inst_solutions; // batched roots: per-root solution counters (else nullptr); the root id travels in header word 6
  int32_t *ready;             // [pool_cap] 1 = the pool slot holds a complete frame (shared pool ring)
  int32_t *pool;              // the shared pool (same memory as `items` in the depth-first phase), writable
  int32_t pool_cap;           // frames in the pool ring
  int32_t n_initial;          // the first n_initial pool entries are the expanded root frontier (rank partition applies)
  int32_t part_rank;          // this process searches the frontier frames whose path hash % part_count == part_rank
  int32_t part_count;
  // Ranks of a csolve_gpu_comm (n_peers == 0: a search on its own). ANY / MIN / MAX models: the expanded root frontier
  // and its claim counter live on rank 0's GPU and every rank claims frames of that ONE frontier with a system-scope
  // atomicAdd over NVLink. (ALL models are dealt by path hash and searched without a comm, capi.cu.)
  const int32_t *front_pool;  // the root frontier: [n_initial][frame_words] (== pool without a comm)
  SearchCtl *front_ctl;       // control block whose init_next hands the frontier out (== ctl without a comm)
  CommBlock *comm;            // this rank's block (peers write it)
  CommBlock *peer_comm[COMM_MAX_RANKS];   // every rank's block, peer-mapped, indexed by rank (peer_comm[rank] == comm)
  // every rank's donation ring (control block, frames, ready flags; same slot numbering on all ranks: the frontier
  // and the ring size are the same everywhere): a busy warp serves the tickets of a rank that ran dry over NVLink
  SearchCtl *peer_ctl[COMM_MAX_RANKS];
  int32_t *peer_pool[COMM_MAX_RANKS];
  int32_t *peer_ready[COMM_MAX_RANKS];
  int32_t n_peers;            // world - 1 (0: no comm)
  int32_t rank, world;
  int32_t epoch;
  int32_t total_warps;        // search warps of all ranks (size of a guided chunk)
  int32_t peer_demand;        // ranks of a comm: a starving rank asks its peers for frames while its kernel runs (claim_frame)
  // Parity instrumentation (csolve_solve_options.sample_mod > 0; runs the SAMPLE instances of the search kernels):
  // every search node -- executed or counted in bulk -- whose identity hash (parent domains, variable, value) is 0
  // modulo sample_mod is written to sample_rec as
  //   [0] flags (SAMPLE_*)  [1] variable  [2] value  [3] incumbent the node was propagated against
  //   [4 .. 4+2V) parent domains (state before the assignment)   [4+2V .. 4+4V) post-fixpoint domains
  int32_t *sample_rec;        // [sample_cap][4 + 4 * n_vars]
  int32_t *sample_n;          // records wanted so far (may exceed sample_cap: the excess was dropped)
  int32_t sample_cap;
  uint32_t sample_mod;
  uint32_t sample_fkeep;      // of the hits that failed, 1 in sample_fkeep is kept (>= 1)
};
static const int SAMPLE_FAILED = 1;    // the node failed (PROP_ERROR); its post-fixpoint domains are meaningless
static const int SAMPLE_COUNTED = 2;   // the node was counted by a bulk shortcut of the kernel, not executed
static const int SAMPLE_LEAF = 4;      // the node is an accepted leaf (every variable a value, every clause true)
CSOLVE_HOSTDEV static inline int sample_words(int n_vars) { return 4 + 4 * n_vars; }

size_t search_smem_bytes(const DevModel &m, bool learn = false, bool sat = false);
bool search_uses_sat(const DevModel &m, bool learn, int order);
cudaError_t launch_search(const SearchArgs &a, int grid, bool expand, cudaStream_t s, bool backjump = false);
bool search_learns(const SearchArgs &a);
cudaError_t launch_rebalance(const SearchArgs &a, int32_t *scratch, cudaStream_t s);
// comm: this rank's state between two kernel launches. out[0] = 1 if the rank has work (frames in its ring or in the
// shared frontier), out[1] = number of ranks still active (-1: another epoch), out[2] = 1 if a peer found a solution
// (ANY). want_frames > 0 and no work: ask every peer for that many frames. force_idle: leave the epoch whatever is left.
cudaError_t launch_comm_state(const SearchArgs &a, int want_frames, int force_idle, int32_t *out, cudaStream_t s);
// comm: wait on the device until rank 0's block carries front_epoch >= epoch (or timeout_s passed), then copy the block
// to `copy` (device memory of this rank)
cudaError_t launch_comm_wait_front(const CommBlock *root, int epoch, double timeout_s, CommBlock *copy, cudaStream_t s);
cudaError_t launch_export_frames(const SearchArgs &a, int32_t *out, int max_frames, int32_t *n_out, cudaStream_t s);
cudaError_t launch_import_frames(const SearchArgs &a, const int32_t *in, int n_frames, cudaStream_t s);
cudaError_t launch_reduce_counters(const unsigned long long *wcount, int n_warps, unsigned long long *out, cudaStream_t s);
cudaError_t launch_propagate_batch(const DevModel &m, int n_nodes, const int32_t *dom_in, const int32_t *var,
                                   const int32_t *val, const int32_t *best, int32_t *dom_out, uint8_t *failed,
                                   int grid, cudaStream_t s);
int search_blocks_per_sm(const DevModel &m, bool expand, bool learn = false, bool sample = false, bool sat = false, bool backjump = false);
cudaError_t launch_root_frames(const DevModel &m, int n_roots, const int32_t *root_dom, int order, int32_t *frames_out,
                               int out_cap, int32_t *n_out, unsigned char *root_failed, int grid, cudaStream_t s);

}  // namespace csolve_dev

// the back-jumping instance of the general search kernel (csolve_solve_options.backjump), compiled in a unit of its own
extern "C" const void *csolve_bj_search_kernel(void);
