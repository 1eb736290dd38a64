// emu_search.cpp -- TEST-ONLY: the product's search kernels (csolve_b200/csrc/kernels.cu: k_search, k_search_lov,
// k_search_lovk, k_search_sat, k_rebalance) run on the CPU under the SIMT emulator of simt_emu.h, driven the way capi.cu
// drives the depth-first phase of a search whose root was not expanded (split_target = 1): one root frame in the pool,
// the donation ring behind it, time slices with k_rebalance between them until no warp owns work.
// Built by tests/util.py build_emu() from the kernel source itself (kernels_emu.inc = kernels.cu with its __shared__
// declarations rewritten and the launch wrappers cut off); -DCSOLVE_BJ builds k_search<false, true> as kernels_bj.cu does.
// Not part of the library, never a fallback.
#define EMU_DEFINE_MACHINE 1
#include "simt_emu.h"
extern "C" const void *csolve_bj_search_kernel(void) { return nullptr; }     // kernels_bj.cu's instance: here -DCSOLVE_BJ
#include "kernels_emu.inc"
#include <algorithm>
#include <climits>
#include <string>
#include "compile.hpp"

using namespace csolve_dev;

namespace {

int root_var(const CompiledModel &cm, int order) {      // select_root_var of capi.cu
  const DevModel &m = cm.host;
  if (order == CSOLVE_ORDER_NONE) return cm.order[0];
  unsigned long long bestk = ~0ull;
  int bestv = 0;
  for (int v = 0; v < m.n_vars; v++) {
    const int lo = cm.root_dom[2 * v], hi = cm.root_dom[2 * v + 1];
    unsigned primary;
    switch (order) {
    case CSOLVE_ORDER_SMALLEST_DOMAIN: primary = (unsigned)hi - (unsigned)lo; break;
    case CSOLVE_ORDER_LARGEST_DOMAIN:  primary = ~((unsigned)hi - (unsigned)lo); break;
    case CSOLVE_ORDER_SMALLEST_VALUE:  primary = (unsigned)lo ^ 0x80000000u; break;
    default:                           primary = ~((unsigned)hi ^ 0x80000000u); break;
    }
    const unsigned secondary = ~((unsigned)cm.prio[v] ^ 0x80000000u);
    const unsigned long long k = ((unsigned long long)primary << 32) | secondary;
    if (k < bestk) { bestk = k; bestv = v; }
  }
  return bestv;
}

struct Launch { SearchArgs a; void (*fn)(const SearchArgs); int32_t *scratch; };

void run_kernel(void *arg) {
  const Launch *l = static_cast<const Launch *>(arg);
  l->fn(l->a);
}
struct BatchLaunch { DevModel m; int n_roots; const int32_t *root_dom; int order; int32_t *frames; int cap; int32_t *n_out; unsigned char *failed; };
void run_root_frames(void *arg) {
  const BatchLaunch *b = static_cast<const BatchLaunch *>(arg);
  switch (b->m.lovk) {
  case 0: k_root_frames(b->m, b->n_roots, b->root_dom, b->order, b->frames, b->cap, b->n_out, b->failed); break;
  case 2: k_root_frames_lovk<2>(b->m, b->n_roots, b->root_dom, b->order, b->frames, b->cap, b->n_out, b->failed); break;
  case 3: k_root_frames_lovk<3>(b->m, b->n_roots, b->root_dom, b->order, b->frames, b->cap, b->n_out, b->failed); break;
  default: k_root_frames_lovk<4>(b->m, b->n_roots, b->root_dom, b->order, b->frames, b->cap, b->n_out, b->failed); break;
  }
}
struct ReduceLaunch { const unsigned long long *wcount; int n_warps; unsigned long long *out; };
void run_reduce(void *arg) {
  const ReduceLaunch *r = static_cast<const ReduceLaunch *>(arg);
  k_reduce_counters(r->wcount, r->n_warps, r->out);
}
void run_rebalance(void *arg) {
  const Launch *l = static_cast<const Launch *>(arg);
  k_rebalance(l->a, l->scratch);
}

std::string g_err;
std::vector<int32_t> g_samples;      // records of the last search with sample_mod > 0 (sample_words(V) words each)
int g_sample_seen = 0;
unsigned g_sample_mod = 0, g_sample_fkeep = 1;
int32_t g_fill = 0;                  // what the workspace holds before a search (cudaMalloc'ed memory is not zeroed either)
std::vector<int32_t> g_ng_flat;      // nogoods of the last learning search: len, literals (var << 1 | value), len, ...

}  // namespace

struct emu_result {
  uint64_t solutions, nodes, cuts, props;
  int32_t best, has_solution, n_stored, conflicts, conflicts_abandoned, backjumps, claims, slices;
  int32_t expand_levels, frontier, restarts, pad;
  uint64_t switches, collectives, site_mismatches;
};

extern "C" const char *emu_error() { return g_err.c_str(); }
// lane schedule of the emulator (simt_emu.h): an odd stride, 1 = ascending lanes, 31 = descending
// the pattern stacks / pools / solution buffers are filled with before a search (default 0)
extern "C" void emu_set_workspace_fill(int32_t v) { g_fill = v; }
extern "C" void emu_set_lane_step(int step) { emu::lane_step = (step & 31) | 1; }
// warp schedule: collectives a warp completes before the other warps get their turn (default 64)
extern "C" void emu_set_warp_quantum(int q) { emu::warp_quantum = q > 0 ? (unsigned)q : 64u; }
// parity instrumentation (csolve_solve_options.sample_mod): the next searches run on the SAMPLE instances of the kernels
extern "C" void emu_set_sampling(unsigned mod, unsigned failed_keep) { g_sample_mod = mod; g_sample_fkeep = failed_keep; }
// records of the last sampled search: returns the number kept, *seen = hits including the ones that did not fit
extern "C" int emu_samples(int32_t *out, int cap_words, int *seen) {
  if (seen != nullptr) *seen = g_sample_seen;
  if (out != nullptr) memcpy(out, g_samples.data(), sizeof(int32_t) * std::min((size_t)cap_words, g_samples.size()));
  return (int)g_samples.size();
}
// nogoods learned by the last search, flattened (length, literals, length, ...); returns the number of words
extern "C" int emu_nogoods(int32_t *out, int cap) {
  const int n = (int)g_ng_flat.size();
  if (out != nullptr) memcpy(out, g_ng_flat.data(), sizeof(int32_t) * (size_t)std::min(n, cap));
  return n;
}
extern "C" int emu_backjump_build() {
#ifdef CSOLVE_BJ
  return 1;
#else
  return 0;
#endif
}

// one whole search; solutions: [max_solutions][n_vars + 1] (values..., objective key) or NULL
// general: 1 = the general kernel whatever the model (what capi.cu does for batched roots), 0 = the kernel the product
// picks (lane-owns-variable / K-per-lane / bit-state / general). slice_clock: length of a time slice in emulator clock
// units (every clock64() call adds 64), 0 = the whole search in one slice.
// split_target > 1: the root is expanded breadth-first first (k_search<true>, level by level as capi.cu's expand_root)
// until at least that many frames exist; part_rank / part_count: this call searches the frames whose path hash maps to
// part_rank (the ALL-mode partition between GPUs; the replicated expansion is reported by rank 0 only).
// sink_headroom > 0 (ALL models): the solution buffer holds 4 x sink_headroom assignments and is drained between slices
// into `solutions` (room for sink_rows assignments), as capi.cu does for csolve_gpu_set_solution_sink.
// n_roots > 0: batched roots (csolve_gpu_solve_batch): the root phase runs on the device (k_root_frames / _lovk: every
// root propagated to fixpoint, one tagged frame per consistent root), per-root solution counts and failed flags come back.
static int search_core(const csolve_flat_model *fm, int order, int learn, int prefer_failing, int n_blocks,
                       int max_solutions, int general, long long slice_clock, int sink_headroom, int sink_rows,
                       int split_target, int part_rank, int part_count, int restart_frequency, int n_roots, const int32_t *root_dom,
                       uint32_t *root_solutions, uint8_t *root_failed, emu_result *res, int32_t *solutions) {
  CompiledModel cm;
  int rc = compile_model(*fm, cm, g_err);
  if (rc != 0) return rc;
  DevModel m = cm.host;
  const bool batch = n_roots > 0;
  if (batch && m.objective != CSOLVE_OBJ_ALL) { g_err = "batched roots need an ALL model"; return -105; }
  if (batch) m.lov = 0;                  // capi.cu: batched roots run on the general kernels (or the K-per-lane one)
  if (general || learn) { m.lov = 0; m.lovk = 0; }
  if (prefer_failing && m.lov) prefer_failing = 0;       // capi.cu: the lane-owns-variable kernel has no dynamic priorities
  const int V = m.n_vars, fw = m.frame_words;
  const int n_warps = n_blocks * WARPS_PER_BLOCK;
  memset(res, 0, sizeof(*res));

  SearchCtl ctl;
  memset(&ctl, 0, sizeof(ctl));
  ctl.best = m.objective == CSOLVE_OBJ_MIN ? INT32_MAX : (m.objective == CSOLVE_OBJ_MAX ? INT32_MIN : 0);
  ctl.signal = SIG_RUN;
  const int ring = 4 * n_warps + 1024;   // ring_min_frames of capi.cu
  const int target = split_target > 1 ? split_target : 1;
  const int pool_cap = std::max(std::max(4 * target, 1024), 2 * n_roots) + ring;
  // stacks, pools and the solution buffer start with whatever the workspace held (g_fill); capi.cu clears only the ready
  // flags, the per-warp state and the counters
  std::vector<int32_t> pool((size_t)pool_cap * fw, g_fill), pool_b((size_t)pool_cap * fw, g_fill), ready(pool_cap, 0), stacks((size_t)n_warps * (V + 1) * fw, g_fill);
  std::vector<WarpState> ws(n_warps, WarpState{-1, 0, 0, 0u});
  std::vector<unsigned long long> wcount((size_t)n_warps * CNT_WIDTH, 0);
  const bool sinking = sink_headroom > 0 && m.objective == CSOLVE_OBJ_ALL;
  const int sol_cap = sinking ? 4 * sink_headroom : max_solutions > 0 ? max_solutions : (m.obj_var >= 0 ? 16 : 1);
  long long sunk = 0;
  std::vector<int32_t> solbuf((size_t)sol_cap * (V + 1), g_fill);
  auto upload_root_frame = [&](int rv) {
  int32_t *root = pool.data();
  std::fill(root, root + fw, 0);
  root[FR_VAR] = rv; root[FR_ITER] = 0;
  root[FR_LO] = cm.root_dom[2 * rv]; root[FR_HI] = cm.root_dom[2 * rv + 1];
  root[FR_LAST] = (int32_t)((uint32_t)root[FR_HI] - (uint32_t)root[FR_LO]);
  root[FR_LEVEL] = 0; root[FR_BEST] = ctl.best; root[7] = 0x1234567;
  memcpy(&root[frame_dom_offset(m.mask_words)], cm.root_dom.data(), sizeof(int32_t) * 2 * V);
  if (m.lovk) memcpy(&root[frame_dom_offset(m.mask_words) + 2 * V], cm.lov_fconst.data(), sizeof(int32_t) * V);   // value sets
  };
  upload_root_frame(root_var(cm, order));

  std::vector<int32_t> gprio(cm.prio.begin(), cm.prio.end());
  NogoodPool ng;
  memset(&ng, 0, sizeof(ng));
  std::vector<int32_t> ng_lits, ng_start, ng_len, ng_watch, ng_watch_n, ng_counters(8, 0);
  if (learn) {
    ng.cap_ng = 1 << 14; ng.cap_lits = 1 << 18; ng.cap_w = 1024;
    ng_lits.assign(ng.cap_lits, 0); ng_start.assign(ng.cap_ng, 0); ng_len.assign(ng.cap_ng, 0);
    ng_watch.assign((size_t)V * ng.cap_w, -1); ng_watch_n.assign(V, 0);
    ng.lits = ng_lits.data(); ng.start = ng_start.data(); ng.len = ng_len.data();
    ng.watch = ng_watch.data(); ng.watch_n = ng_watch_n.data(); ng.counters = ng_counters.data();
  }

  Launch l;
  SearchArgs &a = l.a;
  memset(&a, 0, sizeof(a));
  a.m = m; a.ctl = &ctl; a.stacks = stacks.data(); a.wstate = ws.data(); a.wcount = wcount.data();
  a.solbuf = solbuf.data(); a.max_solutions = sol_cap; a.n_warps = n_warps; a.order = order;
  a.out_cap = pool_cap; a.expand_branch_max = 64;
  a.part_rank = 0; a.part_count = 1;
  a.slice_cycles = slice_clock > 0 ? slice_clock : LLONG_MAX / 2;
  if (sinking) a.sink_headroom = sink_headroom;
  std::vector<int32_t> scratch(4 + 3 * n_warps, 0);
  l.scratch = scratch.data();
  const bool sample = g_sample_mod != 0u && !learn;
  const int sample_cap = 1 << 16;
  int32_t sample_n = 0;
  g_samples.clear(); g_sample_seen = 0;
  if (sample) {
    g_samples.assign((size_t)sample_cap * sample_words(V), 0);
    a.sample_rec = g_samples.data(); a.sample_n = &sample_n; a.sample_cap = sample_cap; a.sample_mod = g_sample_mod;
    a.sample_fkeep = g_sample_fkeep > 0 ? g_sample_fkeep : 1;
  }

  // ---- batched frontier expansion (capi.cu: expand_root) ----------------------------------------------------------
  int32_t *pin = pool.data(), *pout = pool_b.data();
  int n_items = 1;
  bool stopped = false;
  std::vector<unsigned int> rsol(std::max(n_roots, 1), 0u);
  if (batch) {
    // root phase on the device (capi.cu: launch_root_frames)
    static BatchLaunch bl;
    std::vector<unsigned char> rfail(n_roots, 0);
    int32_t n_out = 0;
    bl = BatchLaunch{m, n_roots, root_dom, order, pool.data(), pool_cap, &n_out, rfail.data()};
    const int grid = std::min(n_blocks, (n_roots + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    emu::launch(grid, THREADS_PER_BLOCK, m.lovk ? 0 : search_smem_bytes(m, false), run_root_frames, &bl);
    memcpy(root_failed, rfail.data(), n_roots);
    n_items = n_out;
    a.inst_solutions = rsol.data();
  }
  int rc_expand = 0;
  auto expand_root = [&]() {
    long long max_branch = 1;
    for (int v = 0; v < V; v++) max_branch = std::max<long long>(max_branch, (long long)cm.root_dom[2 * v + 1] - cm.root_dom[2 * v] + 1);
    max_branch = std::min<long long>(max_branch, a.expand_branch_max);
    l.fn = reinterpret_cast<void (*)(const SearchArgs)>(const_cast<void *>(search_kernel(m, true, false, sample, false, false)));
    const size_t smem_x = search_smem_bytes(m, false, false);
    for (int lvl = 0; lvl < V && lvl < 24 && n_items > 0 && n_items < target; ++lvl) {
      const int before = n_items;
      if ((long long)n_items * max_branch > pool_cap - ring) break;
      a.items = pin; a.items_out = pout; a.frozen_best = ctl.best;
      ctl.item_next = 0; ctl.item_count = n_items; ctl.out_count = 0; ctl.passed = 0;
      const int grid = std::min(n_blocks, (n_items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
      emu::launch(grid, THREADS_PER_BLOCK, smem_x, run_kernel, &l);
      res->collectives += emu::M.collectives;
      if (ctl.out_dropped > 0) { g_err = "frontier pool overflow during expansion"; rc_expand = -104; return; }
      if (sinking && ctl.n_stored > 0) {
        if (sunk + ctl.n_stored > sink_rows) { g_err = "more solutions than the caller expects"; rc_expand = -103; return; }
        if (!(part_count > 1 && part_rank != 0)) { memcpy(solutions + (size_t)sunk * (V + 1), solbuf.data(), sizeof(int32_t) * (size_t)ctl.n_stored * (V + 1)); sunk += ctl.n_stored; }
        ctl.n_stored = 0;
      }
      n_items = ctl.out_count;
      std::swap(pin, pout);
      res->expand_levels++;
      if (ctl.signal == SIG_STOP) { stopped = true; break; }
      if (ctl.passed == n_items) break;
      if (n_items >= n_warps / 2 && n_items < 2 * (long long)before) break;
    }
  };
  if (!(batch && n_items >= n_warps / 2)) expand_root();
  if (rc_expand != 0) return rc_expand;
  a.part_rank = part_rank; a.part_count = part_count > 0 ? part_count : 1;
  if (a.part_count > 1 && part_rank != 0) {
    std::fill(wcount.begin(), wcount.end(), 0ull);       // the replicated expansion is reported by rank 0 only
    if (m.objective == CSOLVE_OBJ_ALL) ctl.n_stored = 0;
  }
  res->frontier = n_items;
  a.items = pin; a.items_out = nullptr;
  a.pool = pin; a.pool_cap = pool_cap; a.ready = ready.data(); a.n_initial = n_items;
  a.front_pool = pin; a.front_ctl = &ctl; a.total_warps = n_warps;
  a.gprio = prefer_failing ? gprio.data() : nullptr;
  if (learn) a.ng = ng;
  ctl.item_next = 0; ctl.item_count = 0; ctl.init_next = 0; ctl.busy = 0; ctl.hungry = 0;
  ctl.signal = stopped ? SIG_STOP : SIG_RUN;
  const bool sat = !general && search_uses_sat(m, learn != 0, order);
  a.use_sat = sat ? 1 : 0;
  l.fn = reinterpret_cast<void (*)(const SearchArgs)>(const_cast<void *>(search_kernel(m, false, learn != 0, sample, sat, false)));
  const size_t smem = search_smem_bytes(m, learn != 0, sat);
  void (*const fn_dfs)(const SearchArgs) = l.fn;
  // restarts (capi.cu; src/csolve.c:76-83, 264-276): Luby thresholds in units of restart_frequency failed nodes per warp
  const bool restarting = restart_frequency > 0 && m.objective == CSOLVE_OBJ_ANY && a.gprio != nullptr && !batch && !learn;
  unsigned long long luby_counter = 1, luby_threshold = 1;
  auto fail_limit_now = [&]() { return (int32_t)std::min((double)luby_threshold * (double)restart_frequency * (double)n_warps, 2.0e9); };
  if (restarting) a.fail_limit = fail_limit_now();
  for (; !stopped && n_items > 0;) {
    emu::launch(n_blocks, THREADS_PER_BLOCK, smem, run_kernel, &l);
    res->switches += emu::M.switches; res->collectives += emu::M.collectives; res->site_mismatches += emu::M.site_mismatches;
    emu::launch(1, 1024, 0, run_rebalance, &l);
    res->switches += emu::M.switches; res->collectives += emu::M.collectives; res->site_mismatches += emu::M.site_mismatches;
    res->slices++;
    if (sinking && ctl.n_stored > 0) {
      if (ctl.n_stored > sol_cap) { g_err = "solution buffer overflow: " + std::to_string(ctl.n_stored) + " in one slice, room for " + std::to_string(sol_cap); return -102; }
      if (sunk + ctl.n_stored > sink_rows) { g_err = "more solutions than the caller expects"; return -103; }
      memcpy(solutions + (size_t)sunk * (V + 1), solbuf.data(), sizeof(int32_t) * (size_t)ctl.n_stored * (V + 1));
      sunk += ctl.n_stored;
      ctl.n_stored = 0;
    }
    if (ctl.signal == SIG_STOP || ctl.busy == 0) break;
    if (res->slices > 1000000) { g_err = "the search does not end"; return -101; }
    if (restarting && ctl.fails > a.fail_limit) {
      // RESTART: every open frame is dropped, the root is expanded again in the order of the priorities learned so far
      if ((luby_counter & (~luby_counter + 1)) == luby_threshold) { luby_counter++; luby_threshold = 1; } else { luby_threshold <<= 1; }
      res->restarts++;
      std::vector<int32_t> ord(V);
      for (int v = 0; v < V; v++) ord[v] = v;
      std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return gprio[x] > gprio[y]; });
      std::copy(ord.begin(), ord.end(), cm.order.begin());       // DevModel::order points into cm.order
      upload_root_frame(ord[0]);
      std::fill(ws.begin(), ws.end(), WarpState{-1, 0, 0, 0u});
      pin = pool.data(); pout = pool_b.data(); n_items = 1;
      a.gprio = nullptr; a.fail_limit = 0;
      ctl.signal = SIG_RUN;
      expand_root();
      if (rc_expand != 0) return rc_expand;
      a.items = pin; a.items_out = nullptr;
      a.pool = pin; a.front_pool = pin; a.n_initial = n_items; a.gprio = gprio.data();
      a.fail_limit = fail_limit_now();
      std::fill(ready.begin(), ready.end(), 0);
      ctl.item_next = 0; ctl.item_count = 0; ctl.init_next = 0; ctl.busy = 0; ctl.hungry = 0; ctl.fails = 0;
      l.fn = fn_dfs;
      if (stopped || n_items == 0) break;
    }
  }
  for (int w = 0; w < n_warps; w++) {
    const unsigned long long *c = &wcount[(size_t)w * CNT_WIDTH];
    res->nodes += c[CNT_NODES]; res->cuts += c[CNT_CUTS]; res->props += c[CNT_PROPS]; res->solutions += c[CNT_SOLUTIONS];
    res->claims += (int32_t)c[CNT_CLAIMS];
    if ((ws[w].level >= ws[w].base || ws[w].claim_mask != 0u) && ctl.signal != SIG_STOP) {
      g_err = "a warp left the kernel with open frames: warp " + std::to_string(w) + " level " + std::to_string(ws[w].level) + " base " + std::to_string(ws[w].base) +
              ", signal " + std::to_string(ctl.signal) + " hungry " + std::to_string(ctl.hungry) + " tickets " + std::to_string(ctl.item_next) + "/" + std::to_string(ctl.item_count);
      return -100;
    }   // ANY: the first solution stops everybody
  }
  {
    // the device's own reduction of the per-warp counters (k_reduce_counters, what capi.cu reads) against the sums above
    static ReduceLaunch rl;
    std::vector<unsigned long long> totals(CNT_WIDTH, 0);
    rl = ReduceLaunch{wcount.data(), n_warps, totals.data()};
    const int grid = (int)std::min<size_t>(32, ((size_t)n_warps * CNT_WIDTH + REDUCE_THREADS - 1) / REDUCE_THREADS);
    emu::launch(std::max(grid, 1), 1024, 0, run_reduce, &rl);
    if (totals[CNT_NODES] != res->nodes || totals[CNT_CUTS] != res->cuts || totals[CNT_SOLUTIONS] != res->solutions || totals[CNT_PROPS] != res->props) {
      g_err = "k_reduce_counters disagrees with the per-warp counters"; return -120;
    }
  }
  if (batch) for (int r = 0; r < n_roots; r++) root_solutions[r] = rsol[r];
  if (sample) { g_sample_seen = sample_n; g_samples.resize((size_t)std::min(sample_n, sample_cap) * sample_words(V)); }
  res->best = ctl.best;
  res->has_solution = res->solutions > 0;
  res->n_stored = ctl.n_stored;
  g_ng_flat.clear();
  if (learn) {
    const int n_ng = std::min(ng_counters[0], ng.cap_ng);
    for (int k = 0; k < n_ng; k++) {
      g_ng_flat.push_back(ng_len[k]);
      for (int j = 0; j < ng_len[k]; j++) g_ng_flat.push_back(ng_lits[ng_start[k] + j]);
    }
  }
  if (learn) { res->conflicts = ng_counters[0]; res->conflicts_abandoned = ng_counters[3] + ng_counters[4]; res->backjumps = ng_counters[5]; }
  if (sinking) res->n_stored = (int32_t)sunk;
  if (solutions != nullptr && !sinking) {
    // MIN / MAX: the buffer is a ring of improving incumbents (slots are drawn after the incumbent is updated, so the
    // order of two warps' entries may be swapped): best key first, as capi.cu sorts them
    const int n = ctl.n_stored < sol_cap ? ctl.n_stored : sol_cap;
    std::vector<int> idx(n);
    for (int k = 0; k < n; k++) idx[k] = k;
    if (m.obj_var >= 0)
      std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) {
        const int kx = solbuf[(size_t)x * (V + 1) + V], ky = solbuf[(size_t)y * (V + 1) + V];
        return m.objective == CSOLVE_OBJ_MIN ? kx < ky : kx > ky;
      });
    for (int k = 0; k < n; k++) memcpy(solutions + (size_t)k * (V + 1), &solbuf[(size_t)idx[k] * (V + 1)], sizeof(int32_t) * (V + 1));
  }
  return 0;
}

extern "C" int emu_search(const csolve_flat_model *fm, int order, int learn, int prefer_failing, int n_blocks,
                          int max_solutions, int general, long long slice_clock, int sink_headroom, int sink_rows,
                          int split_target, int part_rank, int part_count, int restart_frequency, emu_result *res, int32_t *solutions) {
  return search_core(fm, order, learn, prefer_failing, n_blocks, max_solutions, general, slice_clock, sink_headroom, sink_rows,
                     split_target, part_rank, part_count, restart_frequency, 0, nullptr, nullptr, nullptr, res, solutions);
}

// csolve_gpu_solve_batch: root_dom [n_roots][2 * n_vars]; root_solutions [n_roots], root_failed [n_roots] come back
extern "C" int emu_search_batch(const csolve_flat_model *fm, int order, int n_blocks, int general, long long slice_clock,
                                int split_target, int n_roots, const int32_t *root_dom, uint32_t *root_solutions,
                                uint8_t *root_failed, emu_result *res) {
  return search_core(fm, order, 0, 0, n_blocks, 0, general, slice_clock, 0, 0, split_target, 0, 1, 0, n_roots, root_dom,
                     root_solutions, root_failed, res, nullptr);
}

// ---- several GPUs on one tree (csolve_gpu_comm; ANY / MIN / MAX models) ----------------------------------------------
// `world` emulated ranks, n_blocks blocks each, run side by side in ONE emulated launch: rank 0 has expanded the root,
// every rank claims frames of that one frontier (front_ctl->init_next), incumbents / "found" travel through the ranks'
// CommBlocks (comm_push_best / comm_push_stop / comm_poll), a rank that is running dry asks its peers from its waiting
// loop (CommBlock::demand) and their busy warps serve its donation ring (donation_target, reserve_slot, publish_slot).
// Between slices every rank that searched runs k_rebalance (which closes its ring); a rank that ran dry does what the
// host's idle loop does (capi.cu): k_comm_state clears its busy mark in rank 0's count of active ranks, leaves its ring
// open and raises its demand with the peers -- whose warps serve it during the next launch -- until frames have arrived
// (it joins the next launch) or no rank is active any more (the search is over everywhere).
namespace {
struct CommLaunch {
  std::vector<SearchArgs> a; void (*fn)(const SearchArgs); int n_blocks; int32_t *scratch; int rank_for_rebalance;
  std::vector<int> active;          // the ranks whose search kernels run side by side in this launch
  int want_frames; int32_t *state_out;
};
void run_comm_kernel(void *arg) {
  const CommLaunch *c = static_cast<const CommLaunch *>(arg);
  c->fn(c->a[c->active[emu::M.cur->block / c->n_blocks]]);
}
void run_comm_state(void *arg) {
  const CommLaunch *c = static_cast<const CommLaunch *>(arg);
  k_comm_state(c->a[c->rank_for_rebalance], c->want_frames, 0, c->state_out);
}
void run_comm_rebalance(void *arg) {
  const CommLaunch *c = static_cast<const CommLaunch *>(arg);
  k_rebalance(c->a[c->rank_for_rebalance], c->scratch);
}
}  // namespace

// learn: every rank learns into a nogood pool of its own (-c under -j N; with -DCSOLVE_BJ the ranks back-jump)
extern "C" int emu_search_comm(const csolve_flat_model *fm, int order, int prefer_failing, int n_blocks, int world,
                               int split_target, long long slice_clock, int general, int learn, emu_result *res, int32_t *solution) {
  CompiledModel cm;
  int rc = compile_model(*fm, cm, g_err);
  if (rc != 0) return rc;
  DevModel m = cm.host;
  if (general || learn) { m.lov = 0; m.lovk = 0; }
  if (m.objective == CSOLVE_OBJ_ALL) { g_err = "ALL models are dealt by path hash (emu_search with part_count)"; return -110; }
  if (world < 1 || world > COMM_MAX_RANKS) { g_err = "bad world"; return -111; }
  if (prefer_failing && m.lov) prefer_failing = 0;
  const int V = m.n_vars, fw = m.frame_words, n_warps = n_blocks * WARPS_PER_BLOCK, epoch = 1;
  memset(res, 0, sizeof(*res));
  const int ring = 4 * n_warps + 1024;
  const int target = split_target > 1 ? split_target : 1;
  const int front_cap = std::max(4 * target, 1024) + ring;
  std::vector<int32_t> pool_a((size_t)front_cap * fw, g_fill), pool_b((size_t)front_cap * fw, g_fill);
  const int sol_cap = 64;

  struct Rank {
    SearchCtl ctl; CommBlock blk;
    std::vector<int32_t> stacks, ring_frames, ready, solbuf;
    std::vector<WarpState> ws; std::vector<unsigned long long> wcount;
    std::vector<int32_t> ng_lits, ng_start, ng_len, ng_watch, ng_watch_n, ng_counters;
    NogoodPool ng;
  };
  std::vector<Rank> R(world);
  std::vector<int32_t> gprio(cm.prio.begin(), cm.prio.end());      // per GPU in the product; one here (a heuristic)
  for (auto &r : R) {
    memset(&r.ctl, 0, sizeof(r.ctl)); memset(&r.blk, 0, sizeof(r.blk));
    r.ctl.best = m.objective == CSOLVE_OBJ_MIN ? INT32_MAX : (m.objective == CSOLVE_OBJ_MAX ? INT32_MIN : 0);
    r.blk.rmin64 = ~0ull; r.blk.rmax64 = 0ull;
    r.stacks.assign((size_t)n_warps * (V + 1) * fw, g_fill); r.ring_frames.assign((size_t)ring * fw, g_fill); r.ready.assign(ring, 0);
    r.solbuf.assign((size_t)sol_cap * (V + 1), g_fill);
    r.ws.assign(n_warps, WarpState{-1, 0, 0, 0u}); r.wcount.assign((size_t)n_warps * CNT_WIDTH, 0);
    memset(&r.ng, 0, sizeof(r.ng));
    if (learn) {
      r.ng.cap_ng = 1 << 13; r.ng.cap_lits = 1 << 17; r.ng.cap_w = 512;
      r.ng_lits.assign(r.ng.cap_lits, 0); r.ng_start.assign(r.ng.cap_ng, 0); r.ng_len.assign(r.ng.cap_ng, 0);
      r.ng_watch.assign((size_t)V * r.ng.cap_w, -1); r.ng_watch_n.assign(V, 0); r.ng_counters.assign(8, 0);
      r.ng.lits = r.ng_lits.data(); r.ng.start = r.ng_start.data(); r.ng.len = r.ng_len.data();
      r.ng.watch = r.ng_watch.data(); r.ng.watch_n = r.ng_watch_n.data(); r.ng.counters = r.ng_counters.data();
    }
  }
  // rank 0 expands the root (the same loop as search_core's, without batch / sink)
  int32_t *pin = pool_a.data(), *pout = pool_b.data();
  {
    const int rv = root_var(cm, order);
    int32_t *root = pin;
    std::fill(root, root + fw, 0);
    root[FR_VAR] = rv; root[FR_LO] = cm.root_dom[2 * rv]; root[FR_HI] = cm.root_dom[2 * rv + 1];
    root[FR_LAST] = (int32_t)((uint32_t)root[FR_HI] - (uint32_t)root[FR_LO]);
    root[FR_BEST] = R[0].ctl.best; root[7] = 0x1234567;
    memcpy(&root[frame_dom_offset(m.mask_words)], cm.root_dom.data(), sizeof(int32_t) * 2 * V);
    if (m.lovk) memcpy(&root[frame_dom_offset(m.mask_words) + 2 * V], cm.lov_fconst.data(), sizeof(int32_t) * V);
  }
  Launch l;
  memset(&l.a, 0, sizeof(l.a));
  SearchArgs &a0 = l.a;
  a0.m = m; a0.ctl = &R[0].ctl; a0.stacks = R[0].stacks.data(); a0.wstate = R[0].ws.data(); a0.wcount = R[0].wcount.data();
  a0.solbuf = R[0].solbuf.data(); a0.max_solutions = sol_cap; a0.n_warps = n_warps; a0.order = order;
  a0.out_cap = front_cap; a0.expand_branch_max = 64; a0.part_count = 1; a0.slice_cycles = LLONG_MAX / 2;
  int n_items = 1;
  bool stopped = false;
  {
    long long max_branch = 1;
    for (int v = 0; v < V; v++) max_branch = std::max<long long>(max_branch, (long long)cm.root_dom[2 * v + 1] - cm.root_dom[2 * v] + 1);
    max_branch = std::min<long long>(max_branch, a0.expand_branch_max);
    l.fn = reinterpret_cast<void (*)(const SearchArgs)>(const_cast<void *>(search_kernel(m, true, false, false, false, false)));
    const size_t smem_x = search_smem_bytes(m, false, false);
    SearchCtl &ctl = R[0].ctl;
    for (int lvl = 0; lvl < V && lvl < 24 && n_items > 0 && n_items < target; ++lvl) {
      const int before = n_items;
      if ((long long)n_items * max_branch > front_cap - ring) break;
      a0.items = pin; a0.items_out = pout; a0.frozen_best = ctl.best;
      ctl.item_next = 0; ctl.item_count = n_items; ctl.out_count = 0; ctl.passed = 0;
      emu::launch(std::min(n_blocks, (n_items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK), THREADS_PER_BLOCK, smem_x, run_kernel, &l);
      if (ctl.out_dropped > 0) { g_err = "frontier pool overflow during expansion"; return -104; }
      n_items = ctl.out_count;
      std::swap(pin, pout);
      if (ctl.signal == SIG_STOP) { stopped = true; break; }
      if (ctl.passed == n_items) break;
      if (n_items >= n_warps / 2 && n_items < 2 * (long long)before) break;
    }
  }
  res->frontier = n_items;
  // every rank: the shared frontier is rank 0's, the donation ring its own, slot numbers the same everywhere
  CommLaunch c;
  c.a.assign(world, a0);
  c.n_blocks = n_blocks;
  const bool sat = !general && search_uses_sat(m, learn != 0, order);
  c.fn = reinterpret_cast<void (*)(const SearchArgs)>(const_cast<void *>(search_kernel(m, false, learn != 0, false, sat, false)));
  std::vector<int32_t> scratch(4 + 3 * n_warps, 0);
  c.scratch = scratch.data();
  const int best0 = R[0].ctl.best;
  for (int r = 0; r < world; r++) {
    SearchArgs &a = c.a[r];
    a.ctl = &R[r].ctl; a.stacks = R[r].stacks.data(); a.wstate = R[r].ws.data(); a.wcount = R[r].wcount.data();
    a.solbuf = R[r].solbuf.data();
    a.items = pin; a.items_out = nullptr;
    a.pool_cap = n_items + ring; a.n_initial = n_items;
    a.pool = R[r].ring_frames.data() - (size_t)n_items * fw;
    a.ready = R[r].ready.data() - n_items;
    a.front_pool = pin; a.front_ctl = &R[0].ctl;
    a.total_warps = n_warps * world;
    a.comm = &R[r].blk; a.epoch = epoch; a.rank = r; a.world = world; a.n_peers = world - 1;
    for (int q = 0; q < world; q++) {
      a.peer_comm[q] = &R[q].blk; a.peer_ctl[q] = &R[q].ctl;
      a.peer_pool[q] = R[q].ring_frames.data() - (size_t)n_items * fw;
      a.peer_ready[q] = R[q].ready.data() - n_items;
    }
    a.peer_demand = world > 1 ? 1 : 0;
    a.use_sat = sat ? 1 : 0;
    a.gprio = prefer_failing ? gprio.data() : nullptr;
    if (learn) a.ng = R[r].ng;
    a.slice_cycles = slice_clock > 0 ? slice_clock : LLONG_MAX / 2;
    const int keep_stored = r == 0 ? R[0].ctl.n_stored : 0;
    SearchCtl &ctl = R[r].ctl;
    if (r != 0) memset(&ctl, 0, sizeof(ctl));
    ctl.best = best0; ctl.n_stored = keep_stored;
    ctl.item_next = 0; ctl.item_count = 0; ctl.init_next = 0; ctl.busy = 0; ctl.hungry = 0;
    ctl.signal = stopped ? SIG_STOP : SIG_RUN;
    R[r].blk.busy_epoch = epoch;
  }
  const size_t smem = search_smem_bytes(m, learn != 0, sat);
  R[0].blk.active64 = ((unsigned long long)(unsigned)epoch << 32) | (unsigned)world;      // capi.cu: publish_epoch
  std::vector<int> has_work(world, 1);
  int32_t state[4] = {0, 0, 0, 0};
  c.state_out = state;
  c.want_frames = world > 1 ? std::max(n_warps / (4 * (world - 1)), 32) : 0;
  for (; !stopped && n_items > 0;) {
    c.active.clear();
    for (int r = 0; r < world; r++) if (has_work[r]) { c.active.push_back(r); R[r].blk.ring_open = epoch; }
    if (c.active.empty()) break;
    emu::block_group = n_blocks;
    emu::launch(n_blocks * (int)c.active.size(), THREADS_PER_BLOCK, smem, run_comm_kernel, &c);
    res->switches += emu::M.switches; res->collectives += emu::M.collectives;
    emu::block_group = 0;
    bool any_stop = false;
    for (int r : c.active) {
      c.rank_for_rebalance = r;
      emu::launch(1, 1024, 0, run_comm_rebalance, &c);
      has_work[r] = R[r].ctl.busy != 0;
      any_stop |= R[r].ctl.signal == SIG_STOP;
    }
    res->slices++;
    if (any_stop) break;
    // the ranks without work: the host's idle loop, one look each
    int active_ranks = -1;
    bool peer_found = false;
    for (int r = 0; r < world && world > 1; r++) {
      if (has_work[r]) continue;
      c.rank_for_rebalance = r;
      emu::launch(1, 1, 0, run_comm_state, &c);
      if (state[2]) peer_found = true;
      if (state[0]) has_work[r] = 1;
      active_ranks = state[1];
      if (state[1] < 0) { g_err = "comm: the ranks are out of step"; return -112; }
    }
    if (peer_found) break;
    bool any = false;
    for (int r = 0; r < world; r++) any |= has_work[r] != 0;
    if (!any) {
      if (world > 1 && active_ranks != 0) { g_err = "every rank is idle but rank 0 counts " + std::to_string(active_ranks) + " active ranks"; return -113; }
      break;
    }
    if (res->slices > 100000) { g_err = "the search does not end"; return -101; }
  }
  // reduce (capi.cu: the ranks' results are added up, the incumbent is the best one, its witness comes from its rank)
  res->best = best0;
  int best_rank = -1;
  for (int r = 0; r < world; r++) {
    for (int w = 0; w < n_warps; w++) {
      const unsigned long long *cn = &R[r].wcount[(size_t)w * CNT_WIDTH];
      res->nodes += cn[CNT_NODES]; res->cuts += cn[CNT_CUTS]; res->solutions += cn[CNT_SOLUTIONS]; res->claims += (int32_t)cn[CNT_CLAIMS];
      if ((R[r].ws[w].level >= R[r].ws[w].base || R[r].ws[w].claim_mask != 0u) && m.objective != CSOLVE_OBJ_ANY) {
        g_err = "rank " + std::to_string(r) + " warp " + std::to_string(w) + " left with open frames"; return -100;
      }
    }
    if (learn) { res->conflicts += R[r].ng_counters[0]; res->conflicts_abandoned += R[r].ng_counters[3] + R[r].ng_counters[4]; res->backjumps += R[r].ng_counters[5]; }
    const int b = R[r].ctl.best;
    if (m.objective == CSOLVE_OBJ_MIN ? b < res->best : (m.objective == CSOLVE_OBJ_MAX ? b > res->best : false)) res->best = b;
  }
  res->has_solution = res->solutions > 0;
  // witness: the stored assignment whose key is the optimum (MIN / MAX), the first stored one (ANY)
  for (int r = 0; r < world && best_rank < 0; r++) {
    const int n = std::min(R[r].ctl.n_stored, sol_cap);
    for (int k = 0; k < n; k++) {
      const int32_t *row = &R[r].solbuf[(size_t)k * (V + 1)];
      if (m.objective == CSOLVE_OBJ_ANY || row[V] == res->best) { memcpy(solution, row, sizeof(int32_t) * (V + 1)); best_rank = r; res->n_stored = 1; break; }
    }
  }
  return 0;
}

// ---- csolve_gpu_propagate_batch: n independent node transitions (the parity hook) ------------------------------------
namespace {
struct PropLaunch { DevModel m; int n; const int32_t *dom_in, *var, *val, *best; int32_t *dom_out; uint8_t *failed; };
void run_propagate_batch(void *arg) {
  const PropLaunch *q = static_cast<const PropLaunch *>(arg);
  if (q->m.lovk) {
    switch (q->m.lovk) {
    case 2: k_propagate_batch_lovk<2>(q->m, q->n, q->dom_in, q->var, q->val, q->dom_out, q->failed); break;
    case 3: k_propagate_batch_lovk<3>(q->m, q->n, q->dom_in, q->var, q->val, q->dom_out, q->failed); break;
    default: k_propagate_batch_lovk<4>(q->m, q->n, q->dom_in, q->var, q->val, q->dom_out, q->failed); break;
    }
  } else if (q->m.lov) {
    k_propagate_batch_lov(q->m, q->n, q->dom_in, q->var, q->val, q->dom_out, q->failed);
  } else {
    k_propagate_batch(q->m, q->n, q->dom_in, q->var, q->val, q->best, q->dom_out, q->failed);
  }
}
}  // namespace

// general: 1 = the general kernel whatever the model; best may be NULL (zeros, as capi.cu passes)
extern "C" int emu_propagate_batch(const csolve_flat_model *fm, int general, int n_blocks, int n, const int32_t *dom_in,
                                   const int32_t *var, const int32_t *val, const int32_t *best, int32_t *dom_out, uint8_t *failed) {
  CompiledModel cm;
  int rc = compile_model(*fm, cm, g_err);
  if (rc != 0) return rc;
  DevModel m = cm.host;
  if (general) { m.lov = 0; m.lovk = 0; }
  std::vector<int32_t> zero_best;
  if (best == nullptr) { zero_best.assign(n, 0); best = zero_best.data(); }
  PropLaunch q{m, n, dom_in, var, val, best, dom_out, failed};
  const int grid = std::max(1, std::min((n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, n_blocks));
  emu::launch(grid, THREADS_PER_BLOCK, m.lovk ? 0 : search_smem_bytes(m, false), run_propagate_batch, &q);
  return 0;
}

// ---- frames shipped between ranks at the slice boundaries (csolve_gpu_set_rebalance + csolve_gpu_export_frames /
//      csolve_gpu_import_frames: the per-slice rank hooks of round 1) ----------------------------------------------
// `world` ranks search their path-hash shares of an ALL model's tree slice by slice; after every round a rank that ran
// dry receives frames that k_export_frames split off the parked stacks of the busiest rank, k_import_frames puts them
// into its donation ring as tickets served ahead of their holders (csolve_b200/distributed.py: make_rebalance).
namespace {
struct XferLaunch { const SearchArgs *a; int32_t *buf; int max_frames; int32_t *n_out; int n_frames; };
void run_export(void *arg) { const XferLaunch *x = static_cast<const XferLaunch *>(arg); k_export_frames(*x->a, x->buf, x->max_frames, x->n_out); }
void run_import(void *arg) { const XferLaunch *x = static_cast<const XferLaunch *>(arg); k_import_frames(*x->a, x->buf, x->n_frames); }
}  // namespace

extern "C" int emu_search_exchange(const csolve_flat_model *fm, int order, int n_blocks, int world, int split_target,
                                   long long slice_clock, int general, emu_result *res, int32_t *frames_moved) {
  CompiledModel cm;
  int rc = compile_model(*fm, cm, g_err);
  if (rc != 0) return rc;
  DevModel m = cm.host;
  if (general) { m.lov = 0; m.lovk = 0; }
  if (m.objective != CSOLVE_OBJ_ALL) { g_err = "ALL models only"; return -110; }
  const int V = m.n_vars, fw = m.frame_words, n_warps = n_blocks * WARPS_PER_BLOCK;
  memset(res, 0, sizeof(*res));
  *frames_moved = 0;
  const int ring = 4 * n_warps + 1024;
  const int target = split_target > 1 ? split_target : 1;
  const int pool_cap = std::max(4 * target, 1024) + ring;
  struct Rank {
    SearchCtl ctl;
    std::vector<int32_t> stacks, pool, ready, solbuf, scratch;
    std::vector<WarpState> ws; std::vector<unsigned long long> wcount;
    SearchArgs a;
  };
  std::vector<Rank> R(world);
  std::vector<int32_t> pool_b((size_t)pool_cap * fw, g_fill);
  for (auto &r : R) {
    memset(&r.ctl, 0, sizeof(r.ctl));
    r.stacks.assign((size_t)n_warps * (V + 1) * fw, g_fill); r.pool.assign((size_t)pool_cap * fw, g_fill); r.ready.assign(pool_cap, 0);
    r.solbuf.assign(V + 1, 0); r.scratch.assign(4 + 3 * n_warps, 0);
    r.ws.assign(n_warps, WarpState{-1, 0, 0, 0u}); r.wcount.assign((size_t)n_warps * CNT_WIDTH, 0);
  }
  // rank 0 expands the root; every rank holds a copy of the frontier (the expansion is replicated in the product)
  Launch l;
  SearchArgs &a0 = l.a;
  memset(&a0, 0, sizeof(a0));
  int32_t *pin = R[0].pool.data(), *pout = pool_b.data();
  {
    const int rv = root_var(cm, order);
    int32_t *root = pin;
    std::fill(root, root + fw, 0);
    root[FR_VAR] = rv; root[FR_LO] = cm.root_dom[2 * rv]; root[FR_HI] = cm.root_dom[2 * rv + 1];
    root[FR_LAST] = (int32_t)((uint32_t)root[FR_HI] - (uint32_t)root[FR_LO]); root[7] = 0x1234567;
    memcpy(&root[frame_dom_offset(m.mask_words)], cm.root_dom.data(), sizeof(int32_t) * 2 * V);
    if (m.lovk) memcpy(&root[frame_dom_offset(m.mask_words) + 2 * V], cm.lov_fconst.data(), sizeof(int32_t) * V);
  }
  a0.m = m; a0.ctl = &R[0].ctl; a0.stacks = R[0].stacks.data(); a0.wstate = R[0].ws.data(); a0.wcount = R[0].wcount.data();
  a0.solbuf = R[0].solbuf.data(); a0.max_solutions = 1; a0.n_warps = n_warps; a0.order = order;
  a0.out_cap = pool_cap; a0.expand_branch_max = 64; a0.part_count = 1; a0.slice_cycles = LLONG_MAX / 2;
  int n_items = 1;
  {
    long long max_branch = 1;
    for (int v = 0; v < V; v++) max_branch = std::max<long long>(max_branch, (long long)cm.root_dom[2 * v + 1] - cm.root_dom[2 * v] + 1);
    max_branch = std::min<long long>(max_branch, a0.expand_branch_max);
    l.fn = reinterpret_cast<void (*)(const SearchArgs)>(const_cast<void *>(search_kernel(m, true, false, false, false, false)));
    const size_t smem_x = search_smem_bytes(m, false, false);
    SearchCtl &ctl = R[0].ctl;
    for (int lvl = 0; lvl < V && lvl < 24 && n_items > 0 && n_items < target; ++lvl) {
      const int before = n_items;
      if ((long long)n_items * max_branch > pool_cap - ring) break;
      a0.items = pin; a0.items_out = pout;
      ctl.item_next = 0; ctl.item_count = n_items; ctl.out_count = 0; ctl.passed = 0;
      emu::launch(std::min(n_blocks, (n_items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK), THREADS_PER_BLOCK, smem_x, run_kernel, &l);
      if (ctl.out_dropped > 0) { g_err = "frontier pool overflow during expansion"; return -104; }
      n_items = ctl.out_count;
      std::swap(pin, pout);
      if (ctl.passed == n_items) break;
      if (n_items >= n_warps / 2 && n_items < 2 * (long long)before) break;
    }
  }
  if (pin != R[0].pool.data()) std::copy(pin, pin + (size_t)n_items * fw, R[0].pool.data());
  const bool sat = !general && search_uses_sat(m, false, order);
  void (*const fn_dfs)(const SearchArgs) = reinterpret_cast<void (*)(const SearchArgs)>(const_cast<void *>(search_kernel(m, false, false, false, sat, false)));
  for (int r = 0; r < world; r++) {
    if (r != 0) std::copy(R[0].pool.begin(), R[0].pool.begin() + (size_t)n_items * fw, R[r].pool.begin());
    SearchArgs &a = R[r].a;
    a = a0;
    a.ctl = &R[r].ctl; a.stacks = R[r].stacks.data(); a.wstate = R[r].ws.data(); a.wcount = R[r].wcount.data(); a.solbuf = R[r].solbuf.data();
    a.items = R[r].pool.data(); a.items_out = nullptr; a.pool = R[r].pool.data(); a.pool_cap = pool_cap; a.ready = R[r].ready.data();
    a.n_initial = n_items; a.front_pool = R[r].pool.data(); a.front_ctl = &R[r].ctl; a.total_warps = n_warps;
    a.part_rank = r; a.part_count = world; a.use_sat = sat ? 1 : 0;
    a.slice_cycles = slice_clock > 0 ? slice_clock : LLONG_MAX / 2;
    SearchCtl &ctl = R[r].ctl;
    const int keep = r == 0 ? 1 : 0;
    if (!keep) { memset(&ctl, 0, sizeof(ctl)); }
    ctl.item_next = 0; ctl.item_count = 0; ctl.init_next = 0; ctl.busy = 0; ctl.hungry = 0; ctl.signal = SIG_RUN;
  }
  const size_t smem = search_smem_bytes(m, false, sat);
  std::vector<int> busy(world, 1);
  std::vector<int32_t> xbuf;
  for (int round = 0; n_items > 0; round++) {
    for (int r = 0; r < world; r++) {
      if (!busy[r]) continue;
      l.a = R[r].a; l.fn = fn_dfs; l.scratch = R[r].scratch.data();
      emu::launch(n_blocks, THREADS_PER_BLOCK, smem, run_kernel, &l);
      emu::launch(1, 1024, 0, run_rebalance, &l);
      busy[r] = R[r].ctl.busy;
      res->slices++;
    }
    int donor = -1;
    for (int r = 0; r < world; r++) if (busy[r] > 0 && (donor < 0 || busy[r] > busy[donor])) donor = r;
    if (donor < 0) break;
    // ranks that ran dry get frames of the busiest rank (at most a quarter of their warps each)
    for (int r = 0; r < world; r++) {
      if (busy[r] != 0 || r == donor) continue;
      const int want = std::max(1, n_warps / 4);
      xbuf.assign((size_t)want * fw, 0);
      int32_t n_out = 0;
      XferLaunch x{&R[donor].a, xbuf.data(), want, &n_out, 0};
      emu::launch(1, 1024, 0, run_export, &x);
      if (n_out <= 0) continue;
      x.a = &R[r].a; x.n_frames = n_out;
      emu::launch(1, 1024, 0, run_import, &x);
      *frames_moved += n_out;
      busy[r] = 1;
    }
    if (round > 100000) { g_err = "the search does not end"; return -101; }
  }
  for (int r = 0; r < world; r++)
    for (int w = 0; w < n_warps; w++) {
      const unsigned long long *cn = &R[r].wcount[(size_t)w * CNT_WIDTH];
      res->nodes += cn[CNT_NODES]; res->cuts += cn[CNT_CUTS]; res->solutions += cn[CNT_SOLUTIONS]; res->claims += (int32_t)cn[CNT_CLAIMS];
      if (R[r].ws[w].level >= R[r].ws[w].base || R[r].ws[w].claim_mask != 0u) { g_err = "a warp left with open frames"; return -100; }
    }
  res->has_solution = res->solutions > 0;
  res->frontier = n_items;
  return 0;
}
