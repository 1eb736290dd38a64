// capi.cu -- the C ABI of include/csolve_b200.h: device residency of the compiled
// model and host orchestration of the search (frontier expansion, time-sliced
// persistent search, rebalancing, result collection).
//
// There is no CPU implementation of the search behind these entry points: without a
// CUDA device every device call fails with CSOLVE_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "csolve_b200.h"
#include "compile.hpp"
#include "front.hpp"
#include "kernels.cuh"

using namespace csolve_dev;

namespace {

int fail(int code, const std::string &msg) {
  csolve_front::set_last_error(msg);
  return code;
}

#define CUDA_TRY(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      return fail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver        \
                      ? CSOLVE_ERR_NO_DEVICE : CSOLVE_ERR_CUDA,                         \
                  std::string(#expr) + ": " + cudaGetErrorString(e__));                 \
    }                                                                                   \
  } while (0)

// ---- per-device state ---------------------------------------------------------------------------------------
// One context per CUDA device, created on first use and kept for the life of the process. A problem belongs to
// the device it was loaded on; every entry point makes that device current for the calling thread, so problems on
// different devices can be driven from different host threads (csolve_gpu_group does exactly that).
//
// Device memory blocks (stacks, frontier pools, the small per-problem buffers) are recycled across the problems of
// one device: cudaMalloc / cudaFree cost milliseconds each once the process has peer mappings (measured: 20-40 ms of
// end-to-end time per search), a search takes about as long. The cache is bounded in BYTES; when an allocation fails
// the cache is released and the allocation retried once.
struct WsBlock { void *p; size_t bytes; };

struct DeviceCtx {
  int device = -1;
  int sm_count = 0;
  int clock_khz = 0;
  std::mutex mu;                       // guards the two lists below
  std::vector<WsBlock> free_blocks;    // cached, not in use
  std::vector<WsBlock> live;           // handed out by cmalloc (true size of a recycled block)
  size_t free_bytes = 0;
  static constexpr size_t kMaxCachedBytes = (size_t)6 << 30;

  cudaError_t alloc(void **out, size_t bytes, size_t *got) {
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t i = 0; i < free_blocks.size(); i++) {
        if (free_blocks[i].bytes >= bytes && free_blocks[i].bytes <= bytes + bytes / 4 + 4096) {
          *out = free_blocks[i].p;
          if (got) *got = free_blocks[i].bytes;
          free_bytes -= free_blocks[i].bytes;
          free_blocks.erase(free_blocks.begin() + i);
          return cudaSuccess;
        }
      }
    }
    if (got) *got = bytes;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {
      cudaGetLastError();              // clear the sticky error, drop the cache, try once more
      release_all();
      e = cudaMalloc(out, bytes);
    }
    return e;
  }
  void release(void *p, size_t bytes) {
    if (p == nullptr) return;
    std::vector<void *> evict;
    {
      std::lock_guard<std::mutex> lk(mu);
      free_blocks.push_back(WsBlock{p, bytes});
      free_bytes += bytes;
      while (!free_blocks.empty() && (free_bytes > kMaxCachedBytes || free_blocks.size() > 256)) {
        evict.push_back(free_blocks.front().p);
        free_bytes -= free_blocks.front().bytes;
        free_blocks.erase(free_blocks.begin());
      }
    }
    for (void *q : evict) cudaFree(q);
  }
  void release_all() {
    std::vector<WsBlock> blocks;
    {
      std::lock_guard<std::mutex> lk(mu);
      blocks.swap(free_blocks);
      free_bytes = 0;
    }
    for (auto &b : blocks) cudaFree(b.p);
  }
  cudaError_t cmalloc(void **out, size_t bytes) {
    bytes = std::max<size_t>(bytes, 256);
    size_t got = bytes;
    cudaError_t e = alloc(out, bytes, &got);
    if (e == cudaSuccess) {
      std::lock_guard<std::mutex> lk(mu);
      live.push_back(WsBlock{*out, got});
    }
    return e;
  }
  void cfree(void *p) {
    if (p == nullptr) return;
    size_t bytes = 0;
    bool found = false;
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t i = 0; i < live.size(); i++) {
        if (live[i].p == p) { bytes = live[i].bytes; live.erase(live.begin() + i); found = true; break; }
      }
    }
    if (found) release(p, bytes);     // a recycled block may be larger than what was asked for: keep its true size
    else cudaFree(p);
  }
};

std::mutex g_ctx_mu;
std::vector<std::unique_ptr<DeviceCtx>> g_ctx;      // indexed by device ordinal
std::atomic<int> g_default_device{-1};              // set by csolve_gpu_init(): the device csolve_gpu_load() uses

// context of `device` (created on first use); makes the device current for the calling thread
int device_ctx(int device, DeviceCtx **out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    return fail(CSOLVE_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                          " (the search path has no CPU fallback)");
  }
  if (device < 0 || device >= n) return fail(CSOLVE_ERR_INVALID, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  if ((int)g_ctx.size() < n) g_ctx.resize(n);
  if (!g_ctx[device]) {
    std::unique_ptr<DeviceCtx> c(new DeviceCtx);
    c->device = device;
    // queried once: cudaDevAttrClockRate is a live query that takes tens of milliseconds every now and then
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaDeviceGetAttribute(&c->clock_khz, cudaDevAttrClockRate, device));
    g_ctx[device] = std::move(c);
  }
  *out = g_ctx[device].get();
  return CSOLVE_OK;
}

template <class T> cudaError_t cmalloc(DeviceCtx *c, T **out, size_t bytes) { return c->cmalloc(reinterpret_cast<void **>(out), bytes); }

template <class T>
int upload(DeviceCtx *C, const std::vector<T> &h, const T **d) {
  T *p = nullptr;
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  CUDA_TRY(cmalloc(C, &p, bytes));
  if (!h.empty()) CUDA_TRY(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *d = p;
  return CSOLVE_OK;
}

// first statement of every entry point that touches the device: the problem's device becomes current for this thread
#define ENTER(p) DeviceCtx *C = (p)->ctx; CUDA_TRY(cudaSetDevice(C->device))

// runs a cleanup on every way out of a function (the CUDA_TRY returns included)
template <class F>
struct ScopeExit {
  F f;
  explicit ScopeExit(F fn) : f(fn) {}
  ScopeExit(const ScopeExit &) = delete;
  ~ScopeExit() { f(); }
};

// host rendering of warp_select_var() for the root frame
int select_root_var(const CompiledModel &cm, int order) {
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

}  // namespace

struct csolve_gpu_problem {
  DeviceCtx *ctx = nullptr;       // the device this problem lives on
  CompiledModel cm;
  DevModel dev{};                 // device pointers
  std::vector<void *> allocs;     // model arrays on the device (one packed block)
  size_t model_bytes = 0;         // size of that block: the host-to-device bytes of a load
  // search workspace (allocated on first solve)
  int grid = 0, n_warps = 0;
  int32_t *stacks = nullptr;
  size_t stacks_bytes = 0, pool_bytes = 0;
  int ws_lov = -1;
  WarpState *wstate = nullptr;
  unsigned long long *wcount = nullptr;
  unsigned long long *totals = nullptr;
  SearchCtl *ctl = nullptr;
  int32_t *pool_a = nullptr, *pool_b = nullptr;
  int32_t *ready = nullptr;
  int32_t pool_cap = 0;
  int32_t *scratch = nullptr;
  int32_t *solbuf = nullptr;
  int32_t sol_cap = 0;
  std::vector<int32_t> sol_host;
  int32_t n_stored = 0;
  cudaStream_t stream = nullptr;
  NogoodPool ng{};                 // device clause pool of learned nogoods (allocated on demand)
  int32_t *sample_rec = nullptr, *sample_n = nullptr;   // parity instrumentation (csolve_solve_options.sample_mod)
  int32_t sample_cap = 0;
  int32_t sample_seen = 0;
  csolve_solution_fn sink = nullptr;    // every solution is handed to the host between slices (csolve_gpu_set_solution_sink)
  void *sink_user = nullptr;
  csolve_exchange_fn exchange = nullptr;
  void *exchange_user = nullptr;
  csolve_rebalance_fn rebalance = nullptr;
  void *rebalance_user = nullptr;
  const SearchArgs *parked = nullptr;   // search arguments while the rebalance callback runs (export / import are valid)
  bool order_dirty = false;             // a restart rewrote the device copy of the static order (restored by the next solve)

  ~csolve_gpu_problem() {
    DeviceCtx *C = ctx;
    if (C == nullptr) return;
    cudaSetDevice(C->device);
    for (void *p : allocs) C->cfree(p);
    C->release(stacks, stacks_bytes); C->cfree(wstate); C->cfree(wcount); C->cfree(totals); C->cfree(ctl);
    C->release(pool_a, pool_bytes); C->release(pool_b, pool_bytes); C->cfree(scratch); C->cfree(solbuf); C->cfree(ready);
    C->cfree(sample_rec); C->cfree(sample_n);
    C->cfree(ng.lits); C->cfree(ng.start); C->cfree(ng.len); C->cfree(ng.watch); C->cfree(ng.watch_n); C->cfree(ng.counters);
    if (stream) cudaStreamDestroy(stream);
  }
};

// ---- ranks that search one tree together (include/csolve_b200.h: csolve_gpu_comm) -----------------------------------
// Segment of one rank (plain cudaMalloc, so that it can be exported with CUDA IPC):
//   [0, 1024)      the rank's SearchCtl (rank 0's init_next hands out the shared root frontier)
//   [1024, 1152)   CommBlock: written by the peers with system-scope atomics / small copies
//   [4 KiB, +4 MiB)  ready flags of the rank's donation ring
//   then ring_bytes  the donation ring itself (frames): peers serve this rank's tickets over NVLink
//   then             rank 0: the expanded root frontier of the current epoch
static const size_t SEG_COPY = 2048;     // [2048, 2176): where k_comm_wait_front leaves its copy of rank 0's block
static const size_t SEG_CTL = 0, SEG_COMM = 1024, SEG_READY = 4096, SEG_READY_BYTES = (size_t)4 << 20, SEG_RING = SEG_READY + SEG_READY_BYTES;
static const size_t SEG_RING_BYTES_DEFAULT = (size_t)96 << 20;
struct csolve_gpu_comm {
  DeviceCtx *ctx = nullptr;
  int rank = 0, world = 1;
  unsigned char *seg = nullptr;
  size_t front_bytes = 0, ring_bytes = SEG_RING_BYTES_DEFAULT;
  unsigned char *peer_seg[COMM_MAX_RANKS] = {};
  bool opened[COMM_MAX_RANKS] = {};     // mapped with cudaIpcOpenMemHandle (closed in destroy)
  bool connected = false;
  int epoch = 0;                        // solve counter: every rank calls csolve_gpu_solve_comm the same number of times
  // Host-side looks at / small writes into the blocks go through a stream of this rank's own and pinned memory.
  // cudaMemcpy (the synchronous call) on a peer's memory was measured to wait for the peer's running kernel.
  cudaStream_t aux = nullptr;
  CommBlock *hblk = nullptr;            // pinned: [0] block read, [1] staging of a write
  CommBlock *copy_block() const { return reinterpret_cast<CommBlock *>(seg + SEG_COPY); }
  bool read_block(const CommBlock *dev_block, CommBlock *out) {
    if (cudaMemcpyAsync(&hblk[0], dev_block, sizeof(CommBlock), cudaMemcpyDefault, aux) != cudaSuccess) return false;
    if (cudaStreamSynchronize(aux) != cudaSuccess) return false;
    *out = hblk[0];
    return true;
  }
  bool write_words(void *dev_dst, const void *src, size_t bytes) {
    memcpy(&hblk[1], src, bytes);
    if (cudaMemcpyAsync(dev_dst, &hblk[1], bytes, cudaMemcpyDefault, aux) != cudaSuccess) return false;
    return cudaStreamSynchronize(aux) == cudaSuccess;
  }
  SearchCtl *ctl() const { return reinterpret_cast<SearchCtl *>(seg + SEG_CTL); }
  CommBlock *block(int r) const { return reinterpret_cast<CommBlock *>(peer_seg[r] + SEG_COMM); }
  SearchCtl *ctl_of(int r) const { return reinterpret_cast<SearchCtl *>(peer_seg[r] + SEG_CTL); }
  int32_t *front_of(int r) const { return reinterpret_cast<int32_t *>(peer_seg[r] + SEG_RING + ring_bytes); }
  int32_t *ring_of(int r) const { return reinterpret_cast<int32_t *>(peer_seg[r] + SEG_RING); }
  int32_t *ready_of(int r) const { return reinterpret_cast<int32_t *>(peer_seg[r] + SEG_READY); }
  int ring_frames(int fw) const { return (int)std::min<size_t>(ring_bytes / ((size_t)fw * sizeof(int32_t)), SEG_READY_BYTES / sizeof(int32_t)); }
};

extern "C" const char *csolve_last_error(void) { return csolve_front::last_error(); }
extern "C" int csolve_abi_version(void) { return CSOLVE_B200_ABI_VERSION; }

extern "C" int csolve_gpu_init(const csolve_gpu_config *cfg) {
  DeviceCtx *C = nullptr;
  const int rc = device_ctx(cfg ? cfg->device : 0, &C);
  if (rc != CSOLVE_OK) return rc;
  g_default_device.store(C->device);
  return CSOLVE_OK;
}

extern "C" void csolve_gpu_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  for (auto &c : g_ctx) {
    if (c && cudaSetDevice(c->device) == cudaSuccess) c->release_all();
  }
  g_default_device.store(-1);
}

extern "C" int csolve_gpu_load_device(const csolve_flat_model *m, int32_t device, csolve_gpu_problem **out) {
  if (m == nullptr || out == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  *out = nullptr;
  DeviceCtx *C = nullptr;
  int rc = device_ctx(device, &C);
  if (rc != CSOLVE_OK) return rc;
  std::unique_ptr<csolve_gpu_problem> p(new csolve_gpu_problem);
  p->ctx = C;
  std::string err;
  rc = compile_model(*m, p->cm, err);
  if (rc != CSOLVE_OK) return fail(rc, err);

  // ONE upload for the whole compiled model: the arrays are packed into a single pinned-size host image (each
  // 256-byte aligned) and copied with one cudaMemcpy; 18 separate synchronous copies cost ~0.2 ms of a 50 ms search.
  DevModel &d = p->dev;
  d = p->cm.host;
  if (getenv("CSOLVE_NO_LOV")) { d.lov = 0; d.lovk = 0; d.frame_words = frame_words(d.n_vars, d.mask_words); }   // development switch: general kernels only
  if (getenv("CSOLVE_NO_SAT")) d.sat = 0;
  if (getenv("CSOLVE_NO_ADJ")) d.lov_adj_only = 0;
  if (getenv("CSOLVE_NO_DENSE")) d.dense = 0;
  std::vector<unsigned char> image;
  struct Part { size_t off; const void *src; size_t bytes; const void **field; };
  std::vector<Part> parts;
  auto add = [&](const void *src, size_t bytes, const void **field) {
    const size_t off = (image.size() + 255) & ~(size_t)255;
    image.resize(off + std::max<size_t>(bytes, 16));
    if (bytes) memcpy(image.data() + off, src, bytes);
    parts.push_back(Part{off, src, bytes, field});
  };
#define UP(field, vec) add(p->cm.vec.data(), p->cm.vec.size() * sizeof(p->cm.vec[0]), (const void **)&d.field)
  UP(clause, clause); UP(watch_ptr, watch_ptr); UP(watch_idx, watch_idx); UP(wrec, wrec); UP(wrec_ptr, wrec_ptr);
  UP(lov_pair, lov_pair); UP(lov_cptr, lov_cptr); UP(lov_cval, lov_cval); UP(lov_fconst, lov_fconst); UP(lov_adj, lov_adj);
  UP(lin, lin); UP(lin_term, lin_term); UP(linrel, linrel); UP(dense_form, dense_form); UP(sat_occ_ptr, sat_occ_ptr); UP(sat_occ, sat_occ);
  UP(node_op, node_op); UP(node_l, node_l); UP(node_r, node_r); UP(node_first, node_first);
  UP(order, order); UP(prio, prio); UP(root_dom, root_dom);
#undef UP
  unsigned char *base = nullptr;
  CUDA_TRY(cmalloc(C, &base, image.size()));
  p->allocs.push_back(base);
  CUDA_TRY(cudaMemcpy(base, image.data(), image.size(), cudaMemcpyHostToDevice));
  for (const Part &q : parts) *q.field = base + q.off;
  p->model_bytes = image.size();
  CUDA_TRY(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  *out = p.release();
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_load(const csolve_flat_model *m, csolve_gpu_problem **out) {
  int dev = g_default_device.load();
  if (dev < 0) {
    int rc = csolve_gpu_init(nullptr);
    if (rc != CSOLVE_OK) return rc;
    dev = g_default_device.load();
  }
  return csolve_gpu_load_device(m, dev, out);
}

extern "C" void csolve_gpu_unload(csolve_gpu_problem *p) { delete p; }

namespace {
// caller-supplied domains must be sub-intervals of the model's root domains (see include/csolve_b200.h)
int check_domains(const csolve_gpu_problem *p, int32_t n, const int32_t *dom, const char *what) {
  const int V = p->dev.n_vars;
  const int32_t *root = p->cm.root_dom.data();
  for (int32_t b = 0; b < n; b++) {
    const int32_t *d = dom + (size_t)b * 2 * V;
    for (int v = 0; v < V; v++) {
      if (d[2 * v] > d[2 * v + 1] || d[2 * v] < root[2 * v] || d[2 * v + 1] > root[2 * v + 1]) {
        return fail(CSOLVE_ERR_INVALID, std::string(what) + " " + std::to_string(b) + ": the domain of variable " + std::to_string(v) +
                                            " is empty or not inside the model's root domain");
      }
    }
  }
  return CSOLVE_OK;
}
}  // namespace

extern "C" int csolve_gpu_propagate_batch(csolve_gpu_problem *p, int32_t n_nodes, const int32_t *dom_in,
                                          const int32_t *var, const int32_t *val, const int32_t *best,
                                          int32_t *dom_out, uint8_t *failed) {
  if (p == nullptr || n_nodes < 0) return fail(CSOLVE_ERR_INVALID, "bad arguments");
  if (n_nodes == 0) return CSOLVE_OK;
  ENTER(p);
  const int V = p->dev.n_vars;
  if (dom_in == nullptr || var == nullptr || val == nullptr || dom_out == nullptr || failed == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  for (int b = 0; b < n_nodes; b++) {
    if (var[b] < 0 || var[b] >= V) return fail(CSOLVE_ERR_INVALID, "decision variable out of range");
    if (val[b] < p->cm.root_dom[2 * var[b]] || val[b] > p->cm.root_dom[2 * var[b] + 1]) return fail(CSOLVE_ERR_INVALID, "decision value outside the variable's root domain");
  }
  {
    const int rc0 = check_domains(p, n_nodes, dom_in, "node");
    if (rc0 != CSOLVE_OK) return rc0;
  }
  const size_t dom_bytes = (size_t)n_nodes * 2 * V * sizeof(int32_t), vec_bytes = (size_t)n_nodes * sizeof(int32_t);
  int32_t *d_in = nullptr, *d_var = nullptr, *d_val = nullptr, *d_best = nullptr, *d_out = nullptr;
  uint8_t *d_failed = nullptr;
  std::vector<int32_t> zero_best;
  if (best == nullptr) { zero_best.assign(n_nodes, 0); best = zero_best.data(); }
  int rc = CSOLVE_OK;
  auto cleanup = [&]() { C->cfree(d_in); C->cfree(d_var); C->cfree(d_val); C->cfree(d_best); C->cfree(d_out); C->cfree(d_failed); };
#define TRY2(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(CSOLVE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
  TRY2(cmalloc(C, &d_in, dom_bytes)); TRY2(cmalloc(C, &d_out, dom_bytes));
  TRY2(cmalloc(C, &d_var, vec_bytes)); TRY2(cmalloc(C, &d_val, vec_bytes)); TRY2(cmalloc(C, &d_best, vec_bytes));
  TRY2(cmalloc(C, &d_failed, n_nodes));
  TRY2(cudaMemcpyAsync(d_in, dom_in, dom_bytes, cudaMemcpyHostToDevice, p->stream));
  TRY2(cudaMemcpyAsync(d_var, var, vec_bytes, cudaMemcpyHostToDevice, p->stream));
  TRY2(cudaMemcpyAsync(d_val, val, vec_bytes, cudaMemcpyHostToDevice, p->stream));
  TRY2(cudaMemcpyAsync(d_best, best, vec_bytes, cudaMemcpyHostToDevice, p->stream));
  int grid = std::min((n_nodes + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK, std::max(1, C->sm_count) * 8);
  TRY2(launch_propagate_batch(p->dev, n_nodes, d_in, d_var, d_val, d_best, d_out, d_failed, grid, p->stream));
  TRY2(cudaMemcpyAsync(dom_out, d_out, dom_bytes, cudaMemcpyDeviceToHost, p->stream));
  TRY2(cudaMemcpyAsync(failed, d_failed, n_nodes, cudaMemcpyDeviceToHost, p->stream));
  TRY2(cudaStreamSynchronize(p->stream));
#undef TRY2
  cleanup();
  return rc;
}

// ---- search ------------------------------------------------------------------------------------
namespace {

// slots behind the expanded root frontier that stay free for donated frames (a ticket queue: one slot per waiting
// warp, plus frames delivered ahead of their tickets by racing donors -- at most one per warp -- plus one import of at
// most n_warps frames from another rank)
int ring_min_frames(int n_warps) { return 4 * n_warps + 1024; }

// solutions a search can add to the buffer after a warp has noticed that it is nearly full: every warp runs on to
// its next poll of the control block (at most 32 nodes, each of which may be an accepted leaf), twice over
int sink_headroom(int n_warps) { return 64 * n_warps + 4096; }

// Breadth-first expansion target. The ticket queue keeps every warp busy to the end whatever the size of the root
// frontier (measured: 16-queens search time is the same from 16 to 256 frames per warp), so the frontier only has to
// be long enough for an even rank partition; every further level is expansion time that all ranks replicate.
const int DEFAULT_FRAMES_PER_WARP = 16;

int ensure_workspace(csolve_gpu_problem *p, const csolve_solve_options &opt, bool batch, int n_roots, bool learn, bool backjump) {
  DeviceCtx *C = p->ctx;
  DevModel m = p->dev;
  if (batch) m.lov = 0;
  const bool sample = opt.sample_mod != 0u;
  const bool sat = !batch && search_uses_sat(m, learn, opt.order);
  const int ws_key = m.lov * 16 + m.lovk + (learn ? 64 : 0) + (sample ? 128 : 0) + (sat ? 256 : 0) + (backjump ? 512 : 0);
  if (p->stacks != nullptr && p->ws_lov != ws_key) {
    // the lane-owns-variable and the general kernels have different occupancies: rebuild the per-warp state
    C->release(p->stacks, p->stacks_bytes); C->cfree(p->wstate); C->cfree(p->wcount); C->cfree(p->totals); C->cfree(p->ctl); C->cfree(p->scratch);
    p->stacks = nullptr; p->wstate = nullptr; p->wcount = nullptr; p->totals = nullptr; p->ctl = nullptr; p->scratch = nullptr;
  }
  if (p->stacks == nullptr) {
    p->ws_lov = ws_key;
    int per_sm = search_blocks_per_sm(m, false, learn, sample, sat, backjump);
    if (per_sm <= 0) return fail(CSOLVE_ERR_CUDA, "search kernel does not fit on the device (shared memory per node too large)");
    p->grid = per_sm * C->sm_count;
    p->n_warps = p->grid * WARPS_PER_BLOCK;
    const size_t stack_words = (size_t)p->n_warps * (m.n_vars + 1) * m.frame_words;
    p->stacks_bytes = stack_words * sizeof(int32_t);
    CUDA_TRY(C->alloc((void **)&p->stacks, p->stacks_bytes, nullptr));
    CUDA_TRY(cmalloc(C, &p->wstate, (size_t)p->n_warps * sizeof(WarpState)));
    CUDA_TRY(cmalloc(C, &p->wcount, (size_t)p->n_warps * CNT_WIDTH * sizeof(unsigned long long)));
    CUDA_TRY(cmalloc(C, &p->totals, CNT_WIDTH * sizeof(unsigned long long)));
    CUDA_TRY(cmalloc(C, &p->ctl, sizeof(SearchCtl)));
    CUDA_TRY(cmalloc(C, &p->scratch, (size_t)(4 + 3 * p->n_warps) * sizeof(int32_t)));
  }
  // frontier pools: room for the split target times the largest branching the expansion may apply
  int target = opt.split_target > 0 ? opt.split_target : p->n_warps * DEFAULT_FRAMES_PER_WARP;
  target = std::max(target, n_roots);
  int cap = std::max(target * 4, 1 << 16) + ring_min_frames(p->n_warps);
  const size_t max_bytes = (size_t)2 << 30;   // per pool
  while ((size_t)cap * m.frame_words * sizeof(int32_t) > max_bytes && cap > 1024) cap /= 2;
  if (cap > p->pool_cap) {
    C->release(p->pool_a, p->pool_bytes); C->release(p->pool_b, p->pool_bytes); p->pool_a = p->pool_b = nullptr;
    p->pool_bytes = (size_t)cap * m.frame_words * sizeof(int32_t);
    CUDA_TRY(C->alloc((void **)&p->pool_a, p->pool_bytes, nullptr));
    CUDA_TRY(C->alloc((void **)&p->pool_b, p->pool_bytes, nullptr));
    C->cfree(p->ready); p->ready = nullptr;
    CUDA_TRY(cmalloc(C, &p->ready, (size_t)cap * sizeof(int32_t)));
    p->pool_cap = cap;
  }
  if (sample) {
    const int want = opt.sample_cap > 0 ? opt.sample_cap : 65536;
    if (want > p->sample_cap) {
      C->cfree(p->sample_rec); p->sample_rec = nullptr;
      CUDA_TRY(cmalloc(C, &p->sample_rec, (size_t)want * sample_words(m.n_vars) * sizeof(int32_t)));
      p->sample_cap = want;
    }
    if (p->sample_n == nullptr) CUDA_TRY(cmalloc(C, &p->sample_n, sizeof(int32_t)));
  }
  int sol_cap = std::max(opt.max_solutions, m.obj_var >= 0 ? 4096 : (m.objective == CSOLVE_OBJ_ANY ? 1 : 0));
  if (p->sink != nullptr && m.objective == CSOLVE_OBJ_ALL) sol_cap = std::max(sol_cap, std::max(1 << 20, 4 * sink_headroom(p->n_warps)));
  if (sol_cap > p->sol_cap) {
    C->cfree(p->solbuf); p->solbuf = nullptr;
    CUDA_TRY(cmalloc(C, &p->solbuf, (size_t)sol_cap * (m.n_vars + 1) * sizeof(int32_t)));
    p->sol_cap = sol_cap;
  }
  return CSOLVE_OK;
}

}  // namespace

namespace {
// n_roots == 0: the model's own root. n_roots > 0: batched roots over the shared network (ALL models).
// host-side wait on a CommBlock in THIS rank's device memory: polled with small copies on the comm's own stream;
// `ok(block)` decides. Returns false after `limit_s` seconds. (Waits on a peer's block run on the device:
// k_comm_wait_front.)
template <class F>
bool comm_wait(csolve_gpu_comm *c, const CommBlock *dev_block, F ok, double limit_s, CommBlock *out) {
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spin = 0;; spin++) {
    CommBlock b;
    if (!c->read_block(dev_block, &b)) return false;
    if (ok(b)) {
      // the fields of a block are written by separate stores: read once more so that everything that was written
      // before the field `ok` looked at is seen as well
      if (!c->read_block(dev_block, &b)) return false;
      if (out) *out = b;
      return true;
    }
    if ((spin & 63u) == 63u && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) return false;
    if (spin > 2000) std::this_thread::sleep_for(std::chrono::microseconds(50));
  }
}
const double COMM_WAIT_S = 120.0;

// CSOLVE_DEBUG_COMM: host-side timeline of a rank (microseconds since the first call in this process)
double comm_now_us() {
  static const auto t0 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
}
#define COMM_TRACE(c, what)                                                                              \
  do {                                                                                                   \
    if ((c) != nullptr && g_comm_trace) fprintf(stderr, "[csolve comm] rank %d %10.1f us  %s\n", (c)->rank, comm_now_us(), what); \
  } while (0)
const bool g_comm_trace = getenv("CSOLVE_DEBUG_COMM") != nullptr;

int solve_impl(csolve_gpu_problem *p, csolve_gpu_comm *c, const csolve_solve_options *opt_in, csolve_gpu_result *res, int n_roots,
               const int32_t *root_dom, uint32_t *root_solutions, uint8_t *root_failed) {
  if (p == nullptr || res == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  ENTER(p);
  if (c != nullptr && (c->ctx != p->ctx || !c->connected)) return fail(CSOLVE_ERR_INVALID, "the comm is not connected or lives on another device than the problem");
  if (c != nullptr && n_roots > 0) return fail(CSOLVE_ERR_UNSUPPORTED, "batched roots are sharded by the caller (one slice of the roots per rank), not through a comm");
  if (c != nullptr && c->world == 1) c = nullptr;
  COMM_TRACE(c, "enter");
  csolve_solve_options opt;
  memset(&opt, 0, sizeof(opt));
  if (opt_in) opt = *opt_in;
  if (opt.part_count <= 0) opt.part_count = 1;
  if (c != nullptr) { opt.part_rank = 0; opt.part_count = 1; }       // the comm decides (shared frontier, or rank / world as a fallback)
  if (opt.part_rank < 0 || opt.part_rank >= opt.part_count) return fail(CSOLVE_ERR_INVALID, "part_rank out of range");
  if (opt.order < CSOLVE_ORDER_NONE || opt.order > CSOLVE_ORDER_LARGEST_VALUE) return fail(CSOLVE_ERR_INVALID, "invalid ordering strategy");
  memset(res, 0, sizeof(*res));
  const bool batch = n_roots > 0;
  if (batch && p->dev.objective != CSOLVE_OBJ_ALL) return fail(CSOLVE_ERR_UNSUPPORTED, "batched roots need an ALL model");
  // learning needs the general kernel; the specialised NOT(EQ) kernels never meet a 0/1-only conflict
  const bool learn = opt.create_conflicts != 0 && !batch && !p->dev.lov && !p->dev.lovk;
  if (learn && opt.sample_mod != 0u) return fail(CSOLVE_ERR_UNSUPPORTED, "sample_mod is not available together with create_conflicts");
  if (batch) {
    const int rc0 = check_domains(p, n_roots, root_dom, "root");
    if (rc0 != CSOLVE_OK) return rc0;
  }
  // back-jumping (src/csolve.c:350-364) re-decides frames: never for ALL models, whose counts must stay the tree's
  const bool backjump = learn && opt.backjump != 0 && p->dev.objective != CSOLVE_OBJ_ALL;
  int rc = ensure_workspace(p, opt, batch, n_roots, learn, backjump);
  if (rc != CSOLVE_OK) return rc;
  if (learn) {
    NogoodPool &g = p->ng;
    if (g.lits == nullptr) {
      g.cap_ng = 1 << 18; g.cap_lits = 1 << 23; g.cap_w = 4096;
      CUDA_TRY(cmalloc(C, &g.lits, (size_t)g.cap_lits * 4)); CUDA_TRY(cmalloc(C, &g.start, (size_t)g.cap_ng * 4));
      CUDA_TRY(cmalloc(C, &g.len, (size_t)g.cap_ng * 4)); CUDA_TRY(cmalloc(C, &g.watch, (size_t)p->dev.n_vars * g.cap_w * 4));
      CUDA_TRY(cmalloc(C, &g.watch_n, (size_t)p->dev.n_vars * 4)); CUDA_TRY(cmalloc(C, &g.counters, 8 * 4));
    }
    CUDA_TRY(cudaMemsetAsync(g.watch, 0xff, (size_t)p->dev.n_vars * g.cap_w * 4, p->stream));
    CUDA_TRY(cudaMemsetAsync(g.watch_n, 0, (size_t)p->dev.n_vars * 4, p->stream));
    CUDA_TRY(cudaMemsetAsync(g.counters, 0, 8 * 4, p->stream));
  }

  DevModel m = p->dev;
  if (batch) m.lov = 0;            // batched roots run on the general kernels
  SearchCtl *const dctl = c != nullptr ? c->ctl() : p->ctl;      // with a comm the control block lives in the rank's segment
  // ---- comm: epoch bookkeeping. Whatever way this call ends, the other ranks must not wait for ever: rank 0
  //      publishes "failed" if it never published a frontier, every other rank reports that it has left the epoch.
  const int epoch = c != nullptr ? ++c->epoch : 0;
  bool front_published = false;
  const SearchArgs *comm_args = nullptr;       // set once the search arguments are complete
  int32_t *d_cstate = nullptr;                 // comm: result of k_comm_state
  ScopeExit comm_guard([&]() {
    if (c == nullptr) return;
    // leave the epoch: this rank no longer counts as active, whatever state its search is in (errors, time limit)
    if (comm_args != nullptr && d_cstate != nullptr) {
      launch_comm_state(*comm_args, 0, 1, d_cstate, p->stream);
      cudaStreamSynchronize(p->stream);
    }
    C->cfree(d_cstate);
    if (c->rank == 0) {
      if (!front_published) {
        const int32_t hdr[3] = {epoch, -2, 0};
        c->write_words(&c->block(0)->front_n, &hdr[1], 2 * sizeof(int32_t));
        c->write_words(&c->block(0)->front_epoch, &hdr[0], sizeof(int32_t));
      }
    } else {
      c->write_words(&c->block(0)->done_epoch[c->rank], &epoch, sizeof(int32_t));
    }
  });
  const CompiledModel &cm = p->cm;
  const int V = m.n_vars, fw = m.frame_words;
  cudaStream_t st = p->stream;
  const auto wall0 = std::chrono::steady_clock::now();
  if (p->order_dirty) {
    CUDA_TRY(cudaMemcpyAsync(const_cast<int32_t *>(m.order), cm.order.data(), (size_t)m.n_vars * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    p->order_dirty = false;
  }

  // ---- initial state ----------------------------------------------------------------------------
  SearchCtl ctl;
  memset(&ctl, 0, sizeof(ctl));
  ctl.best = m.objective == CSOLVE_OBJ_MIN ? INT32_MAX : (m.objective == CSOLVE_OBJ_MAX ? INT32_MIN : 0);   // src/objective.c:38-50
  CUDA_TRY(cudaMemsetAsync(p->wcount, 0, (size_t)p->n_warps * CNT_WIDTH * sizeof(unsigned long long), st));
  std::vector<WarpState> ws(p->n_warps, WarpState{-1, 0, 0, 0u});
  CUDA_TRY(cudaMemcpyAsync(p->wstate, ws.data(), ws.size() * sizeof(WarpState), cudaMemcpyHostToDevice, st));

  int32_t *d_roots = nullptr; unsigned char *d_rfail = nullptr; unsigned int *d_rsol = nullptr; int32_t *d_nout = nullptr;
  ScopeExit batch_guard([&]() { C->cfree(d_roots); C->cfree(d_rfail); C->cfree(d_rsol); C->cfree(d_nout); });
  // the root frame (root_var < 0: the configured order's first variable)
  auto upload_root_frame = [&](int root_var) -> int {
    std::vector<int32_t> root(fw, 0);
    const int rv = root_var >= 0 ? root_var : select_root_var(cm, opt.order);
    root[FR_VAR] = rv; root[FR_ITER] = 0;
    root[FR_LO] = cm.root_dom[2 * rv]; root[FR_HI] = cm.root_dom[2 * rv + 1];
    root[FR_LAST] = (int32_t)((uint32_t)root[FR_HI] - (uint32_t)root[FR_LO]);
    root[FR_LEVEL] = 0; root[FR_BEST] = ctl.best; root[7] = 0x1234567;
    memcpy(&root[frame_dom_offset(m.mask_words)], cm.root_dom.data(), sizeof(int32_t) * 2 * V);
    if (m.lovk) memcpy(&root[frame_dom_offset(m.mask_words) + 2 * V], cm.lov_fconst.data(), sizeof(int32_t) * V);   // value sets
    CUDA_TRY(cudaMemcpyAsync(p->pool_a, root.data(), fw * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));            // `root` goes out of scope
    ctl.item_count = 1;
    return CSOLVE_OK;
  };
  if (!batch) {
    rc = upload_root_frame(-1);
    if (rc != CSOLVE_OK) return rc;
  } else {
    // root phase on the device: propagate every root to fixpoint, emit one tagged frame per consistent root
    const size_t rb = (size_t)n_roots * 2 * V * sizeof(int32_t);
    CUDA_TRY(cmalloc(C, &d_roots, rb)); CUDA_TRY(cmalloc(C, &d_rfail, n_roots));
    CUDA_TRY(cmalloc(C, &d_rsol, (size_t)n_roots * sizeof(unsigned int))); CUDA_TRY(cmalloc(C, &d_nout, sizeof(int32_t)));
    CUDA_TRY(cudaMemcpyAsync(d_roots, root_dom, rb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_rsol, 0, (size_t)n_roots * sizeof(unsigned int), st));
    CUDA_TRY(cudaMemsetAsync(d_nout, 0, sizeof(int32_t), st));
    // one frame per consistent root goes into pool_a, the donation ring needs its slots behind them
    if ((long long)n_roots + ring_min_frames(p->n_warps) > p->pool_cap) {
      return fail(CSOLVE_ERR_CAPACITY, "too many roots for one call: " + std::to_string(n_roots) + " roots, the frontier pool holds " +
                                           std::to_string(p->pool_cap - ring_min_frames(p->n_warps)) + " (split the batch)");
    }
    const int grid = std::min(p->grid, (n_roots + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    CUDA_TRY(launch_root_frames(m, n_roots, d_roots, opt.order, p->pool_a, p->pool_cap, d_nout, d_rfail, grid, st));
    int32_t n_ok = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_ok, d_nout, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ctl.item_count = n_ok;
  }
  CUDA_TRY(cudaMemcpyAsync(dctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, st));

  SearchArgs a;
  memset(&a, 0, sizeof(a));
  a.m = m; a.ctl = dctl; a.stacks = p->stacks; a.wstate = p->wstate; a.wcount = p->wcount;
  a.solbuf = p->solbuf; a.max_solutions = p->sol_cap; a.n_warps = p->n_warps; a.order = opt.order;
  a.out_cap = p->pool_cap; a.expand_branch_max = 64;
  a.inst_solutions = d_rsol;
  a.use_sat = !batch && search_uses_sat(m, learn, opt.order) ? 1 : 0;
  int part_rank = opt.part_rank, part_count = opt.part_count;
  const bool sinking = p->sink != nullptr && m.objective == CSOLVE_OBJ_ALL && !batch;
  if (sinking) a.sink_headroom = sink_headroom(p->n_warps);
  uint64_t sunk = 0, n_sliced = 0;      // n_sliced: depth-first slices run so far (0 = still in the expansion)
  // hands what the solution buffer holds to the sink and empties it (`ctl` was just read back from the device)
  auto drain_solutions = [&]() -> int {
    if (!sinking || ctl.n_stored <= 0) return CSOLVE_OK;
    if (ctl.n_stored > p->sol_cap) return fail(CSOLVE_ERR_CAPACITY, "solution buffer overflow: " + std::to_string(ctl.n_stored) + " solutions in one slice, room for " + std::to_string(p->sol_cap));
    p->sol_host.resize((size_t)ctl.n_stored * (V + 1));
    CUDA_TRY(cudaMemcpyAsync(p->sol_host.data(), p->solbuf, p->sol_host.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    const int32_t zero = 0;
    CUDA_TRY(cudaMemcpyAsync(&dctl->n_stored, &zero, sizeof(zero), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // (a replicated expansion is reported by rank 0 only, like its counters)
    if (!(part_count > 1 && part_rank != 0 && n_sliced == 0)) { p->sink(p->sink_user, p->sol_host.data(), ctl.n_stored, V + 1); sunk += (uint64_t)ctl.n_stored; }
    ctl.n_stored = 0;
    return CSOLVE_OK;
  };
  p->sample_seen = 0;
  if (opt.sample_mod != 0u) {
    a.sample_rec = p->sample_rec; a.sample_n = p->sample_n; a.sample_cap = p->sample_cap; a.sample_mod = opt.sample_mod;
    a.sample_fkeep = std::max(opt.sample_failed_keep, 1u);
    CUDA_TRY(cudaMemsetAsync(p->sample_n, 0, sizeof(int32_t), st));
  }
  int32_t *d_gprio = nullptr;
  ScopeExit gprio_guard([&]() { C->cfree(d_gprio); });
  if (opt.prefer_failing && !m.lov) {
    // device-wide dynamic priorities, seeded with the parse-time weights (env_t.prio)
    CUDA_TRY(cmalloc(C, &d_gprio, (size_t)V * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpyAsync(d_gprio, cm.prio.data(), (size_t)V * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  // a slice ends so that the host can check the time limit / run the rank exchange; without either the kernel only
  // has to come back when it is done. Branch-and-bound profits from short slices: between two slices k_rebalance
  // hands EVERY idle warp half of a busy warp's shallowest frame at once, which finds good incumbents sooner
  // (wcet: 11 ms with 1-2 ms slices, 15 ms with 5 ms, 23-33 ms with 10 ms, several times that with one long slice).
  const int slice_ms = opt.slice_ms > 0 ? opt.slice_ms
                       : p->dev.obj_var >= 0 ? 2
                       : (p->exchange != nullptr || opt.time_limit_ms > 0) ? 20 : 1000;
  a.slice_cycles = (long long)C->clock_khz * slice_ms;

  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  ScopeExit event_guard([&]() { if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); if (ev2) cudaEventDestroy(ev2); });
  CUDA_TRY(cudaEventCreate(&ev0)); CUDA_TRY(cudaEventCreate(&ev1)); CUDA_TRY(cudaEventCreate(&ev2));
  CUDA_TRY(cudaEventRecord(ev0, st));

  // ---- batched frontier expansion -------------------------------------------------------------------
  const int target = opt.split_target > 0 ? opt.split_target : p->n_warps * DEFAULT_FRAMES_PER_WARP;
  long long max_branch = 1;
  for (int v = 0; v < V; v++) max_branch = std::max<long long>(max_branch, (long long)cm.root_dom[2 * v + 1] - cm.root_dom[2 * v] + 1);
  max_branch = std::min<long long>(max_branch, a.expand_branch_max);
  int32_t *pin = p->pool_a, *pout = p->pool_b;
  int n_items = ctl.item_count;
  uint64_t launches = batch ? 1 : 0;
  if (batch && n_items >= p->n_warps / 2) n_items = -n_items;   // enough roots: no breadth-first phase at all
  bool stopped = false;
  // Breadth-first levels pay off while the frontier multiplies; once there is about one frame per two warps
  // and a level no longer doubles it (unit-propagation-heavy models, batched roots), or after 24 levels,
  // the depth-first phase with rebalancing takes over.
  auto expand_root = [&]() -> int {
    for (int lvl = 0; lvl < V && lvl < 24 && n_items > 0 && n_items < target; ++lvl) {
      const int before = n_items;
      // would another level overflow the pool? domains only shrink, so a frame has at most as many
      // children as the largest root domain (and never more than expand_branch_max)
      if ((long long)n_items * max_branch > p->pool_cap - ring_min_frames(p->n_warps)) break;
      a.items = pin; a.items_out = pout; a.frozen_best = ctl.best;
      ctl.item_next = 0; ctl.item_count = n_items; ctl.out_count = 0; ctl.passed = 0;
      CUDA_TRY(cudaMemcpyAsync(dctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, st));
      const int grid = std::min(p->grid, (n_items + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
      CUDA_TRY(launch_search(a, grid, true, st)); launches++;
      CUDA_TRY(cudaMemcpyAsync(&ctl, dctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      if (ctl.out_dropped > 0) return fail(CSOLVE_ERR_CAPACITY, "frontier pool overflow during expansion");
      { const int rcd = drain_solutions(); if (rcd != CSOLVE_OK) return rcd; }
      n_items = ctl.out_count;
      if (getenv("CSOLVE_DEBUG")) fprintf(stderr, "[csolve] expand level %d -> %d frames (target %d, pool %d, branch %lld)\n", lvl, n_items, target, p->pool_cap, max_branch);
      std::swap(pin, pout);
      if (ctl.signal == SIG_STOP) { stopped = true; break; }
      // frames with huge domains are passed through unsplit; when nothing else is left the
      // breadth-first phase cannot make progress and the depth-first phase (which bisects) takes over
      if (ctl.passed == n_items) break;
      if (n_items >= p->n_warps / 2 && n_items < 2 * (long long)before) break;
    }
    return CSOLVE_OK;
  };
  // Ranks of a comm.
  // ANY / MIN / MAX: rank 0 alone expands the root; the others take the frontier from rank 0's segment and every rank
  //   claims frames of that ONE frontier in its own order (part_rank / part_count = 0 / 1) -- the value order's
  //   preferred sub-trees are searched first, by everybody; incumbents and "found" travel over peer memory, and the
  //   ranks serve each other's donation rings over NVLink (3-SAT n=200 seed 1 on 8 x B200: 7.4 ms, 14.1 ms on one).
  // ALL: every rank expands for itself, keeps the frames whose path hash maps to it (part_rank / part_count = rank /
  //   world) and searches them on its own -- no data-path exchange at all, the results are summed by the caller.
  //   Measured on 8 x B200, 16-queens (one GPU: 51.7 ms):
  //     shared frontier, remote claims                      10.0 ms  (rank 0 143 M nodes, the others 75-110 M + donations)
  //     shared frontier of 4 x / 16 x as many frames          7.9 / 8.5 ms  (balanced, but rank 0's expansion is serial)
  //     own frontiers + rings served across ranks             7.75 ms (hand-offs over NVLink in the end game stall the donors)
  //     own frontiers, nothing shared                         7.14 ms
  //   The path hash deals sub-trees to within 2 %, which no exchange on a 7 ms search can beat.
  // The same partition is used when a shared frontier does not fit rank 0's segment.
  const int32_t *front_pool = nullptr;
  SearchCtl *front_ctl = dctl;
  COMM_TRACE(c, "set up");
  const bool own_front = c != nullptr && m.objective == CSOLVE_OBJ_ALL && getenv("CSOLVE_COMM_SHARED_FRONT") == nullptr;
  // rank 0 opens the epoch: every rank starts it active (the count lives in rank 0's block, CommBlock::active64), then
  // the header the peers wait for. front_n >= 0: that many frames are in rank 0's segment; -1: every rank expands.
  auto publish_epoch = [&](int front_n) -> int {
    // the previous epoch's frontier / counters are overwritten: every peer must have left that search
    if (!comm_wait(c, c->block(0), [&](const CommBlock &b) {
          for (int r = 1; r < c->world; r++) if (b.done_epoch[r] < epoch - 1) return false;
          return true; }, COMM_WAIT_S, nullptr))
      return fail(CSOLVE_ERR_CUDA, "comm: a peer rank did not finish the previous search");
    const int32_t hdr[3] = {epoch, front_n, fw};
    const unsigned long long act = ((unsigned long long)(uint32_t)epoch << 32) | (unsigned)c->world;
    CUDA_TRY(cudaMemcpyAsync(&c->block(0)->active64, &act, sizeof(act), cudaMemcpyHostToDevice, st));
    for (int r = 0; r < c->world; r++) CUDA_TRY(cudaMemcpyAsync(&c->block(r)->busy_epoch, &epoch, sizeof(int32_t), cudaMemcpyDefault, st));
    if (front_n >= 0) {
      // the frontier's claim counter must be zero before anybody sees the frontier
      ctl.init_next = 0;
      CUDA_TRY(cudaMemcpyAsync(&dctl->init_next, &ctl.init_next, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    CUDA_TRY(cudaMemcpyAsync(&c->block(0)->front_n, &hdr[1], 2 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaMemcpyAsync(&c->block(0)->front_epoch, &hdr[0], sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    front_published = true;
    return CSOLVE_OK;
  };
  // the other ranks wait for it on the device (one thread polls rank 0's block over NVLink), then ONE copy of what it saw
  auto await_epoch = [&](CommBlock *b) -> int {
    CUDA_TRY(launch_comm_wait_front(c->block(0), epoch, COMM_WAIT_S, c->copy_block(), st));
    CUDA_TRY(cudaMemcpyAsync(&c->hblk[0], c->copy_block(), sizeof(CommBlock), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *b = c->hblk[0];
    if (b->front_epoch < epoch) return fail(CSOLVE_ERR_CUDA, "comm: rank 0 did not open the search");
    if (b->front_epoch != epoch || b->front_n == -2) return fail(CSOLVE_ERR_CUDA, "comm: rank 0 failed or the ranks are out of step");
    return CSOLVE_OK;
  };
  if (c == nullptr) {
    rc = expand_root();
    if (rc != CSOLVE_OK) return rc;
    if (n_items < 0) n_items = -n_items;
  } else if (own_front) {
    // nothing is shared in this search: no frontier, no ring, no incumbent -- the ranks do not even have to meet
    part_rank = c->rank; part_count = c->world;
    front_published = true;
    rc = expand_root();
    if (rc != CSOLVE_OK) return rc;
    if (n_items < 0) n_items = -n_items;
    COMM_TRACE(c, "expanded");
  } else if (c->rank == 0) {
    rc = expand_root();
    if (rc != CSOLVE_OK) return rc;
    if (n_items < 0) n_items = -n_items;
    COMM_TRACE(c, "expanded");
    const size_t bytes = (size_t)n_items * fw * sizeof(int32_t);
    int front_n = stopped ? 0 : n_items;           // ANY solved by the expansion itself: nothing to share
    if (bytes <= c->front_bytes) {
      if (bytes) CUDA_TRY(cudaMemcpyAsync(c->front_of(0), pin, bytes, cudaMemcpyDeviceToDevice, st));
      front_pool = c->front_of(0);
    } else {
      front_n = -1;                                // does not fit: every rank expands for itself
      part_rank = 0; part_count = c->world;
    }
    rc = publish_epoch(front_n);
    if (rc != CSOLVE_OK) return rc;
  } else {
    CommBlock b;
    rc = await_epoch(&b);
    if (rc != CSOLVE_OK) return rc;
    if (b.front_n >= 0) {
      if (b.front_fw != fw) return fail(CSOLVE_ERR_INVALID, "comm: the ranks loaded different models");
      n_items = b.front_n;
      front_pool = c->front_of(0);
      front_ctl = c->ctl_of(0);
    } else {
      part_rank = c->rank; part_count = c->world;
      rc = expand_root();
      if (rc != CSOLVE_OK) return rc;
      if (n_items < 0) n_items = -n_items;
    }
  }
  // ---- partition: every rank holds the whole frontier; the search kernel skips the frames whose path
  //      hash maps to another rank (no copy, no compaction)
  a.part_rank = part_rank; a.part_count = part_count;

  if (part_count > 1 && part_rank != 0 && front_pool == nullptr) {
    // the expansion was replicated on every rank (also when it exhausted the whole tree):
    // only rank 0 reports its counters and the leaves found in it
    CUDA_TRY(cudaMemsetAsync(p->wcount, 0, (size_t)p->n_warps * CNT_WIDTH * sizeof(unsigned long long), st));
    if (m.objective == CSOLVE_OBJ_ALL) ctl.n_stored = 0;
  }
  COMM_TRACE(c, "frontier published / seen");
  CUDA_TRY(cudaEventRecord(ev1, st));

  // ---- time-sliced persistent search -----------------------------------------------------------------------
  a.items = pin; a.items_out = nullptr;
  // shared pool ring of the depth-first phase: the expanded frontier occupies the first n_items slots
  a.pool = pin; a.pool_cap = p->pool_cap; a.ready = p->ready; a.n_initial = n_items;
  a.front_pool = front_pool != nullptr ? front_pool : pin;
  a.front_ctl = front_ctl;
  a.total_warps = p->n_warps * (front_pool != nullptr ? c->world : 1);
  const bool cross = c != nullptr && !own_front;     // the ranks serve each other's donation rings
  if (c != nullptr && !own_front) {
    a.comm = c->block(c->rank); a.epoch = epoch; a.rank = c->rank; a.world = c->world; a.n_peers = c->world - 1;
    for (int r = 0; r < c->world; r++) a.peer_comm[r] = c->block(r);
  }
  // Branch-and-bound: a rank that ran dry does not ask its peers for frames during the first COMM_BNB_GRACE_MS of a
  // search. Sub-trees handed to another GPU are searched against an incumbent that arrives late, and on a tree as
  // narrow as wcet's (330 k nodes, 6 ms on one GPU) that is all the extra GPUs ever do: 2 GPUs searched 2.4 x the
  // nodes in 10 ms. A search that lasts longer than the grace period is worth sharing.
  const double COMM_BNB_GRACE_MS = 20.0;
  auto demand_allowed = [&]() {
    if (!cross || getenv("CSOLVE_COMM_NO_DEMAND") != nullptr) return false;
    if (m.obj_var < 0) return true;
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count() >= COMM_BNB_GRACE_MS;
  };
  a.peer_demand = demand_allowed() ? 1 : 0;
  if (cross) {
    // the donation ring lives in the rank's segment, where the peers can reach it; slot numbers are the same on
    // every rank (n_initial .. n_initial + ring), so the pointers are shifted by the frontier's length
    const int ring = c->ring_frames(fw);
    if (ring < ring_min_frames(p->n_warps))
      return fail(CSOLVE_ERR_CAPACITY, "comm: the donation ring holds " + std::to_string(ring) + " frames of this model, " + std::to_string(ring_min_frames(p->n_warps)) + " are needed");
    a.pool_cap = n_items + ring;
    a.pool = c->ring_of(c->rank) - (size_t)n_items * fw;
    a.ready = c->ready_of(c->rank) - n_items;
    for (int r = 0; r < c->world; r++) {
      a.peer_ctl[r] = c->ctl_of(r);
      a.peer_pool[r] = c->ring_of(r) - (size_t)n_items * fw;
      a.peer_ready[r] = c->ready_of(r) - n_items;
    }
    CUDA_TRY(cudaMemsetAsync(c->ready_of(c->rank), 0, (size_t)ring * sizeof(int32_t), st));
    const int32_t zeros[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // demand[8], inflight
    CUDA_TRY(cudaMemcpyAsync(c->block(c->rank)->demand, zeros, 9 * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(&c->block(c->rank)->ring_open, zeros, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cmalloc(C, &d_cstate, 4 * sizeof(int32_t)));
  } else {
    CUDA_TRY(cudaMemsetAsync(p->ready, 0, (size_t)p->pool_cap * sizeof(int32_t), st));
    if (p->pool_cap - n_items < ring_min_frames(p->n_warps)) return fail(CSOLVE_ERR_CAPACITY, "no room for donated frames behind the root frontier");
  }
  a.gprio = d_gprio;     // the breadth-first expansion above stays deterministic (identical on every rank)
  if (learn) a.ng = p->ng;
  if (cross) comm_args = &a;
  ctl.item_next = 0; ctl.item_count = 0; ctl.init_next = 0; ctl.busy = 0; ctl.hungry = 0;
  ctl.signal = stopped ? SIG_STOP : SIG_RUN;
  if (c != nullptr && c->rank == 0 && front_pool != nullptr) {
    // the peers may already be claiming from init_next (line 1 of the block): upload everything but that line
    CUDA_TRY(cudaMemcpyAsync(dctl, &ctl, offsetof(SearchCtl, init_next), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(&dctl->item_next, &ctl.item_next, sizeof(ctl) - offsetof(SearchCtl, item_next), cudaMemcpyHostToDevice, st));
  } else {
    CUDA_TRY(cudaMemcpyAsync(dctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, st));
  }
  // ---- restarts (src/csolve.c:76-83, 264-276): Knuth's rendering of the Luby sequence, one unit = restart_frequency
  //      failed nodes per search warp; see csolve_solve_options.restart_frequency
  const bool restarting = opt.restart_frequency > 0 && m.objective == CSOLVE_OBJ_ANY && d_gprio != nullptr && c == nullptr && !batch && !learn;
  unsigned long long luby_counter = 1, luby_threshold = 1;
  uint64_t n_restarts = 0;
  auto fail_limit_now = [&]() {
    const double lim = (double)luby_threshold * (double)opt.restart_frequency * (double)p->n_warps;
    return (int32_t)std::min(lim, 2.0e9);
  };
  if (restarting) a.fail_limit = fail_limit_now();
  int idle_now = p->n_warps;          // every warp starts without a stack
  int busy = 0;
  bool timed_out = false;
  uint64_t slices = 0;
  // With an exchange callback (one process per GPU) every rank keeps calling it once per slice until ALL ranks
  // are done, so the collectives inside it stay matched; a rank that ran dry simply waits there.
  bool local_done = stopped || n_items == 0;
  const bool is_min = m.objective == CSOLVE_OBJ_MIN;
  for (;;) {
    if (!local_done) {
      COMM_TRACE(c, "launch search slice");
      // the peers may serve this rank's tickets while its kernel runs (k_rebalance closes the ring again)
      if (cross) {
        CUDA_TRY(cudaMemcpyAsync(&c->block(c->rank)->ring_open, &epoch, sizeof(int32_t), cudaMemcpyHostToDevice, st));
        a.peer_demand = demand_allowed() ? 1 : 0;
      }
      CUDA_TRY(launch_search(a, p->grid, false, st, backjump)); launches++;
      if (getenv("CSOLVE_DEBUG_SYNC")) {
        const cudaError_t es = cudaStreamSynchronize(st);
        fprintf(stderr, "[csolve] depth-first kernel: %s (solbuf %p cap %d, pool %p cap %d, stacks %p)\n", cudaGetErrorString(es), (void *)p->solbuf,
                p->sol_cap, (void *)a.pool, a.pool_cap, (void *)a.stacks);
      }
      CUDA_TRY(launch_rebalance(a, p->scratch, st)); launches++;
      CUDA_TRY(cudaMemcpyAsync(&ctl, dctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
      slices++; n_sliced++;
      COMM_TRACE(c, "slice back");
      { const int rcd = drain_solutions(); if (rcd != CSOLVE_OK) return rcd; }
      busy = ctl.busy;
      idle_now = p->n_warps - busy;
      if (ctl.signal == SIG_STOP || ctl.busy == 0) local_done = true;
      if (restarting && !local_done && ctl.fails > a.fail_limit) {
        // RESTART (src/csolve.c:380-385): every open frame is dropped, the root is expanded again -- its top levels now
        // in the order of the failure-driven priorities the search has learned (they survive, src/csolve.c:459-462)
        if ((luby_counter & (~luby_counter + 1)) == luby_threshold) { luby_counter++; luby_threshold = 1; } else { luby_threshold <<= 1; }
        n_restarts++;
        std::vector<int32_t> gp(V), ord(V);
        CUDA_TRY(cudaMemcpyAsync(gp.data(), d_gprio, (size_t)V * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int v = 0; v < V; v++) ord[v] = v;
        std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return gp[x] > gp[y]; });
        CUDA_TRY(cudaMemcpyAsync(const_cast<int32_t *>(m.order), ord.data(), (size_t)V * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        p->order_dirty = true;
        rc = upload_root_frame(ord[0]);
        if (rc != CSOLVE_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(p->wstate, ws.data(), ws.size() * sizeof(WarpState), cudaMemcpyHostToDevice, st));
        pin = p->pool_a; pout = p->pool_b; n_items = 1;
        a.gprio = nullptr; a.fail_limit = 0;
        ctl.signal = SIG_RUN;
        rc = expand_root();
        if (rc != CSOLVE_OK) return rc;
        a.items = pin; a.items_out = nullptr;
        a.pool = pin; a.front_pool = pin; a.n_initial = n_items; a.gprio = d_gprio;
        a.fail_limit = fail_limit_now();
        CUDA_TRY(cudaMemsetAsync(p->ready, 0, (size_t)p->pool_cap * sizeof(int32_t), st));
        if (p->pool_cap - n_items < ring_min_frames(p->n_warps)) return fail(CSOLVE_ERR_CAPACITY, "no room for donated frames behind the root frontier");
        ctl.item_next = 0; ctl.item_count = 0; ctl.init_next = 0; ctl.busy = 0; ctl.hungry = 0; ctl.fails = 0;
        if (ctl.signal == SIG_STOP || n_items == 0) local_done = true;      // the expansion itself found a solution / emptied the tree
        CUDA_TRY(cudaMemcpyAsync(dctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice, st));
      }
      if (!local_done && opt.time_limit_ms > 0) {
        const auto now = std::chrono::steady_clock::now();
        if (std::chrono::duration_cast<std::chrono::milliseconds>(now - wall0).count() > opt.time_limit_ms) { timed_out = true; local_done = true; }
      }
      if (cross && local_done && ctl.signal != SIG_STOP && !timed_out) {
        // This rank ran dry. It asks its peers for frames (CommBlock::demand) and waits here, its ring open to
        // them, until one arrives -- or until no rank is active any more: the search is over everywhere.
        for (unsigned spin = 0;; spin++) {
          int32_t stt[4] = {0, 0, 0, 0};
          // per peer and per look: the ring holds 4 x n_warps frames
          CUDA_TRY(launch_comm_state(a, demand_allowed() ? std::max(p->n_warps / (4 * (c->world - 1)), 32) : 0, 0, d_cstate, st)); launches++;
          CUDA_TRY(cudaMemcpyAsync(stt, d_cstate, sizeof(stt), cudaMemcpyDeviceToHost, st));
          CUDA_TRY(cudaStreamSynchronize(st));
          if (stt[2]) break;                                   // ANY: a peer has a solution
          if (stt[0]) { local_done = false; break; }           // frames arrived (or the shared frontier is not drained yet)
          if (stt[1] == 0) break;                              // every rank is idle
          if (stt[1] < 0) return fail(CSOLVE_ERR_CUDA, "comm: the ranks are out of step");
          if (opt.time_limit_ms > 0 && std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - wall0).count() > opt.time_limit_ms) { timed_out = true; break; }
          if (std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count() > 3600.0) return fail(CSOLVE_ERR_CUDA, "comm: waited an hour for the other ranks");
          if (spin > 50) std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
      }
    }
    if (p->exchange == nullptr) {
      if (local_done) break;
      continue;
    }
    // periodic incumbent / first-solution exchange between the ranks (NCCL all-reduce in the callback)
    int32_t best = ctl.best, found = (m.objective == CSOLVE_OBJ_ANY && ctl.signal == SIG_STOP) ? 1 : 0;
    const int all_done = p->exchange(p->exchange_user, &best, &found, local_done ? 1 : 0);
    if (m.obj_var >= 0 && (is_min ? best < ctl.best : best > ctl.best)) {
      ctl.best = best;     // another rank found a better incumbent: prune with it from the next slice on
      CUDA_TRY(cudaMemcpyAsync(&dctl->best, &ctl.best, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    if (found && m.objective == CSOLVE_OBJ_ANY && !local_done) {
      local_done = true;   // found_any() on another rank (src/csolve.c:207-209)
    }
    if (all_done) break;
    if (p->rebalance != nullptr && !batch && !(found && m.objective == CSOLVE_OBJ_ANY)) {
      // frontier rebalancing: ranks that ran dry receive frames split off the busy warps of the others. Every rank
      // makes the call (the collectives inside must stay matched); one that hit its time limit poses as a rank with
      // a single busy warp: it neither receives nor has anything to spare.
      p->parked = &a;
      const int got = timed_out ? p->rebalance(p->rebalance_user, p, 0, 1, fw)
                                : p->rebalance(p->rebalance_user, p, local_done ? p->n_warps : idle_now, local_done ? 0 : busy, fw);
      p->parked = nullptr;
      if (got < 0) return fail(CSOLVE_ERR_INVALID, "rebalance callback failed");
      if (got > 0) local_done = false;
    }
  }
  COMM_TRACE(c, "search over");
  CUDA_TRY(cudaEventRecord(ev2, st));

  // ---- results ---------------------------------------------------------------------------------------------
  CUDA_TRY(launch_reduce_counters(p->wcount, p->n_warps, p->totals, st)); launches++;
  unsigned long long tot[CNT_WIDTH];
  CUDA_TRY(cudaMemcpyAsync(tot, p->totals, sizeof(tot), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&ctl, dctl, sizeof(ctl), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (getenv("CSOLVE_DEBUG")) {
    // load-balance diagnostics of the depth-first phase
    std::vector<unsigned long long> wc((size_t)p->n_warps * CNT_WIDTH);
    CUDA_TRY(cudaMemcpy(wc.data(), p->wcount, wc.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double wait = 0, claims = 0, lw = 0, lw_max = 0, lw_min = 1e30, nmax = 0;
    float dbg_ms = 0;
    cudaEventElapsedTime(&dbg_ms, ev1, ev2);
    for (int w = 0; w < p->n_warps; w++) {
      const unsigned long long *c = &wc[(size_t)w * CNT_WIDTH];
      wait += (double)c[CNT_WAIT]; claims += (double)c[CNT_CLAIMS];
      const double l = (double)c[CNT_LASTWORK];
      lw += l; lw_max = std::max(lw_max, l); lw_min = std::min(lw_min, l); nmax = std::max(nmax, (double)c[CNT_NODES]);
    }
    const double khz = (double)C->clock_khz;
    fprintf(stderr, "[csolve] depth-first phase: %.2f ms, %llu slices, %d warps, per warp: waited %.3f ms, %.1f claims, last node at %.3f ms "
                    "(min %.3f, max %.3f), most nodes on one warp %.0f (avg %.0f), %d frames donated\n", dbg_ms, (unsigned long long)slices, p->n_warps,
            wait / p->n_warps / khz, claims / p->n_warps, lw / p->n_warps / khz, lw_min / khz, lw_max / khz, nmax,
            (double)tot[CNT_NODES] / p->n_warps, ctl.item_count);
    fprintf(stderr, "[csolve]   polls %llu, donation wanted %llu, donated %llu; nodes %llu, frame refreshes %llu, props %llu, clause visits %llu\n", tot[CNT_POLLS], tot[CNT_WANTED], tot[CNT_DONATED],
            tot[CNT_NODES], tot[CNT_REFRESH], tot[CNT_PROPS], tot[CNT_VISITS]);
    {
      std::vector<double> lws, wts;
      for (int w = 0; w < p->n_warps; w++) { lws.push_back((double)wc[(size_t)w * CNT_WIDTH + CNT_LASTWORK] / khz); wts.push_back((double)wc[(size_t)w * CNT_WIDTH + CNT_WAIT] / khz); }
      std::sort(lws.begin(), lws.end()); std::sort(wts.begin(), wts.end());
      fprintf(stderr, "[csolve]   last-node deciles:");
      for (int d = 0; d <= 10; d++) fprintf(stderr, " %.3f", lws[std::min<size_t>(lws.size() - 1, lws.size() * d / 10)]);
      fprintf(stderr, "  p99 %.3f p99.9 %.3f\n[csolve]   waited deciles:", lws[lws.size() * 99 / 100], lws[lws.size() * 999 / 1000]);
      for (int d = 0; d <= 10; d++) fprintf(stderr, " %.3f", wts[std::min<size_t>(wts.size() - 1, wts.size() * d / 10)]);
      fprintf(stderr, "\n");
    }
  }
  { const int rcd = drain_solutions(); if (rcd != CSOLVE_OK) return rcd; }
  if (sinking && sunk != tot[CNT_SOLUTIONS])
    return fail(CSOLVE_ERR_CAPACITY, "solutions lost between the device and the sink: " + std::to_string(sunk) + " of " + std::to_string(tot[CNT_SOLUTIONS]));
  p->n_stored = std::min(ctl.n_stored, p->sol_cap);
  p->sol_host.resize((size_t)p->n_stored * (V + 1));
  if (p->n_stored > 0) {
    CUDA_TRY(cudaMemcpy(p->sol_host.data(), p->solbuf, p->sol_host.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  if (m.obj_var >= 0 && p->n_stored > 1) {
    // order the stored incumbents as a chain of improvements; the optimum is last
    std::vector<int> idx(p->n_stored);
    for (int i = 0; i < p->n_stored; i++) idx[i] = i;
    const bool is_min = m.objective == CSOLVE_OBJ_MIN;
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) {
      const int kx = p->sol_host[(size_t)x * (V + 1) + V], ky = p->sol_host[(size_t)y * (V + 1) + V];
      return is_min ? kx > ky : kx < ky;
    });
    std::vector<int32_t> sorted(p->sol_host.size());
    for (int i = 0; i < p->n_stored; i++)
      memcpy(&sorted[(size_t)i * (V + 1)], &p->sol_host[(size_t)idx[i] * (V + 1)], sizeof(int32_t) * (V + 1));
    p->sol_host.swap(sorted);
  }

  // (ranks of a comm / of an exchange: the incumbent may be a peer's, whose buffer then holds the witness -- the caller
  //  that merges the ranks' results checks it, csolve_gpu_group_solve / csolve_b200.distributed)
  if (m.obj_var >= 0 && tot[CNT_SOLUTIONS] > 0 && c == nullptr && p->exchange == nullptr &&
      (p->n_stored == 0 || p->sol_host[(size_t)(p->n_stored - 1) * (V + 1) + V] != ctl.best)) {
    return fail(CSOLVE_ERR_CAPACITY, "the witness of the optimum was overwritten in the solution ring (more than " +
                                         std::to_string(p->sol_cap) + " concurrent incumbents); raise max_solutions");
  }
  if (opt.sample_mod != 0u) {
    CUDA_TRY(cudaMemcpy(&p->sample_seen, p->sample_n, sizeof(int32_t), cudaMemcpyDeviceToHost));
  }

  float ms_expand = 0, ms_search = 0;
  cudaEventElapsedTime(&ms_expand, ev0, ev1);
  cudaEventElapsedTime(&ms_search, ev1, ev2);

  if (batch) {
    if (root_solutions) CUDA_TRY(cudaMemcpy(root_solutions, d_rsol, (size_t)n_roots * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    if (root_failed) CUDA_TRY(cudaMemcpy(root_failed, d_rfail, n_roots, cudaMemcpyDeviceToHost));
  }
  res->solutions = tot[CNT_SOLUTIONS];
  res->nodes = tot[CNT_NODES];
  res->cuts = tot[CNT_CUTS];
  res->props = tot[CNT_PROPS];
  res->clause_visits = tot[CNT_VISITS];
  res->best = ctl.best;
  res->has_solution = tot[CNT_SOLUTIONS] > 0;
  res->timed_out = timed_out;
  res->n_stored = p->n_stored;
  res->kernel_ms = ms_search;
  res->expand_ms = ms_expand;
  res->kernel_launches = launches;
  res->restarts = n_restarts;
  if (learn) {
    int32_t cnt[8];
    CUDA_TRY(cudaMemcpy(cnt, p->ng.counters, sizeof(cnt), cudaMemcpyDeviceToHost));
    res->conflicts = std::min(cnt[0], p->ng.cap_ng);
    res->conflicts_abandoned = (uint64_t)cnt[3] + (uint64_t)cnt[4];
    res->backjumps = (uint64_t)cnt[5];
  }
  (void)slices;
  return CSOLVE_OK;
}

}  // namespace

extern "C" int csolve_gpu_solve(csolve_gpu_problem *p, const csolve_solve_options *opt, csolve_gpu_result *res) {
  return solve_impl(p, nullptr, opt, res, 0, nullptr, nullptr, nullptr);
}

// ---- comm ------------------------------------------------------------------------------------------------------------
extern "C" int csolve_gpu_comm_create(int32_t device, int32_t rank, int32_t world, size_t frontier_bytes, csolve_gpu_comm **out) {
  if (out == nullptr || world < 1 || world > COMM_MAX_RANKS || rank < 0 || rank >= world) return fail(CSOLVE_ERR_INVALID, "comm: bad rank / world (at most 8 ranks)");
  *out = nullptr;
  DeviceCtx *C = nullptr;
  const int rc = device_ctx(device, &C);
  if (rc != CSOLVE_OK) return rc;
  std::unique_ptr<csolve_gpu_comm> c(new csolve_gpu_comm);
  c->ctx = C; c->rank = rank; c->world = world;
  c->front_bytes = rank == 0 ? (frontier_bytes ? frontier_bytes : (size_t)256 << 20) : 0;
  if (world == 1) { c->front_bytes = 0; c->ring_bytes = 0; }      // a comm of one rank is inert: no ring, no frontier
  const size_t bytes = SEG_RING + c->ring_bytes + c->front_bytes;
  CUDA_TRY(cudaMalloc((void **)&c->seg, bytes));          // not from the block cache: the allocation is exported whole
  CUDA_TRY(cudaMemset(c->seg, 0, SEG_RING));
  const unsigned long long none = ~0ull;
  CUDA_TRY(cudaMemcpy(&reinterpret_cast<CommBlock *>(c->seg + SEG_COMM)->rmin64, &none, sizeof(none), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
  CUDA_TRY(cudaHostAlloc((void **)&c->hblk, 2 * sizeof(CommBlock), cudaHostAllocDefault));
  c->peer_seg[rank] = c->seg;
  c->connected = world == 1;
  *out = c.release();
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_comm_handle(csolve_gpu_comm *c, void *handle) {
  if (c == nullptr || handle == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == CSOLVE_COMM_HANDLE_BYTES, "handle size");
  CUDA_TRY(cudaSetDevice(c->ctx->device));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, c->seg));
  memcpy(handle, &h, sizeof(h));
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_comm_connect(csolve_gpu_comm *c, const void *handles) {
  if (c == nullptr || handles == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(c->ctx->device));
  for (int r = 0; r < c->world; r++) {
    if (r == c->rank || c->peer_seg[r] != nullptr) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char *)handles + (size_t)r * CSOLVE_COMM_HANDLE_BYTES, sizeof(h));
    void *ptr = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_seg[r] = (unsigned char *)ptr;
    c->opened[r] = true;
  }
  c->connected = true;
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_comm_connect_local(csolve_gpu_comm **comms, int32_t world) {
  if (comms == nullptr || world < 1 || world > COMM_MAX_RANKS) return fail(CSOLVE_ERR_INVALID, "bad arguments");
  for (int i = 0; i < world; i++) {
    if (comms[i] == nullptr || comms[i]->rank != i || comms[i]->world != world) return fail(CSOLVE_ERR_INVALID, "comms must be given in rank order");
  }
  for (int i = 0; i < world; i++) {
    CUDA_TRY(cudaSetDevice(comms[i]->ctx->device));
    for (int j = 0; j < world; j++) {
      if (i == j) continue;
      const int di = comms[i]->ctx->device, dj = comms[j]->ctx->device;
      if (di != dj) {
        int can = 0;
        CUDA_TRY(cudaDeviceCanAccessPeer(&can, di, dj));
        if (!can) return fail(CSOLVE_ERR_UNSUPPORTED, "no peer access between devices " + std::to_string(di) + " and " + std::to_string(dj));
        const cudaError_t e = cudaDeviceEnablePeerAccess(dj, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else CUDA_TRY(e);
      }
      comms[i]->peer_seg[j] = comms[j]->seg;
    }
    comms[i]->connected = true;
  }
  return CSOLVE_OK;
}

extern "C" void csolve_gpu_comm_destroy(csolve_gpu_comm *c) {
  if (c == nullptr) return;
  cudaSetDevice(c->ctx->device);
  for (int r = 0; r < c->world; r++) if (c->opened[r]) cudaIpcCloseMemHandle(c->peer_seg[r]);
  cudaFree(c->seg);
  if (c->aux) cudaStreamDestroy(c->aux);
  if (c->hblk) cudaFreeHost(c->hblk);
  delete c;
}

extern "C" int csolve_gpu_solve_comm(csolve_gpu_problem *p, csolve_gpu_comm *c, const csolve_solve_options *opt, csolve_gpu_result *res) {
  if (c == nullptr) return fail(CSOLVE_ERR_INVALID, "null comm");
  return solve_impl(p, c, opt, res, 0, nullptr, nullptr, nullptr);
}

// ---- group: all GPUs of this process on one tree, one host thread per device ---------------------------------------
struct csolve_gpu_group {
  std::vector<int> devices;
  std::vector<csolve_gpu_comm *> comms;
  std::vector<csolve_gpu_problem *> probs;
  int32_t n_vars = 0, objective = 0, obj_var = -1;
  std::vector<int32_t> sols;          // merged stored assignments: (n_vars values, key) each
  int32_t n_stored = 0;
  csolve_solution_fn sink = nullptr;  // the devices' sinks funnel into this one, one call at a time
  void *sink_user = nullptr;
  std::mutex sink_mu;
};

extern "C" int csolve_gpu_device_count(int32_t *n) {
  if (n == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  int k = 0;
  const cudaError_t e = cudaGetDeviceCount(&k);
  if (e != cudaSuccess || k == 0) { *n = 0; return fail(CSOLVE_ERR_NO_DEVICE, "no CUDA device (the search path has no CPU fallback)"); }
  *n = k;
  return CSOLVE_OK;
}

extern "C" void csolve_gpu_group_destroy(csolve_gpu_group *g) {
  if (g == nullptr) return;
  for (auto *p : g->probs) delete p;
  for (auto *c : g->comms) csolve_gpu_comm_destroy(c);
  delete g;
}

extern "C" int csolve_gpu_group_create(int32_t n_devices, const int32_t *devices, size_t frontier_bytes, csolve_gpu_group **out) {
  if (out == nullptr || n_devices < 1 || n_devices > COMM_MAX_RANKS) return fail(CSOLVE_ERR_INVALID, "group: 1..8 devices");
  *out = nullptr;
  std::unique_ptr<csolve_gpu_group, void (*)(csolve_gpu_group *)> g(new csolve_gpu_group, csolve_gpu_group_destroy);
  for (int i = 0; i < n_devices; i++) g->devices.push_back(devices ? devices[i] : i);
  for (int i = 0; i < n_devices; i++) {
    csolve_gpu_comm *c = nullptr;
    const int rc = csolve_gpu_comm_create(g->devices[i], i, n_devices, frontier_bytes, &c);
    if (rc != CSOLVE_OK) return rc;
    g->comms.push_back(c);
  }
  const int rc = csolve_gpu_comm_connect_local(g->comms.data(), n_devices);
  if (rc != CSOLVE_OK) return rc;
  *out = g.release();
  return CSOLVE_OK;
}

namespace {
// runs fn(i) on one host thread per device and returns the first error (with its text)
template <class F>
int group_parallel(csolve_gpu_group *g, F fn) {
  const int n = (int)g->devices.size();
  std::vector<int> rc(n, CSOLVE_OK);
  std::vector<std::string> err(n);
  std::vector<std::thread> th;
  for (int i = 1; i < n; i++) th.emplace_back([&, i]() { rc[i] = fn(i); if (rc[i] != CSOLVE_OK) err[i] = csolve_front::last_error(); });
  rc[0] = fn(0);
  if (rc[0] != CSOLVE_OK) err[0] = csolve_front::last_error();
  for (auto &t : th) t.join();
  for (int i = 0; i < n; i++) if (rc[i] != CSOLVE_OK) return fail(rc[i], "device " + std::to_string(g->devices[i]) + ": " + err[i]);
  return CSOLVE_OK;
}
}  // namespace

namespace {
void group_sink(void *user, const int32_t *values, int32_t n, int32_t stride) {
  csolve_gpu_group *g = static_cast<csolve_gpu_group *>(user);
  std::lock_guard<std::mutex> lk(g->sink_mu);
  if (g->sink) g->sink(g->sink_user, values, n, stride);
}
}  // namespace

extern "C" int csolve_gpu_group_set_solution_sink(csolve_gpu_group *g, csolve_solution_fn fn, void *user) {
  if (g == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  g->sink = fn;
  g->sink_user = user;
  for (auto *p : g->probs) if (p) { p->sink = fn ? group_sink : nullptr; p->sink_user = g; }
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_group_load(csolve_gpu_group *g, const csolve_flat_model *m) {
  if (g == nullptr || m == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  for (auto *p : g->probs) delete p;
  g->probs.assign(g->devices.size(), nullptr);
  g->n_vars = m->n_vars; g->objective = m->objective; g->obj_var = m->obj_var;
  // the search workspace is allocated here as well: with peer mappings in the process a cudaMalloc takes milliseconds,
  // and a device that is still allocating when the others start searching gets nothing of a short search
  const int rc = group_parallel(g, [&](int i) {
    const int r = csolve_gpu_load_device(m, g->devices[i], &g->probs[i]);
    if (r != CSOLVE_OK) return r;
    csolve_solve_options o;
    memset(&o, 0, sizeof(o));
    if (g->sink) { g->probs[i]->sink = group_sink; g->probs[i]->sink_user = g; }
    return ensure_workspace(g->probs[i], o, false, 0, false, false);
  });
  if (rc == CSOLVE_OK && g->sink) for (auto *p : g->probs) { p->sink = group_sink; p->sink_user = g; }
  return rc;
}

extern "C" int csolve_gpu_group_solve(csolve_gpu_group *g, const csolve_solve_options *opt, csolve_gpu_result *res,
                                      csolve_gpu_result *per_device) {
  if (g == nullptr || res == nullptr || g->probs.empty() || g->probs[0] == nullptr) return fail(CSOLVE_ERR_INVALID, "group: no model loaded");
  const int n = (int)g->devices.size();
  std::vector<csolve_gpu_result> r(n);
  const int rc = group_parallel(g, [&](int i) { return solve_impl(g->probs[i], g->comms[i], opt, &r[i], 0, nullptr, nullptr, nullptr); });
  if (rc != CSOLVE_OK) return rc;
  // the reference's shared page (struct shared_t, src/csolve.h:259-266): counters add up, the incumbent is the best one
  memset(res, 0, sizeof(*res));
  const bool is_min = g->objective == CSOLVE_OBJ_MIN, is_opt = g->obj_var >= 0;
  for (int i = 0; i < n; i++) {
    res->solutions += r[i].solutions; res->nodes += r[i].nodes; res->cuts += r[i].cuts; res->props += r[i].props;
    res->clause_visits += r[i].clause_visits; res->kernel_launches += r[i].kernel_launches;
    res->conflicts += r[i].conflicts; res->conflicts_abandoned += r[i].conflicts_abandoned; res->backjumps += r[i].backjumps;
    res->timed_out |= r[i].timed_out;
    res->kernel_ms = std::max(res->kernel_ms, r[i].kernel_ms); res->expand_ms = std::max(res->expand_ms, r[i].expand_ms);
    if (r[i].has_solution) {
      if (!res->has_solution || (is_opt && (is_min ? r[i].best < res->best : r[i].best > res->best))) res->best = r[i].best;
      res->has_solution = 1;
    }
    if (per_device) per_device[i] = r[i];
  }
  if (g->objective == CSOLVE_OBJ_ANY && res->solutions > 1) res->solutions = 1;     // two ranks may both have finished a leaf
  // stored assignments of all devices; MIN / MAX: one chain of improvements, the optimum last
  const int V = g->n_vars;
  struct Ref { int dev, idx, key; };
  std::vector<Ref> refs;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < g->probs[i]->n_stored; k++) refs.push_back(Ref{i, k, g->probs[i]->sol_host[(size_t)k * (V + 1) + V]});
  if (is_opt) std::stable_sort(refs.begin(), refs.end(), [&](const Ref &x, const Ref &y) { return is_min ? x.key > y.key : x.key < y.key; });
  if (g->objective == CSOLVE_OBJ_ANY && refs.size() > 1) refs.resize(1);
  g->sols.resize(refs.size() * (size_t)(V + 1));
  for (size_t k = 0; k < refs.size(); k++)
    memcpy(&g->sols[k * (V + 1)], &g->probs[refs[k].dev]->sol_host[(size_t)refs[k].idx * (V + 1)], sizeof(int32_t) * (V + 1));
  g->n_stored = (int32_t)refs.size();
  res->n_stored = g->n_stored;
  if (is_opt && res->has_solution && (refs.empty() || refs.back().key != res->best))
    return fail(CSOLVE_ERR_CAPACITY, "the witness of the optimum was overwritten in a device's solution ring; raise max_solutions");
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_group_get_solution(csolve_gpu_group *g, int32_t i, int32_t *values, int32_t *key) {
  if (g == nullptr || values == nullptr || i < 0 || i >= g->n_stored) return fail(CSOLVE_ERR_INVALID, "solution index out of range");
  memcpy(values, &g->sols[(size_t)i * (g->n_vars + 1)], sizeof(int32_t) * g->n_vars);
  if (key) *key = g->sols[(size_t)i * (g->n_vars + 1) + g->n_vars];
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_set_solution_sink(csolve_gpu_problem *p, csolve_solution_fn fn, void *user) {
  if (p == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  p->sink = fn;
  p->sink_user = user;
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_set_exchange(csolve_gpu_problem *p, csolve_exchange_fn fn, void *user) {
  if (p == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  p->exchange = fn;
  p->exchange_user = user;
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_set_rebalance(csolve_gpu_problem *p, csolve_rebalance_fn fn, void *user) {
  if (p == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  p->rebalance = fn;
  p->rebalance_user = user;
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_export_frames(csolve_gpu_problem *p, int32_t max_frames, int32_t *frames, int32_t *n_out) {
  if (p == nullptr || frames == nullptr || n_out == nullptr || max_frames < 0) return fail(CSOLVE_ERR_INVALID, "bad arguments");
  *n_out = 0;
  if (p->parked == nullptr) return fail(CSOLVE_ERR_INVALID, "csolve_gpu_export_frames is only valid inside the rebalance callback");
  if (max_frames == 0) return CSOLVE_OK;
  ENTER(p);
  const SearchArgs &a = *p->parked;
  const size_t bytes = (size_t)max_frames * a.m.frame_words * sizeof(int32_t);
  int32_t *d_buf = nullptr, *d_n = nullptr;
  CUDA_TRY(cmalloc(C, &d_buf, bytes));
  CUDA_TRY(cmalloc(C, &d_n, sizeof(int32_t)));
  cudaError_t e = launch_export_frames(a, d_buf, max_frames, d_n, p->stream);
  int32_t n = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, p->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
  if (e == cudaSuccess && n > 0) e = cudaMemcpy(frames, d_buf, (size_t)n * a.m.frame_words * sizeof(int32_t), cudaMemcpyDeviceToHost);
  C->cfree(d_buf); C->cfree(d_n);
  if (e != cudaSuccess) return fail(CSOLVE_ERR_CUDA, std::string("export frames: ") + cudaGetErrorString(e));
  *n_out = n;
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_import_frames(csolve_gpu_problem *p, const int32_t *frames, int32_t n_frames) {
  if (p == nullptr || frames == nullptr || n_frames < 0) return fail(CSOLVE_ERR_INVALID, "bad arguments");
  if (p->parked == nullptr) return fail(CSOLVE_ERR_INVALID, "csolve_gpu_import_frames is only valid inside the rebalance callback");
  if (n_frames == 0) return CSOLVE_OK;
  ENTER(p);
  const SearchArgs &a = *p->parked;
  if (n_frames > a.n_warps) return fail(CSOLVE_ERR_CAPACITY, "more frames than the donation ring holds");
  const size_t bytes = (size_t)n_frames * a.m.frame_words * sizeof(int32_t);
  int32_t *d_buf = nullptr;
  CUDA_TRY(cmalloc(C, &d_buf, bytes));
  cudaError_t e = cudaMemcpyAsync(d_buf, frames, bytes, cudaMemcpyHostToDevice, p->stream);
  if (e == cudaSuccess) e = launch_import_frames(a, d_buf, n_frames, p->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
  C->cfree(d_buf);
  if (e != cudaSuccess) return fail(CSOLVE_ERR_CUDA, std::string("import frames: ") + cudaGetErrorString(e));
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_solve_batch(csolve_gpu_problem *p, const csolve_solve_options *opt, int32_t n_roots,
                                      const int32_t *root_dom, uint32_t *root_solutions, uint8_t *root_failed,
                                      csolve_gpu_result *res) {
  if (n_roots <= 0 || root_dom == nullptr) return fail(CSOLVE_ERR_INVALID, "bad arguments");
  return solve_impl(p, nullptr, opt, res, n_roots, root_dom, root_solutions, root_failed);
}

extern "C" int csolve_gpu_get_nogoods(csolve_gpu_problem *p, int32_t *lits, int32_t cap_lits, int32_t *starts, int32_t cap_ng,
                                      int32_t *n_out) {
  if (p == nullptr || lits == nullptr || starts == nullptr || n_out == nullptr) return fail(CSOLVE_ERR_INVALID, "null argument");
  *n_out = 0;
  starts[0] = 0;
  if (p->ng.lits == nullptr) return CSOLVE_OK;
  ENTER(p);
  int32_t cnt[8];
  CUDA_TRY(cudaMemcpy(cnt, p->ng.counters, sizeof(cnt), cudaMemcpyDeviceToHost));
  const int n = std::min(std::min(cnt[0], p->ng.cap_ng), cap_ng);
  std::vector<int32_t> st(n), ln(n);
  if (n > 0) {
    CUDA_TRY(cudaMemcpy(st.data(), p->ng.start, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(ln.data(), p->ng.len, (size_t)n * 4, cudaMemcpyDeviceToHost));
  }
  int total = 0, kept = 0;
  for (int k = 0; k < n; k++) {
    if (total + ln[k] > cap_lits) break;
    CUDA_TRY(cudaMemcpy(lits + total, p->ng.lits + st[k], (size_t)ln[k] * 4, cudaMemcpyDeviceToHost));
    total += ln[k];
    starts[++kept] = total;
  }
  *n_out = kept;
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_get_samples(csolve_gpu_problem *p, int32_t *records, int32_t cap_records, int32_t *n_out, int32_t *n_seen) {
  if (p == nullptr || n_out == nullptr || cap_records < 0 || (records == nullptr && cap_records > 0)) return fail(CSOLVE_ERR_INVALID, "bad arguments");
  ENTER(p);
  const int n = std::min(std::min(p->sample_seen, p->sample_cap), cap_records);
  if (n > 0) {
    CUDA_TRY(cudaMemcpy(records, p->sample_rec, (size_t)n * sample_words(p->dev.n_vars) * sizeof(int32_t), cudaMemcpyDeviceToHost));
  }
  *n_out = n;
  if (n_seen) *n_seen = p->sample_seen;
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_get_solution_key(csolve_gpu_problem *p, int32_t i, int32_t *key) {
  if (p == nullptr || key == nullptr || i < 0 || i >= p->n_stored) return fail(CSOLVE_ERR_INVALID, "solution index out of range");
  *key = p->sol_host[(size_t)i * (p->dev.n_vars + 1) + p->dev.n_vars];
  return CSOLVE_OK;
}

extern "C" int csolve_gpu_get_solution(csolve_gpu_problem *p, int32_t i, int32_t *values) {
  if (p == nullptr || values == nullptr || i < 0 || i >= p->n_stored) return fail(CSOLVE_ERR_INVALID, "solution index out of range");
  memcpy(values, &p->sol_host[(size_t)i * (p->dev.n_vars + 1)], sizeof(int32_t) * p->dev.n_vars);
  return CSOLVE_OK;
}
