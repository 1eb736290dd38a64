// simt_emu.h -- TEST-ONLY: a small SIMT emulator that lets g++ compile and run the product's search kernel source
// (csolve_b200/csrc/kernels.cu, up to the end of k_search) on the CPU, so that the whole node loop -- claims, fixpoint,
// push / pop, leaves, donation, conflict analysis, back-jump -- can be checked against the oracle without a GPU.
// Every CUDA thread of a launch is a fiber with a stack of its own; a fiber runs until it reaches a warp collective
// (__shfl_sync, __ballot_sync, __syncwarp, ...), a block barrier or a wait (__nanosleep), where the next lane of the
// warp / the next warp is resumed. Collectives complete when all 32 lanes of the warp have arrived (the kernels only
// ever use the full mask) and check that every lane arrived from the same source line.
// What this does NOT model: the memory system (every store is visible at once), real concurrency between warps, and
// lanes running ahead of a missing __syncwarp -- a kernel that passes here has its logic checked, not its fences.
// Not part of the library, never a fallback: the product has no CPU path (tests/test_abi.py).
#pragma once
#include <cuda_runtime.h>      // vector types (int2, int4, uint3, dim3), make_int4 -- host declarations only under g++
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <list>
#include <utility>
#include <vector>

#if !defined(__x86_64__)
#error "simt_emu.h: the fiber switch is written for x86-64"
#endif

#undef __device__
#undef __host__
#undef __global__
#undef __shared__
#undef __forceinline__
#undef __noinline__
#undef __launch_bounds__
#undef __align__
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
// __noinline__ stays undefined (libstdc++ spells attributes with it); the kernel source's uses are rewritten to EMU_NOINLINE
#define EMU_NOINLINE __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

namespace emu {

struct Fiber {
  void *sp = nullptr;
  char *stack = nullptr;
  uint3 tid{0, 0, 0}, bid{0, 0, 0};
  int lane = 0, warp = 0, block = 0;      // warp: index over the whole grid
  bool done = false;
};

struct Warp {
  int first = 0;            // index of lane 0 in the fiber table
  int alive = 32;
  int cur = 0;              // lane to resume when the warp gets its turn
  int arrived = 0;
  unsigned gen = 0;
  uint64_t val[2][32];
  long site = 0;
  const char *op = nullptr;
  long long clock_val = 0;
  unsigned clock_gen = 0;
  int blocked = 0;          // lanes waiting at a collective or a block barrier
};

struct Block {
  int arrived = 0; unsigned gen = 0; int threads = 0; char *smem = nullptr;
  std::list<std::pair<long, std::vector<unsigned long long>>> statics;     // std::list: references handed out stay valid
};

struct Machine {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  std::vector<Block> blocks;
  Fiber *cur = nullptr;
  void *main_sp = nullptr;
  dim3 block_dim{1, 1, 1}, grid_dim{1, 1, 1};
  long long clock = 0;
  unsigned long long switches = 0, collectives = 0, site_mismatches = 0;
  void (*entry)(void *) = nullptr;
  void *entry_arg = nullptr;
  int live = 0;
};

extern Machine M;

extern "C" void emu_switch(void **save_sp, void *load_sp);

inline void switch_to(Fiber *to) {
  if (to == M.cur) return;
  Fiber *from = M.cur;
  M.cur = to;
  M.switches++;
  emu_switch(&from->sp, to->sp);
}

// Lane schedule: between two collectives the lanes of a warp run one after the other, in ascending lane order (step 1),
// descending (step 31) or any odd stride. Results must not depend on it: a kernel whose answer changes with the stride has
// lanes communicating through memory without a warp sync in between (a race the hardware's lockstep usually hides).
extern int lane_step;
// Warp schedule: a running warp gives the others a turn after every warp_quantum-th collective (and whenever it waits).
extern unsigned warp_quantum;
// block_group > 0: the grid is several kernels running side by side (one per emulated GPU), block_group blocks each;
// blockIdx counts inside the group, emu::M.cur->block / block_group tells the entry function which kernel it is
extern int block_group;

// next live lane of the current warp (cyclic, may be the caller itself)
inline Fiber *next_lane(const Fiber *f) {
  Warp &w = M.warps[f->warp];
  for (int k = 1; k <= 32; k++) {
    Fiber *c = &M.fibers[w.first + ((f->lane + k * lane_step) & 31)];
    if (!c->done) return c;
  }
  return nullptr;
}

inline void yield_inner() {
  Fiber *n = next_lane(M.cur);
  if (n != nullptr) switch_to(n);
}

// the next warp gets its turn; this warp will resume with its next lane
inline void yield_outer() {
  Fiber *f = M.cur;
  Warp &w = M.warps[f->warp];
  Fiber *n = next_lane(f);
  w.cur = n != nullptr ? n->lane : f->lane;
  const int nw = (int)M.warps.size();
  for (int k = 1; k <= nw; k++) {
    Warp &o = M.warps[(f->warp + k) % nw];
    if (o.alive <= 0) continue;
    Fiber *c = &M.fibers[o.first + o.cur];
    if (c->done) { Fiber *c2 = next_lane(c); if (c2 == nullptr) continue; c = c2; }
    switch_to(c);
    return;
  }
}

[[noreturn]] inline void fiber_exit() {
  Fiber *f = M.cur;
  f->done = true;
  M.warps[f->warp].alive--;
  M.live--;
  Fiber *n = next_lane(f);
  if (n != nullptr) { M.cur = n; M.switches++; emu_switch(&f->sp, n->sp); }
  const int nw = (int)M.warps.size();
  for (int k = 1; k <= nw; k++) {
    Warp &o = M.warps[(f->warp + k) % nw];
    if (o.alive <= 0) continue;
    Fiber *c = &M.fibers[o.first + o.cur];
    if (c->done) c = next_lane(c);
    if (c == nullptr) continue;
    M.cur = c; M.switches++;
    emu_switch(&f->sp, c->sp);
  }
  M.cur = nullptr;
  emu_switch(&f->sp, M.main_sp);
  abort();
}

inline void fiber_main() {
  M.entry(M.entry_arg);
  fiber_exit();
}

// ---- warp collectives ------------------------------------------------------------------------------------------
// every lane deposits a value, waits for the other 31, then reads what it needs from the deposited set
inline const uint64_t *warp_exchange(uint64_t v, long site, const char *op) {
  Fiber *f = M.cur;
  Warp &w = M.warps[f->warp];
  if (w.alive != 32) { fprintf(stderr, "[emu] %s with only %d lanes of warp %d alive\n", op, w.alive, f->warp); abort(); }
  const unsigned g = w.gen;
  if (w.arrived == 0) { w.site = site; w.op = op; }
  else if (w.site != site) {
    // __syncwarp() calls of different sites may legally meet; any other collective must not
    if (strcmp(op, "__syncwarp") != 0 || strcmp(w.op, "__syncwarp") != 0) {
      fprintf(stderr, "[emu] divergent collective: lane %d of warp %d arrived at %s (line %ld), the warp is at %s (line %ld) of kernels_emu.inc\n", f->lane, f->warp, op, site, w.op, w.site);
      abort();
    }
    M.site_mismatches++;
  }
  w.val[g & 1][f->lane] = v;
  if (++w.arrived == 32) {
    w.arrived = 0; w.gen = g + 1; M.collectives++;
    // the other warps get their turn HERE, where this warp is converged: between two collectives its lanes run one
    // after the other without anybody else in between, so that they all read the same values from memory (on the
    // device a converged warp issues such a load once for all lanes)
    if (M.collectives % warp_quantum == 0) yield_outer();
  }
  else { w.blocked++; while (w.gen == g) yield_inner(); w.blocked--; }
  return w.val[g & 1];
}

template <class T> inline uint64_t to_u64(T v) { uint64_t r = 0; static_assert(sizeof(T) <= 8, "size"); memcpy(&r, &v, sizeof(T)); return r; }
template <class T> inline T from_u64(uint64_t r) { T v; memcpy(&v, &r, sizeof(T)); return v; }

}  // namespace emu

#define threadIdx (emu::M.cur->tid)
#define blockIdx (emu::M.cur->bid)
#define blockDim (emu::M.block_dim)
#define gridDim (emu::M.grid_dim)
#define EMU_COLL

template <class T> EMU_COLL T emu__shfl_sync(long site_, T v, int src) { return emu::from_u64<T>(emu::warp_exchange(emu::to_u64(v), site_, "__shfl_sync")[src & 31]); }
template <class T> EMU_COLL T emu__shfl_xor_sync(long site_, T v, int mask) {
  const int lane = emu::M.cur->lane;
  return emu::from_u64<T>(emu::warp_exchange(emu::to_u64(v), site_, "__shfl_xor_sync")[(lane ^ mask) & 31]);
}
EMU_COLL inline unsigned emu__ballot_sync(long site_, int pred) {
  const uint64_t *v = emu::warp_exchange(pred ? 1u : 0u, site_, "__ballot_sync");
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(v[i] & 1u) << i;
  return r;
}
EMU_COLL inline int emu__any_sync(long site_, int pred) {
  const uint64_t *v = emu::warp_exchange(pred ? 1u : 0u, site_, "__any_sync");
  for (int i = 0; i < 32; i++) if (v[i]) return 1;
  return 0;
}
EMU_COLL inline int emu__all_sync(long site_, int pred) {
  const uint64_t *v = emu::warp_exchange(pred ? 1u : 0u, site_, "__all_sync");
  for (int i = 0; i < 32; i++) if (!v[i]) return 0;
  return 1;
}
EMU_COLL inline void emu__syncwarp(long site_) { emu::warp_exchange(0, site_, "__syncwarp"); }
EMU_COLL inline int emu__reduce_add_sync(long site_, int x) {
  const uint64_t *v = emu::warp_exchange(emu::to_u64(x), site_, "__reduce_add_sync");
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r += (unsigned)v[i];
  return (int)r;
}
EMU_COLL inline unsigned emu__reduce_add_sync(long site_, unsigned x) {
  const uint64_t *v = emu::warp_exchange(x, site_, "__reduce_add_sync");
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r += (unsigned)v[i];
  return r;
}
EMU_COLL inline unsigned emu__reduce_or_sync(long site_, unsigned x) {
  const uint64_t *v = emu::warp_exchange(x, site_, "__reduce_or_sync");
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)v[i];
  return r;
}
EMU_COLL inline unsigned emu__reduce_xor_sync(long site_, unsigned x) {
  const uint64_t *v = emu::warp_exchange(x, site_, "__reduce_xor_sync");
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r ^= (unsigned)v[i];
  return r;
}


// the collectives carry the source line they are called from: every lane of a warp must arrive from the same one
#define __shfl_sync(mask, v, src) emu__shfl_sync(__LINE__, (v), (src))
#define __shfl_xor_sync(mask, v, lm) emu__shfl_xor_sync(__LINE__, (v), (lm))
#define __ballot_sync(mask, p) emu__ballot_sync(__LINE__, (p))
#define __any_sync(mask, p) emu__any_sync(__LINE__, (p))
#define __all_sync(mask, p) emu__all_sync(__LINE__, (p))
#define __syncwarp(...) emu__syncwarp(__LINE__)
#define __reduce_add_sync(mask, x) emu__reduce_add_sync(__LINE__, (x))
#define __reduce_or_sync(mask, x) emu__reduce_or_sync(__LINE__, (x))
#define __reduce_xor_sync(mask, x) emu__reduce_xor_sync(__LINE__, (x))
#define EMU_UNIFORM(x) __shfl_sync(0xffffffffu, (x), 0)

inline void __syncthreads() {
  emu::Block &b = emu::M.blocks[emu::M.cur->block];
  const unsigned g = b.gen;
  if (++b.arrived == b.threads) { b.arrived = 0; b.gen = g + 1; }
  else { emu::Warp &w = emu::M.warps[emu::M.cur->warp]; w.blocked++; while (b.gen == g) emu::yield_outer(); w.blocked--; }
}

// ---- memory, atomics, clocks -----------------------------------------------------------------------------------
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T, class U> inline void __stcg(T *p, U v) { *p = (T)v; }
inline void __stcg(int4 *p, int4 v) { *p = v; }
inline void __stcg(int2 *p, int2 v) { *p = v; }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline void __threadfence_block() {}
// a waiting lane gives the other warps a turn -- once the other lanes of its own warp have caught up with it (they
// wait at the collective that follows the wait loop): what they read on the way must not be newer than what this lane read
inline void __nanosleep(unsigned) {
  emu::Warp &w = emu::M.warps[emu::M.cur->warp];
  for (int tries = 0; tries < 64 && w.blocked < w.alive - 1; tries++) emu::yield_inner();
  emu::yield_outer();
}
// the clock is read once per warp between two collectives (a converged warp reads one SM clock value on the device)
inline long long clock64() {
  emu::Warp &w = emu::M.warps[emu::M.cur->warp];
  if (w.clock_gen != w.gen || w.clock_val == 0) { w.clock_val = (emu::M.clock += 64); w.clock_gen = w.gen; }
  return w.clock_val;
}
[[noreturn]] inline void __trap() { fprintf(stderr, "[emu] __trap()\n"); abort(); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }

#define EMU_ATOMIC(NAME, EXPR) \
  template <class T, class U> inline T NAME(T *p, U v_) { const T v = (T)v_; const T old = *p; *p = (EXPR); return old; } \
  template <class T, class U> inline T NAME##_system(T *p, U v_) { return NAME(p, v_); } \
  template <class T, class U> inline T NAME##_block(T *p, U v_) { return NAME(p, v_); }
EMU_ATOMIC(atomicAdd, (T)(old + v))
EMU_ATOMIC(atomicSub, (T)(old - v))
EMU_ATOMIC(atomicMax, old > v ? old : v)
EMU_ATOMIC(atomicMin, old < v ? old : v)
EMU_ATOMIC(atomicOr, (T)(old | v))
EMU_ATOMIC(atomicAnd, (T)(old & v))
EMU_ATOMIC(atomicExch, v)
template <class T, class U, class W> inline T atomicCAS(T *p, U cmp, W v) { const T old = *p; if (old == (T)cmp) *p = (T)v; return old; }
template <class T, class U, class W> inline T atomicCAS_system(T *p, U cmp, W v) { return atomicCAS(p, cmp, v); }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }

// dynamic shared memory of the running block / the block's statically declared __shared__ ints (the kernel source is
// rewritten to call these where it declares them: tests/util.py build_emu)
inline int *emu_dynamic_smem() { return reinterpret_cast<int *>(emu::M.blocks[emu::M.cur->block].smem); }
// a statically declared __shared__ object of the running block, keyed by the line that declares it (zeroed at launch)
inline void *emu_static_shared(long key, size_t bytes) {
  emu::Block &b = emu::M.blocks[emu::M.cur->block];
  for (auto &e : b.statics) if (e.first == key) return e.second.data();
  b.statics.emplace_back(key, std::vector<unsigned long long>((bytes + 7) / 8, 0ull));
  return b.statics.back().second.data();
}
// n-th set bit of mask at or above `base` (offset >= 1), -1 when there is none (PTX fns with a positive offset)
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
  for (unsigned b = base; b < 32; b++) if ((mask >> b) & 1u) { if (--offset == 0) return b; }
  return 0xffffffffu;
}

namespace emu {

// run `entry(arg)` on grid x block threads (block a multiple of 32) with smem_bytes of dynamic shared memory per block
inline void launch(int grid, int block, size_t smem_bytes, void (*entry)(void *), void *arg) {
  const size_t STACK = 512 * 1024;
  const int n = grid * block;
  if (block % 32 != 0 && grid != 1) { fprintf(stderr, "[emu] a block that is not a multiple of 32 threads needs grid 1\n"); abort(); }
  const int n_slots = (n + 31) / 32 * 32;          // lanes that do not exist are fibers that are done from the start
  M = Machine();
  M.fibers.resize(n_slots); M.warps.resize(n_slots / 32); M.blocks.resize(grid);
  M.block_dim = dim3(block, 1, 1); M.grid_dim = dim3(grid, 1, 1);
  M.entry = entry; M.entry_arg = arg; M.live = n;
  char *stacks = (char *)mmap(nullptr, STACK * n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (stacks == MAP_FAILED) { perror("mmap"); abort(); }
  std::vector<std::vector<char>> smem(grid);
  for (int b = 0; b < grid; b++) {
    smem[b].assign(smem_bytes + 64, 0);
    M.blocks[b].threads = block;
    M.blocks[b].smem = (char *)(((uintptr_t)smem[b].data() + 15) & ~(uintptr_t)15);
  }
  for (int i = 0; i < n; i++) {
    Fiber &f = M.fibers[i];
    f.block = i / block; f.tid = uint3{(unsigned)(i % block), 0, 0};
    f.bid = uint3{(unsigned)(block_group > 0 ? f.block % block_group : f.block), 0, 0};
    f.warp = i / 32; f.lane = i % 32;
    f.stack = stacks + STACK * i;
    // initial frame for emu_switch: six callee-saved registers, then the return address; the entry point finds the
    // stack 8 bytes off a 16-byte boundary, as after a call
    uintptr_t top = ((uintptr_t)f.stack + STACK) & ~(uintptr_t)15;
    void **sp = (void **)(top - 8);
    *--sp = (void *)&fiber_main;
    for (int k = 0; k < 6; k++) *--sp = nullptr;
    f.sp = sp;
  }
  for (size_t w = 0; w < M.warps.size(); w++) M.warps[w].first = (int)w * 32;
  for (int i = n; i < n_slots; i++) { M.fibers[i].done = true; M.fibers[i].warp = i / 32; M.fibers[i].lane = i % 32; M.warps[i / 32].alive--; }
  M.cur = &M.fibers[0];
  M.switches++;
  emu_switch(&M.main_sp, M.fibers[0].sp);
  if (M.live != 0) { fprintf(stderr, "[emu] %d threads never finished\n", M.live); abort(); }
  munmap(stacks, STACK * n);
}

}  // namespace emu

#ifdef EMU_DEFINE_MACHINE
namespace emu { Machine M; int lane_step = 1; unsigned warp_quantum = 64; int block_group = 0; }
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");
#endif
