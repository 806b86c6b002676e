// Stub <cuda_runtime.h> for the CPU CTA emulator (tests/emu/emu_*.cpp): lets g++ compile the kernel SOURCES of
// climaatmos.jl_b200/csrc/*.cuh unchanged.  One CTA = NT cooperative FIBERS (ucontext) on one host thread: a fiber runs until it
// reaches a barrier (__syncthreads(), or the per-warp barrier behind the emulated shuffles / __syncwarp()) and then yields to the next
// one, so the emulation is deterministic, needs no kernel-level synchronisation and costs one context switch per thread per barrier
// (the first emulator ran 256 host threads against std::barrier — futex round trips that made the CPU test tier take 5–14 minutes
// depending on the load of the machine).  Exited fibers stop counting towards barriers, as exited CUDA threads do.
// Test infrastructure only.
#pragma once
#include <sys/mman.h>
#include <ucontext.h>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__
#define __align__(x) __attribute__((aligned(x)))

struct uint3_emu { unsigned x, y, z; };
extern thread_local uint3_emu threadIdx, blockIdx;
struct float2 { float x, y; };
struct double2 { double x, y; };

namespace emu {
constexpr int MAXT = 1024;
constexpr size_t STACK = 1 << 20;  // per fiber; mapped lazily
struct Sched {
  int n = 0, cur = 0, ndone = 0;
  ucontext_t main_ctx;
  ucontext_t ctx[MAXT];
  bool done[MAXT];
  uint3_emu tid[MAXT];
  char* stacks = nullptr;
  int cta_arrived = 0;
  unsigned cta_gen = 0;
  int warp_arrived[MAXT / 32], warp_live[MAXT / 32];
  unsigned warp_gen[MAXT / 32];
  alignas(16) unsigned char warp_buf[MAXT / 32][32][16];  // shuffle exchange slots
  const std::function<void()>* body = nullptr;
};
inline Sched& S() {
  static Sched* s = nullptr;  // the emulators drive one CTA at a time from one host thread
  if (!s) {
    s = new Sched;
    s->stacks = (char*)mmap(nullptr, STACK * MAXT, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (s->stacks == (char*)MAP_FAILED) { perror("emu: mmap"); abort(); }
  }
  return *s;
}
inline void switch_to(int nx) {
  Sched& s = S();
  const int me = s.cur;
  s.cur = nx;
  threadIdx = s.tid[nx];
  swapcontext(&s.ctx[me], &s.ctx[nx]);
}
inline int next_live(int from) {
  Sched& s = S();
  int nx = from;
  do { nx = nx + 1 == s.n ? 0 : nx + 1; } while (s.done[nx] && nx != from);
  return nx;
}
inline void yield() {
  Sched& s = S();
  const int nx = next_live(s.cur);
  if (nx != s.cur) switch_to(nx);
}
[[noreturn]] inline void deadlock(const char* what) {
  fprintf(stderr, "emu: deadlock at a %s barrier (a thread left the kernel's barrier sequence?)\n", what);
  abort();
}
inline void cta_barrier() {
  Sched& s = S();
  const unsigned g = s.cta_gen;
  if (++s.cta_arrived == s.n - s.ndone) { s.cta_arrived = 0; ++s.cta_gen; return; }
  for (long spins = 0; s.cta_gen == g; ++spins) {
    if (spins > 50000000L) deadlock("CTA");
    yield();
  }
}
inline void warp_barrier() {
  Sched& s = S();
  const int w = s.cur >> 5;
  const unsigned g = s.warp_gen[w];
  if (++s.warp_arrived[w] == s.warp_live[w]) { s.warp_arrived[w] = 0; ++s.warp_gen[w]; return; }
  for (long spins = 0; s.warp_gen[w] == g; ++spins) {
    if (spins > 50000000L) deadlock("warp");
    yield();
  }
}
inline void fiber_entry() {
  Sched& s = S();
  (*s.body)();
  const int me = s.cur, w = me >> 5;
  s.done[me] = true;
  ++s.ndone;
  --s.warp_live[w];
  // a thread that has left no longer takes part in barriers: release the ones it was the last to be waited for
  if (s.cta_arrived > 0 && s.cta_arrived == s.n - s.ndone) { s.cta_arrived = 0; ++s.cta_gen; }
  if (s.warp_arrived[w] > 0 && s.warp_arrived[w] == s.warp_live[w]) { s.warp_arrived[w] = 0; ++s.warp_gen[w]; }
  if (s.ndone == s.n) { setcontext(&s.main_ctx); abort(); }
  const int nx = next_live(me);
  s.cur = nx;
  threadIdx = s.tid[nx];
  setcontext(&s.ctx[nx]);
  abort();
}
// run ONE CTA of n threads; tid_of(t) = threadIdx of linear thread t (warps are 32 consecutive linear threads)
template <class TID>
inline void run_cta(int n, TID&& tid_of, const std::function<void()>& body) {
  Sched& s = S();
  if (n > MAXT) abort();
  s.n = n; s.ndone = 0; s.cta_arrived = 0; s.body = &body;
  for (int w = 0; w < (n + 31) / 32; ++w) { s.warp_arrived[w] = 0; s.warp_live[w] = n - 32 * w < 32 ? n - 32 * w : 32; }
  for (int t = 0; t < n; ++t) {
    s.done[t] = false;
    s.tid[t] = tid_of(t);
    getcontext(&s.ctx[t]);
    s.ctx[t].uc_stack.ss_sp = s.stacks + (size_t)t * STACK;
    s.ctx[t].uc_stack.ss_size = STACK;
    s.ctx[t].uc_link = nullptr;
    makecontext(&s.ctx[t], (void (*)())fiber_entry, 0);
  }
  s.cur = 0;
  threadIdx = s.tid[0];
  swapcontext(&s.main_ctx, &s.ctx[0]);
}
// the usual 1-D block of n threads
inline void run_cta(int n, const std::function<void()>& body) {
  run_cta(n, [](int t) { return uint3_emu{(unsigned)t, 0, 0}; }, body);
}
}  // namespace emu

inline void __syncthreads() { emu::cta_barrier(); }

// Warp intrinsics (opt in with EMU_WARP_INTRINSICS before including this header): the lanes of a warp meet at the per-warp barrier,
// publish their value and read the source lane's — all shuffles of the kernels are executed by full, converged warps.
#ifdef EMU_WARP_INTRINSICS
template <class T> inline T shfl_emu(T v, int src_lane) {
  emu::Sched& s = emu::S();
  const int w = s.cur >> 5, l = s.cur & 31;
  memcpy(s.warp_buf[w][l], &v, sizeof(T));
  emu::warp_barrier();
  T r;
  memcpy(&r, s.warp_buf[w][src_lane & 31], sizeof(T));
  emu::warp_barrier();
  return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return shfl_emu(v, (emu::S().cur & 31) ^ m); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return shfl_emu(v, src); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d) { const int l = emu::S().cur & 31; return shfl_emu(v, l >= d ? l - d : l); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) { const int l = emu::S().cur & 31; return shfl_emu(v, l + d < 32 ? l + d : l); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline int __any_sync(unsigned, int pred) {
  emu::Sched& s = emu::S();
  const int w = s.cur >> 5, l = s.cur & 31;
  memcpy(s.warp_buf[w][l], &pred, sizeof(int));
  emu::warp_barrier();
  int any = 0;
  for (int k = 0; k < 32; ++k) { int p; memcpy(&p, s.warp_buf[w][k], sizeof(int)); any |= (p != 0); }
  emu::warp_barrier();
  return any;
}
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
#endif
