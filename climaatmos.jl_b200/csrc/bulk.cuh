// bulk.cuh — 1-D bulk asynchronous copies (the TMA engine without a tensor map) and the mbarrier they complete on.
//
// Every element × component slab of the VIJFH layout is ONE contiguous, 16-byte-aligned block (16·Nv·sizeof(FT) bytes at a
// multiple of 64·Nv bytes), so the persistent kernels (kernels_tma.cuh) stage the inputs of the NEXT element with
// `cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes` (SASS UBLKCP) while the current element is being
// processed: one elected thread arms the stage's mbarrier with the byte count (`mbarrier.arrive.expect_tx`) and issues the
// copies, all threads wait on the phase parity (`mbarrier.try_wait.parity`).  No thread holds a load in flight, so the global
// latency never shows up as a long-scoreboard stall.
//
// The CPU CTA emulator (tests/emu, g++) gets a functional stand-in: the issuing host thread copies synchronously and the
// barrier word counts completed phases.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDACC__
#include <atomic>
#include <cstring>
#include <thread>
#endif

namespace b200 {

typedef unsigned long long mbar_t;

#ifdef __CUDACC__
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(mbar_t* bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy writes (stage reuse)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(mbar_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global → shared bulk copy; `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, mbar_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(mbar_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#else  // ---- CPU CTA emulator: synchronous copies, the barrier word = (completed phases << 32) | pending bytes
inline std::atomic<unsigned long long>& mbar_atomic(mbar_t* bar) { return *reinterpret_cast<std::atomic<unsigned long long>*>(bar); }
inline void mbar_init(mbar_t* bar, unsigned) { mbar_atomic(bar).store(0); }
inline void mbar_fence_init() {}
inline void fence_proxy_async() {}
inline void mbar_expect_tx(mbar_t* bar, unsigned bytes) { mbar_atomic(bar).fetch_add(bytes); }
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, mbar_t* bar) {
  memcpy(dst, src, bytes);
  unsigned long long old = mbar_atomic(bar).fetch_sub(bytes);
  if ((old & 0xffffffffull) == bytes) mbar_atomic(bar).fetch_add(1ull << 32);  // last byte of the phase landed
}
inline void mbar_wait(mbar_t* bar, unsigned parity) {
  while ((((mbar_atomic(bar).load() >> 32) & 1u) == parity)) std::this_thread::yield();
}
#endif

}  // namespace b200
