// Stub <cuda_runtime.h> for the CPU CTA emulator (tests/emu/emu_vdiff.cpp): lets g++ compile the kernel SOURCES of
// climaatmos.jl_b200/csrc/*.cuh unchanged.  One CTA = NT host threads that meet at a std::barrier for __syncthreads().
// Test infrastructure only.
#pragma once
#include <barrier>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

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
extern std::barrier<>* g_cta_barrier;
inline void __syncthreads() { g_cta_barrier->arrive_and_wait(); }
struct float2 { float x, y; };
struct double2 { double x, y; };
