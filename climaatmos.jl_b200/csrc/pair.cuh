// pair.cuh — two-lane packed arithmetic for the Blackwell FFMA2 / FMUL2 / FADD2 pipes.
//
// sm_100 executes fma/mul/add.rn.f32x2 on a 64-bit register pair in ONE issue slot (SASS FFMA2/FMUL2/FADD2),
// and accepts a scalar register broadcast to both lanes as an operand.  The explicit-tendency kernels are
// FP32-issue-bound (ncu: 61 % issue-active, 51 % of the instructions are FFMA/FMUL/FADD), so they keep the four
// GLL nodes of a row as two packed pairs and do all 4×4 contractions and pointwise algebra on pairs.
// Float64 has no packed form: P2<double> is a plain two-member struct with the same interface, so the kernels
// are written once.  Each lane follows IEEE round-to-nearest exactly like the scalar instruction.
#pragma once
#include <cuda_runtime.h>

namespace b200 {

template <class FT> struct P2;

template <>
struct P2<float> {
  unsigned long long v;
  __device__ __forceinline__ P2() {}
  __device__ __forceinline__ P2(float lo, float hi) { asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); }
  __device__ __forceinline__ explicit P2(float s) { asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(s)); }
  __device__ __forceinline__ float lo() const { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
  __device__ __forceinline__ float hi() const { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
};
__device__ __forceinline__ P2<float> operator+(P2<float> a, P2<float> b) { P2<float> r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2<float> operator-(P2<float> a, P2<float> b) { P2<float> r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2<float> operator*(P2<float> a, P2<float> b) { P2<float> r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ P2<float> fma2(P2<float> a, P2<float> b, P2<float> c) { P2<float> r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ P2<float> operator-(P2<float> a) { P2<float> r; r.v = a.v ^ 0x8000000080000000ull; return r; }
__device__ __forceinline__ P2<float> ldpair(const float* p) { P2<float> r; r.v = *reinterpret_cast<const unsigned long long*>(p); return r; }

template <>
struct P2<double> {
  double x, y;
  __device__ __forceinline__ P2() {}
  __device__ __forceinline__ P2(double lo, double hi) : x(lo), y(hi) {}
  __device__ __forceinline__ explicit P2(double s) : x(s), y(s) {}
  __device__ __forceinline__ double lo() const { return x; }
  __device__ __forceinline__ double hi() const { return y; }
};
__device__ __forceinline__ P2<double> operator+(P2<double> a, P2<double> b) { return P2<double>(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ P2<double> operator-(P2<double> a, P2<double> b) { return P2<double>(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ P2<double> operator*(P2<double> a, P2<double> b) { return P2<double>(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ P2<double> fma2(P2<double> a, P2<double> b, P2<double> c) { return P2<double>(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ P2<double> operator-(P2<double> a) { return P2<double>(-a.x, -a.y); }
__device__ __forceinline__ P2<double> ldpair(const double* p) { return P2<double>(p[0], p[1]); }

// scalar-broadcast forms (ptxas folds the broadcast into the .F32 operand modifier)
template <class FT> __device__ __forceinline__ P2<FT> operator*(P2<FT> a, FT s) { return a * P2<FT>(s); }
template <class FT> __device__ __forceinline__ P2<FT> operator*(FT s, P2<FT> a) { return a * P2<FT>(s); }
template <class FT> __device__ __forceinline__ P2<FT> operator+(P2<FT> a, FT s) { return a + P2<FT>(s); }
template <class FT> __device__ __forceinline__ P2<FT> operator-(P2<FT> a, FT s) { return a - P2<FT>(s); }
template <class FT> __device__ __forceinline__ P2<FT> fma2(P2<FT> a, FT s, P2<FT> c) { return fma2(a, P2<FT>(s), c); }
template <class FT> __device__ __forceinline__ P2<FT> fma2(FT s, P2<FT> a, P2<FT> c) { return fma2(a, P2<FT>(s), c); }

}  // namespace b200
