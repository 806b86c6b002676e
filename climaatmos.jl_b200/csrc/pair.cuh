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

// ---- packed transcendental functions for the dry thermodynamics (Π = (p/p0)^κ = exp(κ ln(p/p0))).
// Float32: Cephes-style range reduction per lane (a few integer instructions) and the polynomials as packed FFMA2 — about 45
// instructions for both lanes of log + exp instead of ≈140 for two logf + two expf calls.  Arguments are normal positive numbers
// (p/p0 ∈ [1e-5, 2]) for logp and |x| < 80 for expp; accuracy ≈ 1 ulp, like the CUDA library functions they replace.
// Float64 forwards to log/exp.
__device__ __forceinline__ P2<float> logp(P2<float> x) {
  const int ilo = __float_as_int(x.lo()), ihi = __float_as_int(x.hi());
  const int elo = (ilo - 0x3f3504f3) >> 23, ehi = (ihi - 0x3f3504f3) >> 23;  // x = m·2^e, m ∈ [√½, √2)
  const P2<float> f = P2<float>(__int_as_float(ilo - (elo << 23)), __int_as_float(ihi - (ehi << 23))) - 1.0f;
  const P2<float> e((float)elo, (float)ehi);
  const P2<float> z = f * f;
  P2<float> y = fma2(f, P2<float>(7.0376836292e-2f), P2<float>(-1.1514610310e-1f));
  y = fma2(y, f, P2<float>(1.1676998740e-1f));
  y = fma2(y, f, P2<float>(-1.2420140846e-1f));
  y = fma2(y, f, P2<float>(1.4249322787e-1f));
  y = fma2(y, f, P2<float>(-1.6668057665e-1f));
  y = fma2(y, f, P2<float>(2.0000714765e-1f));
  y = fma2(y, f, P2<float>(-2.4999993993e-1f));
  y = fma2(y, f, P2<float>(3.3333331174e-1f));
  y = (y * f) * z;
  y = fma2(e, P2<float>(-2.12194440e-4f), y);
  y = fma2(z, P2<float>(-0.5f), y);
  return fma2(e, P2<float>(0.693359375f), f + y);
}
__device__ __forceinline__ P2<float> expp(P2<float> x) {
  const P2<float> magic(12582912.0f);  // 1.5·2^23: adding it rounds to the nearest integer
  const P2<float> t = fma2(x, P2<float>(1.44269504088896341f), magic);
  const P2<float> n = t - magic;
  P2<float> r = fma2(n, P2<float>(-0.693359375f), x);
  r = fma2(n, P2<float>(2.12194440e-4f), r);
  P2<float> y = fma2(r, P2<float>(1.9875691500e-4f), P2<float>(1.3981999507e-3f));
  y = fma2(y, r, P2<float>(8.3334519073e-3f));
  y = fma2(y, r, P2<float>(4.1665795894e-2f));
  y = fma2(y, r, P2<float>(1.6666665459e-1f));
  y = fma2(y, r, P2<float>(5.0000001201e-1f));
  y = fma2(y, r * r, r) + 1.0f;
  const int nlo = __float_as_int(t.lo()) - 0x4b400000, nhi = __float_as_int(t.hi()) - 0x4b400000;
  return P2<float>(__int_as_float(__float_as_int(y.lo()) + (nlo << 23)), __int_as_float(__float_as_int(y.hi()) + (nhi << 23)));
}
__device__ __forceinline__ P2<double> logp(P2<double> x) { return P2<double>(log(x.x), log(x.y)); }
__device__ __forceinline__ P2<double> expp(P2<double> x) { return P2<double>(exp(x.x), exp(x.y)); }

}  // namespace b200
