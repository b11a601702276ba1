// Arithmetic core of atan_fast (fastmath.cuh), free of CUDA-only constructs so that the host accuracy
// test (tests/cpp/atan_accuracy.cpp) compiles exactly this fma chain with g++.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#define AMPE_HD __host__ __device__ __forceinline__
#else
#define AMPE_HD inline
#endif

namespace ampe {

// atan(x) for finite x, straight-line (candidate replacement of CUDA's atan in the bias-well term of
// computerhsbiaswell, 9.8 % of the Dendrite2D kernel's instructions: profiles/README.md).
// One division for every range:  |x| < tan(pi/8): t = |x|;  |x| > tan(3pi/8): t = -1/|x|, base pi/2;
// otherwise t = (|x|-1)/(|x|+1), base pi/4;  |t| <= tan(pi/8), atan(t) = t + t s P(s), s = t^2, P of
// degree 11 (Chebyshev interpolant of (atan(sqrt s)/sqrt s - 1)/s on [0, tan^2(pi/8)], truncation
// 1.8e-19 relative).  The arithmetic core is atan_fast_core(x, recip): the device version passes
// rcp_fast, the host accuracy test (tests/cpp/atan_accuracy.cpp) a plain division.
template <class RCP>
AMPE_HD double atan_fast_core(double x, RCP recip)
{
   const double ax = fabs(x);
   const bool lo = ax < 0.41421356237309503;  // tan(pi/8)
   const bool hi = ax > 2.4142135623730951;   // tan(3 pi/8)
   const double num = lo ? ax : (hi ? -1.0 : ax - 1.0);
   const double den = lo ? 1.0 : (hi ? ax : ax + 1.0);
   const double bh = lo ? 0.0 : (hi ? 1.57079632679489656e+00 : 7.85398163397448279e-01);
   const double bl = lo ? 0.0 : (hi ? 6.12323399573676604e-17 : 3.06161699786838302e-17);
   const double r = recip(den);
   double t = num * r;
   t = fma(fma(-den, t, num), r, t);  // quotient corrected to the last bit
   const double s = t * t;
   double q = 0.016285756855221028291;
   q = fma(q, s, -0.034570561981427746882);
   q = fma(q, s, 0.045515932206265491693);
   q = fma(q, s, -0.052304542706502445183);
   q = fma(q, s, 0.058789289978347751327);
   q = fma(q, s, -0.066664248857382553335);
   q = fma(q, s, 0.076922963750321423991);
   q = fma(q, s, -0.090909087535008768442);
   q = fma(q, s, 0.11111111105155446565);
   q = fma(q, s, -0.14285714285659827606);
   q = fma(q, s, 0.19999999999999804526);
   q = fma(q, s, -0.33333333333333333217);
   const double a = bh + (t + fma(t * s, q, bl));
   return (x < 0.0) ? -a : a;
}

}  // namespace ampe
