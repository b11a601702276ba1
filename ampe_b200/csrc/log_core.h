// Arithmetic core of log_fast (fastmath.cuh), free of CUDA-only constructs so that the host accuracy test
// (tests/cpp/log_accuracy.cpp) compiles exactly this chain with g++.
//
// log(x) for NORMAL, finite x > 0 (the KKS Newton only takes logarithms of concentrations above the 1e-8
// xlogx extension, calphad.cuh), straight-line: the published fdlibm scheme -- x = 2^k m with
// m in [sqrt(2)/2, sqrt(2)), f = m - 1, s = f / (2 + f), log(m) = f - s (f - R(s^2)) with R the degree-7
// even polynomial Lg1..Lg7 -- with the division by a Newton-refined reciprocal (passed in) and no
// subnormal / infinity / NaN handling.  CUDA's log() spends ~90 instructions per call on this target, 60 of
// them on range checks, 64-bit literal moves and the subnormal path; this is 27 FP64 + 7 integer ones with
// the coefficients read from constant memory (LOGC).  Accuracy: < 1 ulp (host test against logl).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#ifdef __CUDACC__
#define AMPE_HD __host__ __device__ __forceinline__
#else
#define AMPE_HD inline
#endif

namespace ampe {

// coefficient order of the table handed to log_fast_core
enum { LOGC_LN2HI = 0, LOGC_LN2LO, LOGC_LG1, LOGC_LG2, LOGC_LG3, LOGC_LG4, LOGC_LG5, LOGC_LG6, LOGC_LG7, LOGC_N };
#define AMPE_LOG_COEFFS                                                                                         \
   {6.93147180369123816490e-01, 1.90821492927058770002e-10, 6.666666666666735130e-01, 3.999999999940941908e-01, \
    2.857142874366239149e-01,   2.222219843214978396e-01,   1.818357216161805012e-01, 1.531383769920937332e-01, \
    1.479819860511658591e-01}

// hi / lo words of a double and back (device: register aliases; host: memcpy)
struct Words {
   AMPE_HD static int hi(double x)
   {
#ifdef __CUDA_ARCH__
      return __double2hiint(x);
#else
      uint64_t u;
      memcpy(&u, &x, 8);
      return (int)(u >> 32);
#endif
   }
   AMPE_HD static int lo(double x)
   {
#ifdef __CUDA_ARCH__
      return __double2loint(x);
#else
      uint64_t u;
      memcpy(&u, &x, 8);
      return (int)(u & 0xffffffffu);
#endif
   }
   AMPE_HD static double make(int h, int l)
   {
#ifdef __CUDA_ARCH__
      return __hiloint2double(h, l);
#else
      uint64_t u = ((uint64_t)(uint32_t)h << 32) | (uint32_t)l;
      double x;
      memcpy(&x, &u, 8);
      return x;
#endif
   }
};

template <class RCP>
AMPE_HD double log_fast_core(double x, const double* C, RCP recip)
{
   // x = 2^k m, m in [sqrt(2)/2, sqrt(2)): add the distance of sqrt(2)/2's high word to 1.0's, so that the
   // exponent field of the sum is k + 1023 and its mantissa bits give m back
   int hx = Words::hi(x);
   hx += 0x3ff00000 - 0x3fe6a09e;
   const int k = (hx >> 20) - 0x3ff;
   hx = (hx & 0x000fffff) + 0x3fe6a09e;
   const double m = Words::make(hx, Words::lo(x));
   const double dk = (double)k;
   const double f = m - 1.0;
   const double s = f * recip(2.0 + f);
   const double z = s * s;
   const double w = z * z;
   const double t1 = w * fma(w, fma(w, C[LOGC_LG6], C[LOGC_LG4]), C[LOGC_LG2]);
   const double t2 = z * fma(w, fma(w, fma(w, C[LOGC_LG7], C[LOGC_LG5]), C[LOGC_LG3]), C[LOGC_LG1]);
   const double R = t2 + t1;
   const double hfsq = 0.5 * f * f;
   return fma(dk, C[LOGC_LN2HI], -((hfsq - fma(s, hfsq + R, dk * C[LOGC_LN2LO])) - f));
}

}  // namespace ampe
