// Accuracy of atan_fast_core (ampe_b200/csrc/fastmath.cuh) on the host: same fma chain as the device
// function, the reciprocal replaced by a division.  Reference: atanl (x87 80-bit long double).
// Prints "max_ulp <v>", "max_rel <v>", "worst_x <v>".  Built and run by tests/test_fastmath_host.py.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>

#include "../../ampe_b200/csrc/atan_core.h"
namespace ampe_host_math = ampe;

static double ulp_of(double v)
{
   v = std::fabs(v);
   double n = std::nextafter(v, INFINITY);
   return n - v;
}

int main()
{
   struct Div {
      double operator()(double d) const { return 1.0 / d; }
   };
   double max_ulp = 0.0, max_rel = 0.0, worst = 0.0;
   uint64_t state = 0x9E3779B97F4A7C15ull;
   auto rnd = [&]() {
      state ^= state << 13, state ^= state >> 7, state ^= state << 17;
      return (double)(state >> 11) / 9007199254740992.0;
   };
   auto check = [&](double x) {
      const double got = ampe_host_math::atan_fast_core(x, Div());
      const long double want = atanl((long double)x);
      const double err = (double)fabsl((long double)got - want);
      const double u = err / ulp_of((double)want);
      const double rel = (want != 0.0L) ? err / (double)fabsl(want) : err;
      if (u > max_ulp) max_ulp = u, worst = x;
      if (rel > max_rel) max_rel = rel;
   };
   // dense log-uniform sweep over 1e-12 .. 1e12, both signs, plus the range boundaries
   for (int i = 0; i < 4000000; i++) {
      const double e = -12.0 + 24.0 * rnd();
      const double x = std::pow(10.0, e) * ((i & 1) ? -1.0 : 1.0);
      check(x);
   }
   const double edges[] = {0.41421356237309503, 2.4142135623730951, 1.0, 0.0};
   for (double c : edges)
      for (int k = -2000; k <= 2000; k++) {
         double x = c;
         for (int j = 0; j < (k < 0 ? -k : k); j++) x = std::nextafter(x, k < 0 ? -INFINITY : INFINITY);
         check(x);
         check(-x);
      }
   check(1e300);
   check(-1e300);
   check(5e-324);
   std::printf("max_ulp %.4f\nmax_rel %.4e\nworst_x %.17g\n", max_ulp, max_rel, worst);
   std::printf("atan0 %.17g\natan_neg0_signbit %d\n", ampe_host_math::atan_fast_core(0.0, Div()),
               (int)std::signbit(ampe_host_math::atan_fast_core(-0.0, Div())));
   return 0;
}
