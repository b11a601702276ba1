// Host accuracy test of log_fast_core (ampe_b200/csrc/log_core.h): the same chain compiled with g++ against logl
// over the arguments the KKS Newton produces (1e-8 .. 1, densely near 1) and over the whole normal range.
#include <cmath>
#include <cstdio>
#include <random>

#include "../../ampe_b200/csrc/log_core.h"

struct Div {
   double operator()(double d) const { return 1.0 / d; }
};
// the device passes rcp_fast: MUFU seed + two Newton steps, <= 1 ulp; emulate a reciprocal that is off by one ulp
struct DivOff {
   double operator()(double d) const { return nextafter(1.0 / d, 2.0 / d); }
};

template <class R>
static void run(const char* name, R r)
{
   static const double C[ampe::LOGC_N] = AMPE_LOG_COEFFS;
   std::mt19937_64 g(7);
   std::uniform_real_distribution<double> u(0.0, 1.0);
   double max_ulp = 0.0, max_abs_near1 = 0.0;
   auto check = [&](double x) {
      const double got = ampe::log_fast_core(x, C, r);
      const long double ref = logl((long double)x);
      const double refd = (double)ref;
      const double ulp = (refd == 0.0) ? 0.0 : fabs(nextafter(refd, 2 * refd) - refd);
      const double e = (refd == 0.0) ? fabs(got) / 1.1e-16 : (double)(fabsl((long double)got - ref) / ulp);
      if (e > max_ulp) max_ulp = e;
      if (fabs(x - 1.0) < 1e-3) max_abs_near1 = fmax(max_abs_near1, (double)fabsl((long double)got - ref));
   };
   for (int i = 0; i < 4000000; i++) {
      check(pow(10.0, -8.0 * u(g)));              // concentrations between 1e-8 and 1
      check(1.0 - pow(10.0, -1.0 - 7.0 * u(g)));  // 1 - c for small c
      check(ldexp(0.5 + 0.5 * u(g), (int)(2040 * u(g)) - 1020));
   }
   check(1.0);
   check(0.5);
   check(2.0);
   check(0.70710678118654752);
   check(0.70710678118654757);
   printf("%s_max_ulp %.4f\n%s_max_abs_near1 %.3e\n", name, max_ulp, name, max_abs_near1);
}

int main()
{
   run("div", Div());
   run("off", DivOff());
   static const double C[ampe::LOGC_N] = AMPE_LOG_COEFFS;
   printf("log1 %.17g\n", ampe::log_fast_core(1.0, C, Div()));
   return 0;
}
