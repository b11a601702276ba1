// TEST INFRASTRUCTURE ONLY (see oracle.h).  C entry points of the EXTENDED-PRECISION build of the restatement
// (oracle/Makefile target liboracle_ld.so): the same sources with every `double` replaced by `long double`
// (x87 80-bit, 64-bit mantissa) by sed at build time, ABI header included.  It is the arbiter for outputs whose
// conditioning makes two correct fp64 evaluations differ by more than 1e-12 (the CALPHAD composition RHS:
// differences of Newton-solved concentrations): tests assert |gpu - ld| <= |oracle - ld| (+ 1e-12 scale).
#include "ctx.h"
#include "oracle.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

extern "C" {
void* oracle_ld_create(const ampe_rhs_config* cfg) { return create(*cfg); }
void oracle_ld_destroy(void* c) { destroy((Ctx*)c); }
void oracle_ld_set_ref(void* c, const double* cl, const double* ca) { set_ref((Ctx*)c, cl, ca); }
void oracle_ld_set_rotations(void* c, const int* const* iq) { set_rotations((Ctx*)c, iq); }
int oracle_ld_eval(void* c, double t, const ampe_rhs_fields* y, const ampe_rhs_fields* ydot, int fd_flag)
{
   return eval((Ctx*)c, t, y, ydot, fd_flag);
}
void oracle_ld_get_phase_concentrations(void* c, double* cl, double* ca)
{
   get_phase_concentrations((Ctx*)c, cl, ca);
}
int oracle_ld_sizeof_real(void) { return (int)sizeof(double); }
void oracle_ld_set_num_threads(int n)
{
#ifdef _OPENMP
   if (n > 0) omp_set_num_threads(n);
#else
   (void)n;
#endif
}
}
