// C entry points over the host-side Strategy mirror (QuatIntegrator.h) so that the parity
// tests can drive the unfused Strategy path and the fused path through the same call.
#include <cstring>
#include <string>

#include "QuatIntegrator.h"

static thread_local std::string g_host_err;

extern "C" {

void* ampe_host_create(const ampe_rhs_config* cfg, int use_fused)
{
   try {
      return new ampe_host::QuatIntegrator(*cfg, use_fused != 0);
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return nullptr;
   }
}
void ampe_host_destroy(void* h) { delete static_cast<ampe_host::QuatIntegrator*>(h); }
const char* ampe_host_last_error(void) { return g_host_err.c_str(); }

int ampe_host_reset_ref_phase_concentrations(void* h, const double* cl_ref, const double* ca_ref)
{
   try {
      static_cast<ampe_host::QuatIntegrator*>(h)->resetRefPhaseConcentrations(cl_ref, ca_ref);
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
int ampe_host_set_symmetry_rotations(void* h, const int* const* iqrot)
{
   try {
      static_cast<ampe_host::QuatIntegrator*>(h)->setSymmetryRotations(iqrot);
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
// QuatIntegrator::evaluateRHSFunction(time, y, y_dot, fd_flag)
int ampe_host_evaluate_rhs_function(void* h, double time, const ampe_rhs_fields* y,
                                    const ampe_rhs_fields* y_dot, int fd_flag)
{
   try {
      return static_cast<ampe_host::QuatIntegrator*>(h)->evaluateRHSFunction(time, y, y_dot, fd_flag);
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
}
