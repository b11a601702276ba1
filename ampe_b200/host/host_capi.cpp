// C entry points over the host-side Strategy mirror (QuatIntegrator.h) so that the parity
// tests can drive the unfused Strategy path and the fused path through the same call.
#include <cstring>
#include <string>

#include "FieldsInitializer.h"
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
// slab ranks: ghost-plane exchange of the integrator's context and the sum reduction of its vector operations
int ampe_host_halo_export(void* h, void* handle)
{
   try {
      static_cast<ampe_host::QuatIntegrator*>(h)->haloExport(handle);
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
int ampe_host_halo_connect(void* h, const void* handle_prev, const void* handle_next)
{
   try {
      static_cast<ampe_host::QuatIntegrator*>(h)->haloConnect(handle_prev, handle_next);
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
void ampe_host_set_sum_reduction(void* h, double (*fn)(double, void*), void* user)
{
   static_cast<ampe_host::QuatIntegrator*>(h)->setSumReduction(fn, user);
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
// options: {order, max_krylov_dimension, max_newton_iterations} ints, {rtol, atol, newton_tolerance,
// linear_tolerance_factor} doubles (NULL = AMPE's defaults); stats_out[8]: steps, rhs_evals, jtimes_evals,
// newton_iterations, linear_iterations, projections, last_newton_update, last_linear_residual
int ampe_host_integrate_implicit(void* h, const ampe_rhs_fields* y, double t0, double dt, int nsteps,
                                 const int* iopt, const double* dopt, double* stats_out)
{
   try {
      ampe_host::ImplicitOptions o;
      if (iopt) o.order = iopt[0], o.max_krylov_dimension = iopt[1], o.max_newton_iterations = iopt[2];
      if (dopt) o.rtol = dopt[0], o.atol = dopt[1], o.newton_tolerance = dopt[2], o.linear_tolerance_factor = dopt[3];
      ampe_host::ImplicitStats st;
      const int rc = static_cast<ampe_host::QuatIntegrator*>(h)->integrateImplicit(y, t0, dt, nsteps, o, &st);
      if (stats_out) {
         stats_out[0] = (double)st.steps, stats_out[1] = (double)st.rhs_evals, stats_out[2] = (double)st.jtimes_evals;
         stats_out[3] = (double)st.newton_iterations, stats_out[4] = (double)st.linear_iterations;
         stats_out[5] = (double)st.projections, stats_out[6] = st.last_newton_update;
         stats_out[7] = st.last_linear_residual;
      }
      if (rc != 0) g_host_err = "integrateImplicit: " + std::string(rc == ampe_host::IMPLICIT_ENEWTON ? "Newton did not converge" : "failed");
      return rc;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
// ImplicitIntegrator::advanceTo on device vectors.  iopt[5]: order, max_krylov_dimension,
// max_newton_iterations, max_steps, 1 = do not shorten the last step (stop_at_tend off); dopt[6]: rtol, atol, newton_tolerance, linear_tolerance_factor,
// h_min, h_max; stats_out[16]: the 8 of ampe_host_integrate_implicit, then error_test_failures,
// convergence_failures, last_step, smallest_step, largest_step, last_error_estimate, t_reached, 0.
// Returns 0 or an IMPLICIT_E* code (-20 .. -24); -1 with ampe_host_last_error() for everything else.
int ampe_host_integrate_adaptive(void* h, const ampe_rhs_fields* y, double t0, double tend, double h0,
                                 const int* iopt, const double* dopt, double* stats_out)
{
   try {
      ampe_host::ImplicitOptions o;
      if (iopt) {
         o.order = iopt[0], o.max_krylov_dimension = iopt[1], o.max_newton_iterations = iopt[2];
         if (iopt[3] > 0) o.max_steps = iopt[3];
         o.stop_at_tend = !(iopt[4] & 1);  // bit 0: return after the first step at or beyond tend (CV_ONE_STEP loop)
         o.strict_linear_convergence = (iopt[4] & 2) != 0;  // bit 1: CVODE's rule for unconverged linear solves
         o.scale_newton_tolerance = (iopt[4] & 4) != 0;  // bit 2: CVODE's nonlinear tolerance relative to the error test
         o.hold_step_after_failure = (iopt[4] & 8) != 0;  // bit 3: CVODE's etamax = 1 after a failed attempt
      }
      if (dopt) {
         o.rtol = dopt[0], o.atol = dopt[1], o.newton_tolerance = dopt[2], o.linear_tolerance_factor = dopt[3];
         o.h_min = dopt[4], o.h_max = dopt[5];
      }
      ampe_host::ImplicitStats st;
      const int rc = static_cast<ampe_host::QuatIntegrator*>(h)->integrateAdaptive(y, t0, tend, h0, o, &st);
      if (stats_out) {
         const double out[16] = {(double)st.steps, (double)st.rhs_evals, (double)st.jtimes_evals,
                                 (double)st.newton_iterations, (double)st.linear_iterations, (double)st.projections,
                                 st.last_newton_update, st.last_linear_residual, (double)st.error_test_failures,
                                 (double)st.convergence_failures, st.last_step, st.smallest_step, st.largest_step,
                                 st.last_error_estimate, st.t_reached, 0.0};
         memcpy(stats_out, out, sizeof(out));
      }
      if (rc != 0) g_host_err = "integrateAdaptive: code " + std::to_string(rc);
      return rc;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
// ---- block preconditioners ------------------------------------------------
// QuatIntegrator::setupPreconditioners; ncycles V-cycles per block solve, 0 = preconditioner off
int ampe_host_set_preconditioner(void* h, int ncycles, int precond_has_dquatdphi, int precondition_left)
{
   try {
      static_cast<ampe_host::QuatIntegrator*>(h)->setupPreconditioners(ncycles, precond_has_dquatdphi != 0,
                                                                       precondition_left != 0);
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
// QuatIntegrator::CVSpgmrPrecondSet(t, y, ..., gamma)
int ampe_host_precond_set(void* h, double t, const ampe_rhs_fields* y, double gamma)
{
   try {
      return static_cast<ampe_host::QuatIntegrator*>(h)->CVSpgmrPrecondSet(t, y, gamma);
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
// QuatIntegrator::CVSpgmrPrecondSolve(..., r, z, ...)
int ampe_host_precond_solve(void* h, const ampe_rhs_fields* r, const ampe_rhs_fields* z)
{
   try {
      const int rc = static_cast<ampe_host::QuatIntegrator*>(h)->CVSpgmrPrecondSolve(r, z);
      cudaDeviceSynchronize();
      return rc;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
// QuatSysSolver::multiplyDQuatDPhiBlock: out (depth qlen, ghost 0, device) = [dF_q/dphi] phase
int ampe_host_precond_dquatdphi(void* h, const double* phase, double* out)
{
   try {
      static_cast<ampe_host::QuatIntegrator*>(h)->multiplyDQuatDPhiBlock(phase, out);
      cudaDeviceSynchronize();
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
// borrowed ampe_mg handle of a block (0 phase, 1 quaternion, 2 composition, 3 temperature) or NULL
void* ampe_host_precond_level_solver(void* h, int block)
{
   return static_cast<ampe_host::QuatIntegrator*>(h)->preconditionerLevelSolver(block);
}
void ampe_host_precond_stats(void* h, double* out2)
{
   out2[0] = (double)static_cast<ampe_host::QuatIntegrator*>(h)->precondSetups();
   out2[1] = (double)static_cast<ampe_host::QuatIntegrator*>(h)->precondSolves();
}
// ---- initial conditions ----------------------------------------------------
// FieldsInitializer::initializeLevelFromData(level, init_data_filename, slice_index): fills the HOST arrays
// of y (ghost 0, this rank's slab) from a NetCDF classic file; no GPU involved.  read_mask bits: 1 phase,
// 2 temperature, 4 quaternion, 8 concentration (setFieldsToRead).
int ampe_host_read_initial_conditions(const char* filename, const ampe_rhs_config* cfg, int slice_index,
                                      int read_mask, const ampe_rhs_fields* y_host)
{
   try {
      if (!filename || !cfg || !y_host) throw std::runtime_error("ampe_host_read_initial_conditions: NULL argument");
      ampe_host::FieldsInitializer init(*cfg);
      init.setFieldsToRead(read_mask & 1, read_mask & 2, read_mask & 4, read_mask & 8);
      init.initializeLevelFromData(filename, slice_index, y_host);
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
// NetCDF-4 container inspection (host/NetCDF4File.h): shape of a variable / a whole variable of rank <= 3 as doubles.
// Used by the tests to hold the HDF5 subset reader against files the initial-condition path never sees (any float dataset).
int ampe_host_hdf5_var_shape(const char* filename, const char* name, int* rank, long long* shape8)
{
   try {
      if (!filename || !name || !rank || !shape8) throw std::runtime_error("ampe_host_hdf5_var_shape: NULL argument");
      ampe_host::NetCDF4File f(filename);
      const std::vector<size_t> sh = f.shape(name);
      if (sh.size() > 8) throw std::runtime_error("more than eight dimensions");
      *rank = (int)sh.size();
      for (size_t d = 0; d < sh.size(); d++) shape8[d] = (long long)sh[d];
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
int ampe_host_hdf5_read_var(const char* filename, const char* name, double* out)
{
   try {
      if (!filename || !name || !out) throw std::runtime_error("ampe_host_hdf5_read_var: NULL argument");
      ampe_host::NetCDF4File f(filename);
      f.getAll(name, out);
      return 0;
   } catch (const std::exception& e) {
      g_host_err = e.what();
      return -1;
   }
}
}
