// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Host vector backend for the product's ImplicitIntegrator template
// (ampe_b200/host/ImplicitIntegrator.h) on top of the CPU oracle: the SAME integrator code that
// drives the device vectors, with the oracle's evaluateRHSFunction / project / normalizeQuat as
// the operations.  The GPU trajectory tests compare the device run with this one; the CPU tests
// check the host logic (Newton, GMRES, BDF coefficients) with it.
#include <cmath>
#include <cstring>
#include <vector>

#include "../ampe_b200/host/ImplicitIntegrator.h"
#include "ctx.h"
#include "oracle.h"
#include "precond.h"

namespace {
using namespace oracle;

struct HVec {
   std::vector<double> comp[4];  // phase, quat, conc, temperature (empty = absent)
};

class OracleOps
{
 public:
   typedef HVec Vec;
   explicit OracleOps(Ctx* c) : d_c(c), d_cfg(c->cfg)
   {
      d_ncell = 1;
      for (int d = 0; d < d_cfg.ndim; d++) d_ncell *= (size_t)d_cfg.n[d];
      d_evolved[0] = d_cfg.with_phase != 0;
      d_evolved[1] = d_cfg.evolve_quat != 0;
      d_evolved[2] = d_cfg.with_concentration != 0;
      d_evolved[3] = d_cfg.with_unsteady_temperature != 0;
      d_length = 0;
      for (int k = 0; k < 4; k++)
         if (d_evolved[k]) d_length += (long long)d_ncell * (k == 1 ? d_cfg.qlen : 1);
   }
   Vec clone(const Vec& y) { return y; }
   void release(Vec& v)
   {
      for (auto& c : v.comp) c.clear();
   }
   void linearSum(double a, const Vec& x, double b, const Vec& y, Vec& z)
   {
      for (int k = 0; k < 4; k++) {
         if (!d_evolved[k]) continue;
         const size_t n = x.comp[k].size();
         for (size_t i = 0; i < n; i++) z.comp[k][i] = a * x.comp[k][i] + b * y.comp[k][i];
      }
   }
   void scale(double a, const Vec& x, Vec& z)
   {
      for (int k = 0; k < 4; k++) {
         if (!d_evolved[k]) continue;
         const size_t n = x.comp[k].size();
         for (size_t i = 0; i < n; i++) z.comp[k][i] = a * x.comp[k][i];
      }
   }
   double wdot(const Vec& x, const Vec& y, const Vec& w)
   {
      double total = 0.0;
      for (int k = 0; k < 4; k++) {
         if (!d_evolved[k]) continue;
         const size_t n = x.comp[k].size();
         long double acc = 0.0L;
         for (size_t i = 0; i < n; i++) {
            const double a = x.comp[k][i] * w.comp[k][i], b = y.comp[k][i] * w.comp[k][i];
            acc += (long double)(a * b);
         }
         total += (double)acc;
      }
      return total;
   }
   long long length() const { return d_length; }
   void errorWeights(const Vec& y, double rtol, double atol, Vec& w)
   {
      for (int k = 0; k < 4; k++) {
         if (!d_evolved[k]) continue;
         const size_t n = y.comp[k].size();
         for (size_t i = 0; i < n; i++) w.comp[k][i] = 1.0 / (rtol * fabs(y.comp[k][i]) + atol);
      }
   }
   static ampe_rhs_fields fields(const Vec& v)
   {
      ampe_rhs_fields f;
      f.phase = v.comp[0].empty() ? nullptr : const_cast<double*>(v.comp[0].data());
      f.quat = v.comp[1].empty() ? nullptr : const_cast<double*>(v.comp[1].data());
      f.conc = v.comp[2].empty() ? nullptr : const_cast<double*>(v.comp[2].data());
      f.temperature = v.comp[3].empty() ? nullptr : const_cast<double*>(v.comp[3].data());
      return f;
   }
   int rhs(double t, const Vec& y, Vec& ydot, int fd_flag)
   {
      ampe_rhs_fields fy = fields(y), fd = fields(ydot);
      return eval(d_c, t, &fy, &fd, fd_flag);
   }
   // QuatIntegrator::applyProjection (QuatIntegrator.cc:3911-3962)
   void applyProjection(double, const Vec& y, Vec& corr, Vec& err)
   {
      for (int k = 0; k < 4; k++)
         if (d_evolved[k]) std::fill(corr.comp[k].begin(), corr.comp[k].end(), 0.0);
      if (d_cfg.qlen > 1 && d_cfg.evolve_quat) {
         const Box& b = d_c->box;
         const int Q = d_cfg.qlen;
         project(b, Q, make_view(const_cast<double*>(y.comp[1].data()), b, -1, 0, Q),
                 make_view(corr.comp[1].data(), b, -1, 0, Q), make_view(err.comp[1].data(), b, -1, 0, Q));
      }
   }
   // QuatModel::normalizeQuat (QuatModel.cc:4237-4262) + resetRefPhaseConcentrations (:5218-5231)
   void postStep(Vec& y)
   {
      if (d_cfg.evolve_quat && d_cfg.qlen > 1) {
         const int Q = d_cfg.qlen;
         double* q = y.comp[1].data();
         for (size_t i = 0; i < d_ncell; i++) {
            double n2 = 0.0;
            for (int m = 0; m < Q; m++) n2 = n2 + q[i + m * d_ncell] * q[i + m * d_ncell];
            const double inv = 1.0 / sqrt(n2);
            for (int m = 0; m < Q; m++) q[i + m * d_ncell] = q[i + m * d_ncell] * inv;
         }
      }
      const bool kks = d_cfg.conc_rhs_form == AMPE_CONC_KKS || d_cfg.conc_rhs_form == AMPE_CONC_EBS;
      if (kks && d_cfg.free_energy == AMPE_FE_CALPHAD) set_ref(d_c, nullptr, nullptr);
   }
   // CVSpgmrPrecondSet / CVSpgmrPrecondSolve (QuatIntegrator.cc:3300-3376, 3666-3771), precond.cc
   bool preconditioned() const { return d_c->precond_cycles > 0; }
   int precondSetup(double, const Vec&, double gamma) { return precond_setup(d_c, gamma, d_c->precond_cycles, d_c->precond_dquatdphi); }
   void precondSolve(const Vec& r, Vec& z)
   {
      ampe_rhs_fields fr = fields(r), fz = fields(z);
      precond_solve(d_c, &fr, &fz);
   }

 private:
   Ctx* d_c;
   ampe_rhs_config d_cfg;
   size_t d_ncell;
   long long d_length;
   bool d_evolved[4];
};

}  // namespace

extern "C" int oracle_integrate_implicit(void* ctx, const ampe_rhs_fields* y, double t0, double dt, int nsteps,
                                         const int* iopt, const double* dopt, double* stats_out)
{
   Ctx* c = (Ctx*)ctx;
   const ampe_rhs_config& cfg = c->cfg;
   size_t ncell = 1;
   for (int d = 0; d < cfg.ndim; d++) ncell *= (size_t)cfg.n[d];
   HVec v;
   double* src[4] = {y->phase, y->quat, y->conc, y->temperature};
   const size_t depth[4] = {1, (size_t)(cfg.qlen > 0 ? cfg.qlen : 1), 1, 1};
   for (int k = 0; k < 4; k++)
      if (src[k]) v.comp[k].assign(src[k], src[k] + ncell * depth[k]);
   ampe_host::ImplicitOptions o;
   if (iopt) o.order = iopt[0], o.max_krylov_dimension = iopt[1], o.max_newton_iterations = iopt[2];
   if (dopt) o.rtol = dopt[0], o.atol = dopt[1], o.newton_tolerance = dopt[2], o.linear_tolerance_factor = dopt[3];
   o.precondition_left = c->precond_left;
   OracleOps ops(c);
   ampe_host::ImplicitIntegrator<OracleOps> integ(ops, o);
   const int rc = integ.advance(v, t0, dt, nsteps);
   const ampe_host::ImplicitStats& st = integ.stats();
   if (stats_out) {
      stats_out[0] = (double)st.steps, stats_out[1] = (double)st.rhs_evals, stats_out[2] = (double)st.jtimes_evals;
      stats_out[3] = (double)st.newton_iterations, stats_out[4] = (double)st.linear_iterations;
      stats_out[5] = (double)st.projections, stats_out[6] = st.last_newton_update;
      stats_out[7] = st.last_linear_residual;
   }
   c->precond_stats[0] = (double)st.precond_setups, c->precond_stats[1] = (double)st.precond_solves;
   for (int k = 0; k < 4; k++)
      if (src[k]) memcpy(src[k], v.comp[k].data(), sizeof(double) * ncell * depth[k]);
   return rc;
}

// advanceTo (variable step, local error test).  iopt[5]: order, max_krylov_dimension, max_newton_iterations,
// max_steps, (reserved); dopt[6]: rtol, atol, newton_tolerance, linear_tolerance_factor, h_min, h_max;
// stats_out[16]: the 8 of oracle_integrate_implicit, then error_test_failures, convergence_failures,
// last_step, smallest_step, largest_step, last_error_estimate, t_reached, 0
extern "C" int oracle_integrate_adaptive(void* ctx, const ampe_rhs_fields* y, double t0, double tend, double h0,
                                         const int* iopt, const double* dopt, double* stats_out)
{
   Ctx* c = (Ctx*)ctx;
   const ampe_rhs_config& cfg = c->cfg;
   size_t ncell = 1;
   for (int d = 0; d < cfg.ndim; d++) ncell *= (size_t)cfg.n[d];
   HVec v;
   double* src[4] = {y->phase, y->quat, y->conc, y->temperature};
   const size_t depth[4] = {1, (size_t)(cfg.qlen > 0 ? cfg.qlen : 1), 1, 1};
   for (int k = 0; k < 4; k++)
      if (src[k]) v.comp[k].assign(src[k], src[k] + ncell * depth[k]);
   ampe_host::ImplicitOptions o;
   if (iopt) {
      o.order = iopt[0], o.max_krylov_dimension = iopt[1], o.max_newton_iterations = iopt[2];
      if (iopt[3] > 0) o.max_steps = iopt[3];
      o.stop_at_tend = !(iopt[4] & 1);
      o.strict_linear_convergence = (iopt[4] & 2) != 0;
      o.scale_newton_tolerance = (iopt[4] & 4) != 0;  // bit 2: CVODE's nonlinear tolerance relative to the error test
      o.hold_step_after_failure = (iopt[4] & 8) != 0;  // bit 3: CVODE's etamax = 1 after a failed attempt
   }
   if (dopt) {
      o.rtol = dopt[0], o.atol = dopt[1], o.newton_tolerance = dopt[2], o.linear_tolerance_factor = dopt[3];
      o.h_min = dopt[4], o.h_max = dopt[5];
   }
   o.precondition_left = c->precond_left;
   OracleOps ops(c);
   ampe_host::ImplicitIntegrator<OracleOps> integ(ops, o);
   const int rc = integ.advanceTo(v, t0, tend, h0);
   const ampe_host::ImplicitStats& st = integ.stats();
   if (stats_out) {
      const double out[16] = {(double)st.steps, (double)st.rhs_evals, (double)st.jtimes_evals,
                              (double)st.newton_iterations, (double)st.linear_iterations, (double)st.projections,
                              st.last_newton_update, st.last_linear_residual, (double)st.error_test_failures,
                              (double)st.convergence_failures, st.last_step, st.smallest_step, st.largest_step,
                              st.last_error_estimate, st.t_reached, 0.0};
      memcpy(stats_out, out, sizeof(out));
   }
   c->precond_stats[0] = (double)st.precond_setups, c->precond_stats[1] = (double)st.precond_solves;
   for (int k = 0; k < 4; k++)
      if (src[k]) memcpy(src[k], v.comp[k].data(), sizeof(double) * ncell * depth[k]);
   return rc;
}
