// Minimal implicit time integrator around the RHS evaluation: fixed-step
// BDF1 / BDF2 with an inexact Newton iteration whose linear systems are solved by a matrix-free
// scaled GMRES, shaped like the CVODE configuration AMPE uses
//   (QuatIntegrator::setSundialsOptions, QuatIntegrator.cc:1569-1594: BDF, max_order 2, SPGMR with
//    max_krylov_dimension 5, scalar rtol/atol, projection function, JTimes RHS function;
//    samrai/CVODESolver.cc:157-247):
//   * residual evaluations call evaluateRHSFunction(t, y, ydot, fd_flag = 0), Jacobian-vector
//     products the difference quotient  J v ~ [f(y + sigma v) - f(y)] / sigma,  sigma = 1/||v||_WRMS,
//     with fd_flag = 1 (CVODEJTimesRHSFuncEval, CVODESolver.h:1186-1195: lagged face coefficients,
//     QuatIntegrator.cc:3183-3189);
//   * norms are CVODE's weighted RMS norms with w = 1/(rtol |y_n| + atol);
//   * after the nonlinear solve the projection hook (QuatIntegrator::applyProjection,
//     QuatIntegrator.cc:3911-3962) puts the quaternions back on the unit sphere, then the
//     post-step work of QuatModel::Advance (normalizeQuat, resetRefPhaseConcentrations).
//   * optional RIGHT preconditioning by the backend's block preconditioner (the
//     reference's CVSpgmrPrecondSet / CVSpgmrPrecondSolve, QuatIntegrator.cc:3300-3771): set up after
//     every fd_flag = 0 residual evaluation, applied once per Krylov vector and once to the solution.
// advance() takes fixed steps and returns a Newton failure to the caller; advanceTo() adds CVODE's local error
// test and step-size controller (variable-coefficient BDF2, no order selection).  CVODE's Nordsieck history,
// order selection and stability-limit detection are not restated (SUNDIALS is not in the reference tree).
//
// The algorithm is a template over the vector backend so that the same code drives the device
// vectors (DeviceOps in QuatIntegrator.h, everything through the C ABI of libampe_b200.so) and a
// host backend used by the CPU tests of the host logic.
#pragma once
#include <cmath>
#include <vector>

namespace ampe_host {

struct ImplicitOptions {
   int order = 2;                          // max_order (QuatIntegrator.cc:291): 1 or 2
   double rtol = 3.e-6, atol = 3.e-4;      // QuatIntegrator.cc:285-289
   int max_krylov_dimension = 5;           // QuatIntegrator.cc:300-301
   int max_newton_iterations = 3;          // CVODE NLS_MAXCOR
   double newton_tolerance = 0.1;          // CVODE nlscoef
   double linear_tolerance_factor = 0.05;  // CVODE eplifac
   // side of the preconditioner: false = right (the stopping test measures the true linear residual), true = left
   // like the reference's setPreconditioningType(PREC_LEFT) (QuatIntegrator.cc:1583): GMRES on P A x = P b, the
   // stopping test then measures the preconditioned residual
   bool precondition_left = false;
   // advanceTo only (variable step size, CVODE's controller constants, cvode_impl.h / cvode.c)
   double h_min = 0.0, h_max = 0.0;        // 0 = unbounded (CVodeSetMinStep / CVodeSetMaxStep)
   long max_steps = 500;                   // CVODE mxstep
   int max_error_test_failures = 7;        // MXNEF
   int max_convergence_failures = 10;      // MXNCF
   // true: the last step is shortened to land on tend (CVodeSetStopTime).  false: the controller's steps are left
   // alone and advanceTo returns after the first step that reaches or passes tend -- what AMPE's run loop sees
   // (QuatIntegrator::Advance takes ONE internal CVODE step per call, CV_ONE_STEP; output intervals are tested against
   // the time the step landed on, so an output at "t = 0.01" happens a fraction of a step later)
   bool stop_at_tend = true;
   // CVODE's rule for a linear solve that ran out of Krylov vectors (cvLsSolve): a residual that was only REDUCED is
   // accepted on the first Newton iteration and is a recoverable convergence failure afterwards (step retried with
   // h / 4, preconditioner set up again); a residual that was not reduced always is.  false (the default, what every
   // committed deck result was produced with): the update is taken and the Newton test decides alone.
   bool strict_linear_convergence = false;
   // CVODE defers step growth after a failed attempt: a failed Newton iteration or error test sets etamax = 1
   // (cvHandleNFlag, cvDoErrorTest), so the step that finally succeeds keeps its size once (cvPrepareNextStep:
   // "if etamax = 1, defer step size or order changes") before growth up to ETAMX is allowed again (cvCompleteStep).
   // false (the default, what every committed deck result was produced with): the controller may grow the step at once.
   bool hold_step_after_failure = false;
   // CVODE ties the nonlinear tolerance to the error test: dcon = del min(1, crate) / tq[4] <= 1 with tq[4] = nlscoef / tq[2],
   // tq[2] the coefficient that turns the accumulated correction into the local error estimate (cvSetTqBDF) -- the Newton
   // error may be nlscoef of the ALLOWED LOCAL ERROR, not nlscoef in the WRMS norm -- and epslin = eplifac tq[4].  true: the
   // same rule with this integrator's own error coefficient (newton_tolerance / cerr: 0.2 on the first step, 0.3 for BDF1,
   // about 0.55 for BDF2 at constant step).  false (the default every committed deck number was produced with): del <= nlscoef.
   bool scale_newton_tolerance = false;
};

struct ImplicitStats {
   long steps = 0, rhs_evals = 0, jtimes_evals = 0, newton_iterations = 0, linear_iterations = 0,
        projections = 0, precond_setups = 0, precond_solves = 0;
   double last_newton_update = 0.0;  // WRMS norm of the last Newton correction
   double last_linear_residual = 0.0;
   // advanceTo
   long error_test_failures = 0, convergence_failures = 0;
   double last_step = 0.0, smallest_step = 0.0, largest_step = 0.0, last_error_estimate = 0.0, t_reached = 0.0;
};

enum {
   IMPLICIT_OK = 0,
   IMPLICIT_EINVAL = -1,
   IMPLICIT_ENEWTON = -20,
   IMPLICIT_ERHS = -21,
   IMPLICIT_ETOOMUCHWORK = -22,  // CV_TOO_MUCH_WORK: max_steps taken before tend
   IMPLICIT_EERRTEST = -23,      // CV_ERR_FAILURE: error test failed too often or h = h_min
   IMPLICIT_ECONV = -24          // CV_CONV_FAILURE: Newton failed too often or h = h_min
};

// Ops concept:
//   typedef ... Vec;
//   Vec  clone(const Vec& y)                 new vector, every component (evolved or not) copied
//   void release(Vec& v)
//   void linearSum(double a, const Vec& x, double b, const Vec& y, Vec& z)   evolved components
//   void scale(double a, const Vec& x, Vec& z)
//   double wdot(const Vec& x, const Vec& y, const Vec& w)                     sum (w x)(w y)
//   long long length()                                                        evolved unknowns
//   void errorWeights(const Vec& y, double rtol, double atol, Vec& w)
//   int  rhs(double t, const Vec& y, Vec& ydot, int fd_flag)                  0 = success
//   void applyProjection(double t, const Vec& y, Vec& corr, Vec& err)
//   void postStep(Vec& y)
//   bool preconditioned() const                                               false: the two below are not called
//   int  precondSetup(double t, const Vec& y, double gamma)                   coefficients frozen at y; 0 = ok
//   void precondSolve(const Vec& r, Vec& z)                                   z ~ (I - gamma J_blockdiag)^-1 r
template <class Ops>
class ImplicitIntegrator
{
 public:
   typedef typename Ops::Vec Vec;

   ImplicitIntegrator(Ops& ops, const ImplicitOptions& opt) : d_ops(ops), d_opt(opt) {}

   const ImplicitStats& stats() const { return d_stats; }

   // nsteps steps of size h from t0; y is updated in place
   int advance(Vec& y, double t0, double h, int nsteps)
   {
      if (!(h > 0.0) || nsteps < 0 || (d_opt.order != 1 && d_opt.order != 2) || d_opt.max_krylov_dimension < 1 ||
          d_opt.max_newton_iterations < 1)
         return IMPLICIT_EINVAL;
      const int m = d_opt.max_krylov_dimension;
      Vec yprev = d_ops.clone(y), psi = d_ops.clone(y), ycur = d_ops.clone(y), fy = d_ops.clone(y),
          ewt = d_ops.clone(y), res = d_ops.clone(y), delta = d_ops.clone(y), acor = d_ops.clone(y),
          ytmp = d_ops.clone(y), wk = d_ops.clone(y);
      d_pv = d_ops.clone(y);  // P v of the current Krylov vector
      std::vector<Vec> V;
      for (int j = 0; j <= m; j++) V.push_back(d_ops.clone(y));
      int rc = IMPLICIT_OK;
      double t = t0;
      for (int n = 0; n < nsteps && rc == IMPLICIT_OK; n++) {
         const bool bdf2 = d_opt.order == 2 && n > 0;
         // y_{n+1} = psi + gamma f(y_{n+1})
         const double gamma = bdf2 ? (2.0 / 3.0) * h : h;
         if (bdf2)
            d_ops.linearSum(4.0 / 3.0, y, -1.0 / 3.0, yprev, psi);
         else
            d_ops.scale(1.0, y, psi);
         d_ops.errorWeights(y, d_opt.rtol, d_opt.atol, ewt);
         // predictor: extrapolation through the last two solutions (y_n + h f(y_n) at the first step)
         if (n == 0) {
            if (d_ops.rhs(t, y, fy, 0) != 0) {
               rc = IMPLICIT_ERHS;
               break;
            }
            d_stats.rhs_evals++;
            d_ops.linearSum(1.0, y, h, fy, ycur);
         } else {
            d_ops.linearSum(2.0, y, -1.0, yprev, ycur);
         }
         d_ops.scale(0.0, acor, acor);  // accumulated correction = CVODE's local error estimate
         rc = newton(t + h, gamma, psi, ewt, ycur, fy, res, delta, acor, ytmp, wk, V);
         if (rc != IMPLICIT_OK) break;
         // projection onto the constraint |q| = 1 (CVodeSetProjFn): y <- y + corr
         d_ops.applyProjection(t + h, ycur, delta, acor);
         d_ops.linearSum(1.0, ycur, 1.0, delta, ycur);
         d_stats.projections++;
         d_ops.scale(1.0, y, yprev);
         d_ops.scale(1.0, ycur, y);
         d_ops.postStep(y);
         t += h;
         d_stats.steps++;
      }
      for (auto& v : V) d_ops.release(v);
      Vec* all[] = {&yprev, &psi, &ycur, &fy, &ewt, &res, &delta, &acor, &ytmp, &wk};
      for (Vec* v : all) d_ops.release(*v);
      d_ops.release(d_pv);
      return rc;
   }

   // Variable-step BDF1/BDF2 from t0 to tend (y updated in place), first step h0: the stand-in for
   // CVode(tend, CV_NORMAL) with max order 2.  Variable-coefficient BDF2 on the last two solutions
   // (step ratio w = h_n / h_{n-1}):  y_{n+1} = [(1+w)^2 y_n - w^2 y_{n-1}] / (1+2w) + [(1+w)/(1+2w)] h_n f(y_{n+1}).
   // Local error: the accumulated Newton correction acor = y_{n+1} - y_predicted (CVODE's estimate, after the
   // projection hook has removed its component along q) times the ratio of the leading error terms of corrector
   // and predictor,
   //   BDF1, Euler predictor y_n + h f_n                      1/2                       (first step)
   //   BDF1, linear extrapolation through y_n, y_{n-1}         h / (2h + h1)
   //   BDF2, quadratic extrapolation through y_n .. y_{n-2}    a / (a + h + h1 + h2),  a = h (h + h1) / (2h + h1)
   // (h = h_n, h1 = h_{n-1}, h2 = h_{n-2}; equal steps: 1/3 and 2/11).  Step accepted when the WRMS norm dsm of
   // the estimate is <= 1; next step h * eta, eta = 1 / ((BIAS2 dsm)^(1/(q+1)) + ADDON) with CVODE's constants
   // (BIAS2 = 6, ADDON = 1e-6, growth <= ETAMX = 10, kept when eta < THRESH = 1.5, shrink >= ETAMIN = 0.1 after a
   // failed error test, 0.2 after three of them, ETACF = 0.25 after a failed Newton iteration).  BDF2 from the
   // third step on (the quadratic predictor needs three solutions); no order selection.
   int advanceTo(Vec& y, double t0, double tend, double h0)
   {
      if (!(h0 > 0.0) || !(tend > t0) || (d_opt.order != 1 && d_opt.order != 2) || d_opt.max_krylov_dimension < 1 ||
          d_opt.max_newton_iterations < 1)
         return IMPLICIT_EINVAL;
      const double BIAS2 = 6.0, ADDON = 1.0e-6, ETAMX = 10.0, THRESH = 1.5, ETAMIN = 0.1, ETACF = 0.25;
      const int m = d_opt.max_krylov_dimension;
      Vec y1 = d_ops.clone(y), y2 = d_ops.clone(y), psi = d_ops.clone(y), ycur = d_ops.clone(y), fy = d_ops.clone(y),
          ewt = d_ops.clone(y), res = d_ops.clone(y), delta = d_ops.clone(y), acor = d_ops.clone(y),
          ytmp = d_ops.clone(y), wk = d_ops.clone(y);
      d_pv = d_ops.clone(y);
      std::vector<Vec> V;
      for (int j = 0; j <= m; j++) V.push_back(d_ops.clone(y));
      int rc = IMPLICIT_OK;
      double t = t0, h = d_opt.stop_at_tend ? std::fmin(h0, tend - t0) : h0, h1 = 0.0, h2 = 0.0;  // h1, h2: the last two accepted steps
      if (d_opt.h_max > 0.0) h = std::fmin(h, d_opt.h_max);
      long nacc = 0;  // accepted steps = solutions in the history beyond y
      int nef = 0, ncf = 0;
      bool attempt_failed = false;  // an attempt at the current step has failed (CVODE's etamax = 1)
      d_stats.smallest_step = 0.0, d_stats.largest_step = 0.0;
      while (t < tend && rc == IMPLICIT_OK) {
         if (d_stats.steps >= d_opt.max_steps) {
            rc = IMPLICIT_ETOOMUCHWORK;
            break;
         }
         const bool bdf2 = d_opt.order == 2 && nacc >= 2;
         double gamma, cerr;
         d_ops.errorWeights(y, d_opt.rtol, d_opt.atol, ewt);
         if (bdf2) {
            const double w = h / h1;
            gamma = h * (1.0 + w) / (1.0 + 2.0 * w);
            d_ops.linearSum((1.0 + w) * (1.0 + w) / (1.0 + 2.0 * w), y, -w * w / (1.0 + 2.0 * w), y1, psi);
            // quadratic extrapolation through (t, y), (t - h1, y1), (t - h1 - h2, y2) to t + h
            const double l0 = (h + h1) * (h + h1 + h2) / (h1 * (h1 + h2));
            const double l1 = -h * (h + h1 + h2) / (h1 * h2);
            const double l2 = h * (h + h1) / ((h1 + h2) * h2);
            d_ops.linearSum(l0, y, l1, y1, ycur);
            d_ops.linearSum(1.0, ycur, l2, y2, ycur);
            const double a = h * (h + h1) / (2.0 * h + h1);
            cerr = a / (a + h + h1 + h2);
         } else {
            gamma = h;
            d_ops.scale(1.0, y, psi);
            if (nacc == 0) {
               if (d_ops.rhs(t, y, fy, 0) != 0) {
                  rc = IMPLICIT_ERHS;
                  break;
               }
               d_stats.rhs_evals++;
               d_ops.linearSum(1.0, y, h, fy, ycur);
               cerr = 0.5;
            } else {
               d_ops.linearSum(1.0 + h / h1, y, -h / h1, y1, ycur);
               cerr = h / (2.0 * h + h1);
            }
         }
         d_ops.scale(0.0, acor, acor);
         const int nrc = newton(t + h, gamma, psi, ewt, ycur, fy, res, delta, acor, ytmp, wk, V,
                                d_opt.scale_newton_tolerance ? d_opt.newton_tolerance / cerr : d_opt.newton_tolerance);
         if (nrc == IMPLICIT_ERHS) {
            rc = nrc;
            break;
         }
         double eta;
         if (nrc != IMPLICIT_OK) {
            // CVODE cvNlsFailure handling: retry with h * ETACF
            d_stats.convergence_failures++;
            if (++ncf >= d_opt.max_convergence_failures || (d_opt.h_min > 0.0 && h <= d_opt.h_min * 1.00001)) {
               rc = IMPLICIT_ECONV;
               break;
            }
            eta = ETACF;
            attempt_failed = true;
         } else {
            d_ops.applyProjection(t + h, ycur, delta, acor);
            d_stats.projections++;
            const double dsm = cerr * wrms(acor, ewt);
            d_stats.last_error_estimate = dsm;
            const int q = bdf2 ? 2 : 1;
            eta = 1.0 / (std::pow(BIAS2 * dsm, 1.0 / (q + 1)) + ADDON);
            if (dsm <= 1.0) {
               // accept
               d_ops.linearSum(1.0, ycur, 1.0, delta, ycur);
               d_ops.scale(1.0, y1, y2);
               d_ops.scale(1.0, y, y1);
               d_ops.scale(1.0, ycur, y);
               d_ops.postStep(y);
               t += h;
               h2 = h1, h1 = h;
               nacc++, nef = 0, ncf = 0;
               d_stats.steps++;
               d_stats.last_step = h;
               d_stats.largest_step = std::fmax(d_stats.largest_step, h);
               d_stats.smallest_step = d_stats.smallest_step > 0.0 ? std::fmin(d_stats.smallest_step, h) : h;
               eta = std::fmin(eta, ETAMX);
               if (eta < THRESH) eta = 1.0;
               if (attempt_failed && d_opt.hold_step_after_failure) eta = 1.0;
               attempt_failed = false;
            } else {
               d_stats.error_test_failures++;
               attempt_failed = true;
               if (++nef >= d_opt.max_error_test_failures || (d_opt.h_min > 0.0 && h <= d_opt.h_min * 1.00001)) {
                  rc = IMPLICIT_EERRTEST;
                  break;
               }
               eta = std::fmax(ETAMIN, std::fmin(eta, 0.9));
               if (nef >= 3) eta = std::fmin(eta, 0.2);
            }
         }
         h *= eta;
         if (d_opt.h_max > 0.0) h = std::fmin(h, d_opt.h_max);
         if (d_opt.h_min > 0.0) h = std::fmax(h, d_opt.h_min);
         if (t < tend && d_opt.stop_at_tend) {
            const double left = tend - t;
            if (h >= left * (1.0 - 1.0e-12))
               h = left;  // land on tend
            else if (h > 0.5 * left && eta >= 1.0)
               h = 0.5 * left;  // do not leave a sliver for the last step
         }
      }
      d_stats.t_reached = t;
      for (auto& v : V) d_ops.release(v);
      Vec* all[] = {&y1, &y2, &psi, &ycur, &fy, &ewt, &res, &delta, &acor, &ytmp, &wk};
      for (Vec* v : all) d_ops.release(*v);
      d_ops.release(d_pv);
      return rc;
   }

 private:
   double wrms(const Vec& x, const Vec& w) { return std::sqrt(d_ops.wdot(x, x, w) / (double)d_ops.length()); }

   // inexact Newton on G(y) = y - psi - gamma f(y); ycur holds the predictor on entry, the solution on exit
   int newton(double t, double gamma, const Vec& psi, const Vec& ewt, Vec& ycur, Vec& fy, Vec& res, Vec& delta,
              Vec& acor, Vec& ytmp, Vec& wk, std::vector<Vec>& V, double newton_tolerance = -1.0)
   {
      if (newton_tolerance <= 0.0) newton_tolerance = d_opt.newton_tolerance;
      double delp = 0.0, crate = 1.0;
      for (int it = 0; it < d_opt.max_newton_iterations; it++) {
         if (d_ops.rhs(t, ycur, fy, 0) != 0) return IMPLICIT_ERHS;
         d_stats.rhs_evals++;
         if (d_ops.preconditioned()) {
            if (d_ops.precondSetup(t, ycur, gamma) != 0) return IMPLICIT_ERHS;
            d_stats.precond_setups++;
         }
         // res = -G = psi + gamma f - y
         d_ops.linearSum(1.0, psi, gamma, fy, res);
         d_ops.linearSum(1.0, res, -1.0, ycur, res);
         const double lin_tol = d_opt.linear_tolerance_factor * newton_tolerance;
         int rc = gmres(t, gamma, ewt, ycur, fy, res, delta, ytmp, wk, V, lin_tol);
         if (rc != IMPLICIT_OK) return rc;
         if (d_opt.strict_linear_convergence && d_linear_state != 0 && (it > 0 || d_linear_state == 2))
            return IMPLICIT_ENEWTON;  // SUNLS_RES_REDUCED after the first iteration / SUNLS_CONV_FAIL
         d_ops.linearSum(1.0, ycur, 1.0, delta, ycur);
         d_ops.linearSum(1.0, acor, 1.0, delta, acor);
         d_stats.newton_iterations++;
         const double del = wrms(delta, ewt);
         d_stats.last_newton_update = del;
         if (it > 0) {
            crate = std::fmax(0.3 * crate, del / delp);  // CVODE CRDOWN
            if (del > 2.0 * delp) return IMPLICIT_ENEWTON;  // CVODE RDIV: diverging
         }
         if (del * std::fmin(1.0, crate) <= newton_tolerance) return IMPLICIT_OK;
         delp = del;
      }
      return IMPLICIT_ENEWTON;
   }

   // (I - gamma J) v with the difference-quotient Jacobian (CVODE cvLsDQJtimes with the JTimes RHS)
   int jtimes(double t, double gamma, const Vec& ewt, const Vec& y, const Vec& fy, const Vec& v, Vec& out, Vec& ytmp)
   {
      const double vn = wrms(v, ewt);
      if (!(vn > 0.0)) {
         d_ops.scale(1.0, v, out);
         return IMPLICIT_OK;
      }
      const double sig = 1.0 / vn;
      d_ops.linearSum(1.0, y, sig, v, ytmp);
      if (d_ops.rhs(t, ytmp, out, 1) != 0) return IMPLICIT_ERHS;
      d_stats.jtimes_evals++;
      d_ops.linearSum(1.0 / sig, out, -1.0 / sig, fy, out);  // J v
      d_ops.linearSum(1.0, v, -gamma, out, out);
      return IMPLICIT_OK;
   }

   // GMRES(m) without restart in the ewt-weighted inner product (= SPGMR with s1 = s2 = ewt), modified
   // Gram-Schmidt, Givens rotations; x0 = 0; stops when the WRMS norm of the linear residual <= tol
   int gmres(double t, double gamma, const Vec& ewt, const Vec& y, const Vec& fy, const Vec& b, Vec& x, Vec& ytmp,
             Vec& wk, std::vector<Vec>& V, double tol)
   {
      const int m = d_opt.max_krylov_dimension;
      const double invN = 1.0 / (double)d_ops.length();
      const bool pre = d_ops.preconditioned();
      const bool left = pre && d_opt.precondition_left;
      d_ops.scale(0.0, x, x);
      if (left) {
         d_ops.precondSolve(b, V[0]);  // P b
         d_stats.precond_solves++;
      }
      const Vec& b0 = left ? V[0] : b;
      const double beta = std::sqrt(d_ops.wdot(b0, b0, ewt) * invN);
      d_stats.last_linear_residual = beta;
      d_linear_state = 0;  // 0 converged, 1 residual reduced but above tol, 2 not reduced
      if (beta <= tol) {
         // the predictor already solves the (preconditioned) system to the tolerance: x = 0
         return IMPLICIT_OK;
      }
      std::vector<std::vector<double> > H(m + 1, std::vector<double>(m, 0.0));
      std::vector<double> cs(m, 0.0), sn(m, 0.0), g(m + 1, 0.0);
      g[0] = beta;
      d_ops.scale(1.0 / beta, b0, V[0]);
      int k = 0;
      for (int j = 0; j < m; j++) {
         int rc;
         if (left) {
            rc = jtimes(t, gamma, ewt, y, fy, V[j], d_pv, ytmp);  // left preconditioning: the Krylov space of P A
            if (rc == IMPLICIT_OK) {
               d_ops.precondSolve(d_pv, wk);
               d_stats.precond_solves++;
            }
         } else if (pre) {
            d_ops.precondSolve(V[j], d_pv);  // right preconditioning: the Krylov space of A P
            d_stats.precond_solves++;
            rc = jtimes(t, gamma, ewt, y, fy, d_pv, wk, ytmp);
         } else {
            rc = jtimes(t, gamma, ewt, y, fy, V[j], wk, ytmp);
         }
         if (rc != IMPLICIT_OK) return rc;
         d_stats.linear_iterations++;
         double col2 = 0.0;  // |A v_j|^2 = sum_i H[i][j]^2 (Pythagoras over the orthonormal basis)
         for (int i = 0; i <= j; i++) {
            H[i][j] = d_ops.wdot(wk, V[i], ewt) * invN;
            d_ops.linearSum(1.0, wk, -H[i][j], V[i], wk);
            col2 += H[i][j] * H[i][j];
         }
         H[j + 1][j] = std::sqrt(d_ops.wdot(wk, wk, ewt) * invN);
         // breakdown: A v_j lies in the current Krylov space up to rounding -- the remainder is
         // difference-quotient noise and must not become a basis vector
         const bool breakdown = !(H[j + 1][j] > 1.0e-12 * std::sqrt(col2 + H[j + 1][j] * H[j + 1][j]));
         if (breakdown) H[j + 1][j] = 0.0;
         for (int i = 0; i < j; i++) {
            const double a = cs[i] * H[i][j] + sn[i] * H[i + 1][j];
            H[i + 1][j] = -sn[i] * H[i][j] + cs[i] * H[i + 1][j];
            H[i][j] = a;
         }
         const double r = std::hypot(H[j][j], H[j + 1][j]);
         cs[j] = (r > 0.0) ? H[j][j] / r : 1.0;
         sn[j] = (r > 0.0) ? H[j + 1][j] / r : 0.0;
         const double hsub = H[j + 1][j];
         H[j][j] = r;
         H[j + 1][j] = 0.0;
         g[j + 1] = -sn[j] * g[j];
         g[j] = cs[j] * g[j];
         k = j + 1;
         d_stats.last_linear_residual = std::fabs(g[j + 1]);
         if (std::fabs(g[j + 1]) <= tol || !(hsub > 0.0)) break;
         d_ops.scale(1.0 / hsub, wk, V[j + 1]);
      }
      if (!(d_stats.last_linear_residual <= tol)) d_linear_state = (d_stats.last_linear_residual < beta) ? 1 : 2;
      // back substitution, x = sum c_i V_i
      std::vector<double> c(k, 0.0);
      for (int i = k - 1; i >= 0; i--) {
         double s = g[i];
         for (int l = i + 1; l < k; l++) s -= H[i][l] * c[l];
         c[i] = s / H[i][i];
      }
      for (int i = 0; i < k; i++) d_ops.linearSum(1.0, x, c[i], V[i], x);
      if (pre && !left && k > 0) {
         d_ops.scale(1.0, x, wk);
         d_ops.precondSolve(wk, x);  // x = P (sum c_i v_i)
         d_stats.precond_solves++;
      }
      return IMPLICIT_OK;  // like CVODE, a reduced residual is accepted; Newton decides
   }

   Ops& d_ops;
   int d_linear_state = 0;
   Vec d_pv;
   ImplicitOptions d_opt;
   ImplicitStats d_stats;
};

}  // namespace ampe_host
