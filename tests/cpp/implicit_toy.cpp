// Host-logic check of ampe_b200/host/ImplicitIntegrator.h with a toy vector backend (no GPU, no
// AMPE physics): the BDF coefficients, the Newton iteration and the matrix-free GMRES are pinned by
// problems whose discrete solutions are known in closed form.  Built and run by
// tests/test_implicit_integrator.py; prints "name value" lines.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../ampe_b200/host/ImplicitIntegrator.h"

using ampe_host::ImplicitIntegrator;
using ampe_host::ImplicitOptions;

struct ToyOps {
   typedef std::vector<double> Vec;
   int kind;  // 0: periodic 1D diffusion y' = D lap_h y   1: y' = -y^3 per component
   double D, h;
   long nrhs0 = 0, nrhs1 = 0;
   Vec clone(const Vec& y) { return y; }
   void release(Vec& v) { v.clear(); }
   void linearSum(double a, const Vec& x, double b, const Vec& y, Vec& z)
   {
      for (size_t i = 0; i < x.size(); i++) z[i] = a * x[i] + b * y[i];
   }
   void scale(double a, const Vec& x, Vec& z)
   {
      for (size_t i = 0; i < x.size(); i++) z[i] = a * x[i];
   }
   double wdot(const Vec& x, const Vec& y, const Vec& w)
   {
      double s = 0.0;
      for (size_t i = 0; i < x.size(); i++) s += (x[i] * w[i]) * (y[i] * w[i]);
      return s;
   }
   long long n = 0;
   long long length() const { return n; }
   void errorWeights(const Vec& y, double rtol, double atol, Vec& w)
   {
      for (size_t i = 0; i < y.size(); i++) w[i] = 1.0 / (rtol * std::fabs(y[i]) + atol);
   }
   int rhs(double, const Vec& y, Vec& f, int fd_flag)
   {
      (fd_flag ? nrhs1 : nrhs0)++;
      const size_t N = y.size();
      if (kind == 0) {
         for (size_t i = 0; i < N; i++)
            f[i] = D * (y[(i + 1) % N] - 2.0 * y[i] + y[(i + N - 1) % N]) / (h * h);
      } else {
         for (size_t i = 0; i < N; i++) f[i] = -y[i] * y[i] * y[i];
      }
      return 0;
   }
   void applyProjection(double, const Vec&, Vec& corr, Vec&)
   {
      for (auto& c : corr) c = 0.0;
   }
   void postStep(Vec&) {}
   bool preconditioned() const { return false; }
   int precondSetup(double, const Vec&, double) { return 0; }
   void precondSolve(const Vec& r, Vec& z) { z = r; }
};

int main()
{
   const double PI = std::acos(-1.0);
   const int N = 48;
   // ---- linear diffusion, three Fourier modes: every mode is an eigenvector of the discrete
   // Laplacian, so BDF1 multiplies its amplitude by 1/(1 - lambda h) per step and BDF2 follows the
   // two-step recurrence with a BDF1 start
   for (int order = 1; order <= 2; order++) {
      ToyOps ops;
      ops.kind = 0, ops.D = 0.7, ops.h = 1.0 / N, ops.n = N;
      ImplicitOptions o;
      o.order = order;
      o.rtol = 1e-10, o.atol = 1e-12;
      o.max_krylov_dimension = 12, o.max_newton_iterations = 6;
      const int modes[3] = {1, 3, 7};
      const double amp[3] = {1.0, 0.5, 0.25};
      std::vector<double> y(N);
      for (int i = 0; i < N; i++) {
         y[i] = 2.0;  // constant mode: lambda = 0
         for (int m = 0; m < 3; m++) y[i] += amp[m] * std::cos(2.0 * PI * modes[m] * i / N);
      }
      const double dt = 2.0e-3;  // explicit limit h^2/(2D) = 3.1e-4: a stiff step
      const int nsteps = 25;
      ImplicitIntegrator<ToyOps> integ(ops, o);
      const int rc = integ.advance(y, 0.0, dt, nsteps);
      double a[3];
      for (int m = 0; m < 3; m++) {
         const double lam = -4.0 * ops.D / (ops.h * ops.h) * std::pow(std::sin(PI * modes[m] / N), 2);
         double prev = amp[m], cur = amp[m] / (1.0 - lam * dt);  // BDF1 first step
         for (int n = 1; n < nsteps; n++) {
            double next;
            if (order == 1)
               next = cur / (1.0 - lam * dt);
            else
               next = ((4.0 / 3.0) * cur - (1.0 / 3.0) * prev) / (1.0 - (2.0 / 3.0) * lam * dt);
            prev = cur, cur = next;
         }
         a[m] = cur;
      }
      double err = 0.0;
      for (int i = 0; i < N; i++) {
         double want = 2.0;
         for (int m = 0; m < 3; m++) want += a[m] * std::cos(2.0 * PI * modes[m] * i / N);
         err = std::fmax(err, std::fabs(y[i] - want));
      }
      std::printf("diffusion_bdf%d_rc %d\n", order, rc);
      std::printf("diffusion_bdf%d_err %.3e\n", order, err);
      std::printf("diffusion_bdf%d_linear_iterations %ld\n", order, integ.stats().linear_iterations);
      std::printf("diffusion_bdf%d_jtimes_fd1 %ld\n", order, ops.nrhs1);
      std::printf("diffusion_bdf%d_rhs_fd0 %ld\n", order, ops.nrhs0);
   }
   // ---- nonlinear decay y' = -y^3: every BDF1 step satisfies y1 - y0 + h y1^3 = 0
   {
      ToyOps ops;
      ops.kind = 1, ops.n = 5;
      ImplicitOptions o;
      o.order = 1;
      o.rtol = 1e-10, o.atol = 1e-12;
      o.max_krylov_dimension = 5, o.max_newton_iterations = 8;
      std::vector<double> y = {0.5, 1.0, 1.5, 2.0, 3.0}, y0 = y;
      const double dt = 0.05;
      ImplicitIntegrator<ToyOps> integ(ops, o);
      const int rc = integ.advance(y, 0.0, dt, 1);
      double res = 0.0;
      for (size_t i = 0; i < y.size(); i++) res = std::fmax(res, std::fabs(y[i] - y0[i] + dt * y[i] * y[i] * y[i]));
      std::printf("cubic_rc %d\n", rc);
      std::printf("cubic_residual %.3e\n", res);
      std::printf("cubic_newton_iterations %ld\n", integ.stats().newton_iterations);
   }
   // ---- observed order of accuracy on y' = -y^3 against the exact y(t) = y0 / sqrt(1 + 2 y0^2 t)
   for (int order = 1; order <= 2; order++) {
      double errs[2];
      for (int r = 0; r < 2; r++) {
         ToyOps ops;
         ops.kind = 1, ops.n = 1;
         ImplicitOptions o;
         o.order = order;
         o.rtol = 1e-10, o.atol = 1e-12;
         o.max_newton_iterations = 8;
         std::vector<double> y = {1.0};
         const int nsteps = 40 << r;
         ImplicitIntegrator<ToyOps> integ(ops, o);
         integ.advance(y, 0.0, 1.0 / nsteps, nsteps);
         errs[r] = std::fabs(y[0] - 1.0 / std::sqrt(3.0));
      }
      std::printf("cubic_bdf%d_observed_order %.3f\n", order, std::log2(errs[0] / errs[1]));
   }
   // ---- a step the Newton iteration cannot complete is reported, not hidden
   {
      ToyOps ops;
      ops.kind = 1, ops.n = 1;
      ImplicitOptions o;
      o.order = 1;
      o.rtol = 1e-10, o.atol = 1e-12;
      o.max_newton_iterations = 1;
      std::vector<double> y = {3.0};
      ImplicitIntegrator<ToyOps> integ(ops, o);
      std::printf("starved_newton_rc %d\n", integ.advance(y, 0.0, 1.0, 1));
      std::printf("bad_step_rc %d\n", integ.advance(y, 0.0, -1.0, 1));
   }
   // ---- advanceTo: variable steps with the local error test.  Semi-discrete diffusion: the exact solution of
   // the ODE system is amp_m exp(lambda_m t) per Fourier mode (rates 27 .. 1290), so the GLOBAL error of the
   // adaptive run can be compared with the tolerance it was asked for.
   for (int tolcase = 0; tolcase < 3; tolcase++) {
      ToyOps ops;
      ops.kind = 0, ops.D = 0.7, ops.h = 1.0 / N, ops.n = N;
      ImplicitOptions o;
      o.order = 2;
      const double tol = tolcase == 0 ? 1e-3 : tolcase == 1 ? 1e-5 : 1e-7;
      o.rtol = tol, o.atol = tol * 1e-2;
      o.max_krylov_dimension = 12, o.max_newton_iterations = 4;
      o.newton_tolerance = 0.05;
      o.max_steps = 20000;
      const int modes[3] = {1, 3, 7};
      const double amp[3] = {1.0, 0.5, 0.25};
      std::vector<double> y(N);
      for (int i = 0; i < N; i++) {
         y[i] = 2.0;
         for (int m = 0; m < 3; m++) y[i] += amp[m] * std::cos(2.0 * PI * modes[m] * i / N);
      }
      const double tend = 0.05;
      ImplicitIntegrator<ToyOps> integ(ops, o);
      const int rc = integ.advanceTo(y, 0.0, tend, 1.0e-6);
      double err = 0.0;
      for (int i = 0; i < N; i++) {
         double want = 2.0;
         for (int m = 0; m < 3; m++) {
            const double lam = -4.0 * ops.D / (ops.h * ops.h) * std::pow(std::sin(PI * modes[m] / N), 2);
            want += amp[m] * std::exp(lam * tend) * std::cos(2.0 * PI * modes[m] * i / N);
         }
         err = std::fmax(err, std::fabs(y[i] - want));
      }
      const ampe_host::ImplicitStats& st = integ.stats();
      std::printf("adaptive%d_rc %d\n", tolcase, rc);
      std::printf("adaptive%d_err_over_tol %.4f\n", tolcase, err / tol);
      std::printf("adaptive%d_steps %ld\n", tolcase, st.steps);
      std::printf("adaptive%d_error_test_failures %ld\n", tolcase, st.error_test_failures);
      std::printf("adaptive%d_step_growth %.3e\n", tolcase, st.largest_step / st.smallest_step);
      std::printf("adaptive%d_t_reached_err %.3e\n", tolcase, std::fabs(st.t_reached - tend));
   }
   // nonlinear: y' = -y^3 to t = 1 against the exact solution
   {
      ToyOps ops;
      ops.kind = 1, ops.n = 1;
      ImplicitOptions o;
      o.order = 2;
      o.rtol = 1e-6, o.atol = 1e-9;
      o.max_newton_iterations = 4;
      std::vector<double> y = {1.0};
      ImplicitIntegrator<ToyOps> integ(ops, o);
      const int rc = integ.advanceTo(y, 0.0, 1.0, 1.0e-4);
      std::printf("adaptive_cubic_rc %d\n", rc);
      std::printf("adaptive_cubic_err %.3e\n", std::fabs(y[0] - 1.0 / std::sqrt(3.0)));
      std::printf("adaptive_cubic_steps %ld\n", integ.stats().steps);
   }
   // the failure codes: too much work, a minimum step that cannot meet the tolerance, bad arguments
   {
      ToyOps ops;
      ops.kind = 1, ops.n = 1;
      ImplicitOptions o;
      o.order = 2;
      o.rtol = 1e-8, o.atol = 1e-11;
      o.max_newton_iterations = 4;
      o.max_steps = 5;
      std::vector<double> y = {1.0};
      ImplicitIntegrator<ToyOps> a(ops, o);
      std::printf("adaptive_too_much_work_rc %d\n", a.advanceTo(y, 0.0, 1.0, 1.0e-4));
      o.max_steps = 500, o.h_min = 0.25, o.max_newton_iterations = 12;
      y = {1.0};
      ImplicitIntegrator<ToyOps> b(ops, o);
      std::printf("adaptive_hmin_rc %d\n", b.advanceTo(y, 0.0, 1.0, 0.25));
      ImplicitIntegrator<ToyOps> c(ops, o);
      std::printf("adaptive_bad_interval_rc %d\n", c.advanceTo(y, 1.0, 1.0, 0.1));
   }
   return 0;
}
