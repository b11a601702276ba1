// TEST INFRASTRUCTURE ONLY (see oracle.h).  CPU restatement of AMPE's Fortran
// kernels, dimension-generic (the 2d/ and 3d/ m4 files differ only by the
// extra loop and the extra term, which are checked to keep the same summation
// order; citations give both files).
#include "oracle.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace oracle {

// ---------------------------------------------------------------- views ----
size_t view_size(const Box& b, int axis, int ng, int depth)
{
   size_t s = 1;
   for (int d = 0; d < 3; d++) {
      int n = b.hi[d] - b.lo[d] + 1;
      if (d < b.ndim) n += 2 * ng;
      if (d == axis) n += 1;
      s *= (size_t)n;
   }
   return s * (size_t)depth;
}
View make_view(double* p, const Box& b, int axis, int ng, int depth)
{
   View v;
   v.p = p;
   size_t s = 1;
   for (int d = 0; d < 3; d++) {
      int g = (d < b.ndim) ? ng : 0;
      v.lo[d] = b.lo[d] - g;
      v.n[d] = b.hi[d] - b.lo[d] + 1 + 2 * g + (d == axis ? 1 : 0);
      s *= (size_t)v.n[d];
   }
   v.comp = s;
   (void)depth;
   return v;
}
IView make_iview(int* p, const Box& b, int axis, int ng)
{
   IView v;
   v.p = p;
   for (int d = 0; d < 3; d++) {
      int g = (d < b.ndim) ? ng : 0;
      v.lo[d] = b.lo[d] - g;
      v.n[d] = b.hi[d] - b.lo[d] + 1 + 2 * g + (d == axis ? 1 : 0);
   }
   return v;
}
void Field::alloc(const Box& b, int axis, int ng, int depth)
{
   data.assign(view_size(b, axis, ng, depth), 0.0);
   v = make_view(data.data(), b, axis, ng, depth);
}
void SideField::alloc(const Box& b, int ng, int depth)
{
   for (int d = 0; d < b.ndim; d++) a[d].alloc(b, d, ng, depth);
}

static inline int E(int a, int d) { return a == d ? 1 : 0; }
// loop bounds of the side box of axis a grown by g in the transverse dirs
// The perf build (liboracle_perf.so, -DORACLE_OMP) threads the two outer loops of every
// sweep; the parity build is serial.  Loop bodies only write their own (i,j,k).
#ifdef ORACLE_OMP
#define ORACLE_OMP_FOR _Pragma("omp parallel for collapse(2) schedule(static)")
#else
#define ORACLE_OMP_FOR
#endif
#define FOR_BOX_SERIAL(i, j, k, L0, H0, L1, H1, L2, H2) \
   for (int k = (L2); k <= (H2); k++)                   \
      for (int j = (L1); j <= (H1); j++)                \
         for (int i = (L0); i <= (H0); i++)
#define FOR_BOX(i, j, k, L0, H0, L1, H1, L2, H2) \
   ORACLE_OMP_FOR                                \
   for (int k = (L2); k <= (H2); k++)            \
      for (int j = (L1); j <= (H1); j++)         \
         for (int i = (L0); i <= (H0); i++)
#define FOR_CELLS(b, i, j, k) \
   FOR_BOX(i, j, k, b.lo[0], b.hi[0], b.lo[1], b.hi[1], b.lo[2], b.hi[2])
// faces of axis a of the interior box (lower faces, one extra in direction a)
#define FOR_SIDES(b, a, i, j, k)                                                   \
   FOR_BOX(i, j, k, b.lo[0], b.hi[0] + E(a, 0), b.lo[1], b.hi[1] + E(a, 1), b.lo[2], \
           b.hi[2] + E(a, 2))

// REAL*4 literals of the Fortran sources, promoted to double like gfortran does
static const double R4_1EM12 = (double)1.e-12f;

// ------------------------------------------------------------ quatrhs.m4 ---
// gradient_flux: 2d/quatrhs.m4:12-53, 3d/quatrhs.m4:12-65
void gradient_flux(const Box& b, const double* h, double epsilon, View phase, View* flux)
{
   const double epsilon2 = epsilon * epsilon;
   for (int a = 0; a < b.ndim; a++) {
      const double dinv = epsilon2 / h[a];
      FOR_SIDES(b, a, i, j, k)
      {
         flux[a](i, j, k) =
             (phase(i, j, k) - phase(i - E(a, 0), j - E(a, 1), k - E(a, 2))) * dinv;
      }
   }
}

// compute_flux_isotropic: 2d/quatrhs.m4:106-151 (3d stops: 3d/quatrhs.m4:90-92)
void compute_flux_isotropic(const Box& b, const double* h, double epsilon, View phase,
                            View* flux)
{
   if (b.ndim != 2) {
      fprintf(stderr, "compute_flux_isotropic: incomplete in 3D (reference stops)\n");
      abort();
   }
   const double epsilon2 = epsilon * epsilon;
   const double dxinv = (1.0 / 12.0) * epsilon2 / h[0];
   const double dyinv = (1.0 / 12.0) * epsilon2 / h[1];
   FOR_SIDES(b, 0, i, j, k)
   {
      flux[0](i, j) = dxinv * ((phase(i, j - 1) - phase(i - 1, j - 1)) +
                               (phase(i, j) - phase(i - 1, j)) * 10.0 +
                               (phase(i, j + 1) - phase(i - 1, j + 1)));
   }
   FOR_SIDES(b, 1, i, j, k)
   {
      flux[1](i, j) = dyinv * ((phase(i - 1, j) - phase(i - 1, j - 1)) +
                               (phase(i, j) - phase(i, j - 1)) * 10.0 +
                               (phase(i + 1, j) - phase(i + 1, j - 1)));
   }
}

// compute_dgamma: 3d/quatrhs.m4:119-147
static void compute_dgamma(const double* quat, const double* n, double* dgamma, double* n4)
{
   double qp[4], qtmp[4], np[4];
   quatconj(quat, qp);
   quatmult4(n, qp, qtmp);
   quatmult4(quat, qtmp, np);
   const double a2 = np[1] * np[1], a3 = np[2] * np[2], a4 = np[3] * np[3];
   *n4 = a2 * a2 + a3 * a3 + a4 * a4;  // np**4: gfortran expands x**4 as (x*x)*(x*x)
   double dg[4];
   dg[0] = 0.;
   dg[1] = np[1] * (np[1] * np[1] - *n4);
   dg[2] = np[2] * (np[2] * np[2] - *n4);
   dg[3] = np[3] * (np[3] * np[3] - *n4);
   quatmult4(dg, quat, qtmp);
   quatmult4(qp, qtmp, dgamma);
}

// anisotropic_gradient_flux: 2d/quatrhs.m4:154-256, 3d/quatrhs.m4:149-349
void anisotropic_gradient_flux(const Box& b, const double* h, double epsilon, double nu,
                               int knumber, View phase, View quat, int qlen, View* flux)
{
   if (b.ndim == 2) {
      const double pi = 4.0 * atan(1.0);
      const double dxinv = 1.0 / h[0];
      const double dyinv = 1.0 / h[1];
      // x faces
      FOR_SIDES(b, 0, i, j, k)
      {
         double dphidx = (phase(i, j) - phase(i - 1, j)) * dxinv;
         double dphidy = 0.25 *
                         (phase(i - 1, j + 1) - phase(i - 1, j - 1) + phase(i, j + 1) -
                          phase(i, j - 1)) *
                         dyinv;
         double theta;
         if (fabs(dphidx) > R4_1EM12)
            theta = atan(dphidy / dphidx);
         else
            theta = 0.5 * pi;
         double q = 0.5 * (quat(i - 1, j, 0, 0) + quat(i, j, 0, 0));
         if (q > 1.0) q = 1.0;
         if (q < -1.0) q = -1.0;
         double phi;
         if (qlen == 4)
            phi = 2.0 * acos(q);
         else
            phi = acos(q);
         double epstheta = epsilon * (1.0 + nu * cos(knumber * (theta - phi)));
         double depsdtheta = -knumber * epsilon * nu * sin(knumber * (theta - phi));
         flux[0](i, j) = epstheta * epstheta * dphidx - epstheta * depsdtheta * dphidy;
      }
      // y faces
      FOR_SIDES(b, 1, i, j, k)
      {
         double dphidx = 0.25 *
                         (phase(i + 1, j - 1) - phase(i - 1, j - 1) + phase(i + 1, j) -
                          phase(i - 1, j)) *
                         dxinv;
         double dphidy = (phase(i, j) - phase(i, j - 1)) * dyinv;
         double theta;
         if (fabs(dphidx) > R4_1EM12)
            theta = atan(dphidy / dphidx);
         else
            theta = 0.5 * pi;
         double q = 0.5 * (quat(i, j - 1, 0, 0) + quat(i, j, 0, 0));
         if (q > 1.0) q = 1.0;
         if (q < -1.0) q = -1.0;
         double phi;
         if (qlen == 4)
            phi = 2. * acos(q);
         else
            phi = acos(q);
         double epstheta = epsilon * (1. + nu * cos(knumber * (theta - phi)));
         double depsdtheta = -knumber * epsilon * nu * sin(knumber * (theta - phi));
         flux[1](i, j) = epstheta * epstheta * dphidy + epstheta * depsdtheta * dphidx;
      }
      return;
   }
   // 3D: nu plays the role of eps4
   const double eps4 = nu;
   const double dinv[3] = {1. / h[0], 1. / h[1], 1. / h[2]};
   const double factor = 4. * eps4 / (1. - 3. * eps4);
   const double threshold = R4_1EM12;
   for (int a = 0; a < 3; a++) {
      FOR_SIDES(b, a, i, j, k)
      {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         double g[3];
         for (int t = 0; t < 3; t++) {
            if (t == a) {
               g[t] = (phase(i, j, k) - phase(im, jm, km)) * dinv[t];
            } else {
               const int t0 = E(t, 0), t1 = E(t, 1), t2 = E(t, 2);
               // x faces: (i-1,+t) - (i-1,-t) + (i,+t) - (i,-t)
               // y/z faces: (+t, lower) - (-t, lower) + (+t, upper) - (-t, upper)
               g[t] = 0.25 *
                      (phase(im + t0, jm + t1, km + t2) - phase(im - t0, jm - t1, km - t2) +
                       phase(i + t0, j + t1, k + t2) - phase(i - t0, j - t1, k - t2)) *
                      dinv[t];
            }
         }
         const double gphi2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
         double dgamma[4], n4;
         if (fabs(gphi2) > threshold) {
            const double nni = 1. / sqrt(gphi2);
            double n[4] = {0., g[0] * nni, g[1] * nni, g[2] * nni};
            double q[4];
            for (int m = 0; m < 4; m++)
               q[m] = 0.5 * (quat(im, jm, km, m) + quat(i, j, k, m));
            compute_dgamma(q, n, dgamma, &n4);
         } else {
            dgamma[0] = 0.;
            dgamma[1] = 0.;
            dgamma[2] = (a == 2) ? 1. : 0.;
            dgamma[3] = (a == 2) ? 0. : 1.;
            n4 = 0.;
         }
         const double gamma = epsilon * (1. - 3. * eps4) * (1. + factor * n4);
         flux[a](i, j, k) =
             gamma * gamma * g[a] + 16. * epsilon * gamma * eps4 * sqrt(gphi2) * dgamma[a + 1];
      }
   }
}

// computerhspbg: 2d/quatrhs.m4:264-405, 3d/quatrhs.m4:357-515 (full argument list; the eta block runs
// only for three_phase != 0)
void computerhspbg(const Box& b, const double* dx, double misorientation_factor,
                   double epsilonq, View* flux, View temp, double phi_well_scale,
                   double eta_well_scale, View phi, View eta, View orient_grad_mod, View rhs,
                   char phi_well_type, char eta_well_type, char energy_interp_type,
                   char orient_interp1, char orient_interp2, int with_orient, int three_phase)
{
   double dinv[3];
   for (int d = 0; d < b.ndim; d++) dinv[d] = 1.0 / dx[d];
   const double epsilonq2 = 0.5 * epsilonq * epsilonq;
   FOR_CELLS(b, i, j, k)
   {
      double diff_term = (flux[0](i + 1, j, k) - flux[0](i, j, k)) * dinv[0];
      diff_term = diff_term + (flux[1](i, j + 1, k) - flux[1](i, j, k)) * dinv[1];
      if (b.ndim == 3)
         diff_term = diff_term + (flux[2](i, j, k + 1) - flux[2](i, j, k)) * dinv[2];
      rhs(i, j, k) = diff_term;
      const double g_prime = deriv_well_func(phi(i, j, k), phi_well_type);
      rhs(i, j, k) = rhs(i, j, k) - phi_well_scale * g_prime;
   }
   if (three_phase != 0) {
      // eta energy well
      FOR_CELLS(b, i, j, k)
      {
         const double h_prime = deriv_interp_func(phi(i, j, k), energy_interp_type);
         const double g = well_func(eta(i, j, k), eta_well_type);
         rhs(i, j, k) = rhs(i, j, k) - eta_well_scale * h_prime * g;
      }
   }
   if (with_orient != 0) {
      FOR_CELLS(b, i, j, k)
      {
         const double p1_prime = deriv_interp_func(phi(i, j, k), orient_interp1);
         const double p2_prime = deriv_interp_func(phi(i, j, k), orient_interp2);
         rhs(i, j, k) = rhs(i, j, k) -
                        misorientation_factor * temp(i, j, k) * p1_prime *
                            orient_grad_mod(i, j, k) -
                        p2_prime * epsilonq2 * orient_grad_mod(i, j, k) *
                            orient_grad_mod(i, j, k);
      }
   }
}

// computerhsdeltatemperature: 2d/quatrhs.m4:893-940, 3d/quatrhs.m4:1040-1087
void computerhsdeltatemperature(const Box& b, View phi, View temp, double tm, double latentheat, View rhs,
                                char energy_interp_type)
{
   const double alpha = latentheat / tm;
   const double woff = (double)(0.25f / 6.f);  // "woff = 0.25/6." is a REAL*4 expression (3d/quatrhs.m4:1062)
   FOR_CELLS(b, i, j, k)
   {
      double wtemp;
      if (b.ndim == 2)
         wtemp = 0.75 * temp(i, j, k) +
                 0.0625 * (temp(i - 1, j, k) + temp(i, j - 1, k) + temp(i + 1, j, k) + temp(i, j + 1, k));
      else
         wtemp = 0.75 * temp(i, j, k) + woff * (temp(i - 1, j, k) + temp(i, j - 1, k) + temp(i + 1, j, k) +
                                                temp(i, j + 1, k) + temp(i, j, k - 1) + temp(i, j, k + 1));
      const double m = alpha * (tm - wtemp);
      const double h_prime = deriv_interp_func(phi(i, j, k), energy_interp_type);
      rhs(i, j, k) = rhs(i, j, k) + m * h_prime;
   }
}

// phaserhs_fenergy: 2d/quatrhs.m4:587-631
void phaserhs_fenergy(const Box& b, View fl, View fa, View phi, View rhs, char interp)
{
   FOR_CELLS(b, i, j, k)
   {
      const double hphi_prime = deriv_interp_func(phi(i, j, k), interp);
      rhs(i, j, k) = rhs(i, j, k) + hphi_prime * (fl(i, j, k) - fa(i, j, k));
   }
}

// laplacian: 2d/laplacian.m4:12-58, 3d/laplacian.m4
void laplacian(const Box& b, const double* dx, double coeff, View field, View rhs)
{
   double dinv2[3] = {0, 0, 0};
   for (int d = 0; d < b.ndim; d++) dinv2[d] = 1.0 / (dx[d] * dx[d]);
   FOR_CELLS(b, i, j, k)
   {
      const double dtx = (field(i - 1, j, k) - 2.0 * field(i, j, k) + field(i + 1, j, k));
      const double dty = (field(i, j - 1, k) - 2.0 * field(i, j, k) + field(i, j + 1, k));
      double diff_term = dtx * dinv2[0] + dty * dinv2[1];
      if (b.ndim == 3) {
         const double dtz =
             (field(i, j, k - 1) - 2.0 * field(i, j, k) + field(i, j, k + 1));
         diff_term = diff_term + dtz * dinv2[2];
      }
      rhs(i, j, k) = coeff * diff_term;
   }
}

// computerhstemp: 2d/quatrhs.m4:749-806
void computerhstemp(const Box& b, const double* dx, double thermal_diffusivity,
                    double latent_heat, View temp, View cp, int with_phase, View phi_rhs,
                    View rhs)
{
   laplacian(b, dx, thermal_diffusivity, temp, rhs);
   if (with_phase != 0) {
      FOR_CELLS(b, i, j, k)
      {
         const double gamma = latent_heat / cp(i, j, k);
         rhs(i, j, k) = rhs(i, j, k) + gamma * phi_rhs(i, j, k);
      }
   }
}

// computerhsbiaswell: 2d/quatrhs.m4:810-849.  pi = 4.*atan(1.) is REAL*4.
void computerhsbiaswell(const Box& b, View phi, View temp, double alpha, double gamma,
                        View te, View rhs)
{
   const double pi = (double)(4.f * atanf(1.f));
   const double coeff = alpha / pi;
   FOR_CELLS(b, i, j, k)
   {
      const double m = coeff * atan(gamma * (te(i, j, k) - temp(i, j, k)));
      rhs(i, j, k) = rhs(i, j, k) + m * phi(i, j, k) * (1.0 - phi(i, j, k));
   }
}

// ---------------------------------------------------------- quatdiffs.m4 ---
// quatdiffs: 2d/quatdiffs.m4:52-84, 3d/quatdiffs.m4:70-124
void quatdiffs(const Box& b, int depth, View q, View* diff)
{
   for (int m = 0; m < depth; m++)
      for (int a = 0; a < b.ndim; a++) {
         // side box of axis a grown by one in the transverse directions
         int L[3], H[3];
         for (int d = 0; d < 3; d++) {
            int g = (d < b.ndim && d != a) ? 1 : 0;
            L[d] = b.lo[d] - g;
            H[d] = b.hi[d] + g + E(a, d);
         }
         FOR_BOX(i, j, k, L[0], H[0], L[1], H[1], L[2], H[2])
         {
            diff[a](i, j, k, m) =
                q(i, j, k, m) - q(i - E(a, 0), j - E(a, 1), k - E(a, 2), m);
         }
      }
}

// quatdiffs_symm: 2d/quatdiffs.m4:90-150, 3d/quatdiffs.m4:131-218
void quatdiffs_symm(const Box& b, int depth, View q, View* diff, IView* iqrot)
{
   for (int a = 0; a < b.ndim; a++) {
      int L[3], H[3];
      for (int d = 0; d < 3; d++) {
         int g = (d < b.ndim && d != a) ? 1 : 0;
         L[d] = b.lo[d] - g;
         H[d] = b.hi[d] + g + E(a, d);
      }
      FOR_BOX(i, j, k, L[0], H[0], L[1], H[1], L[2], H[2])
      {
         double q2[4], q2_prime[4];
         for (int m = 0; m < depth; m++)
            q2[m] = q(i - E(a, 0), j - E(a, 1), k - E(a, 2), m);
         quatsymmrotate(q2, iqrot[a](i, j, k), q2_prime, depth);
         for (int m = 0; m < depth; m++) diff[a](i, j, k, m) = q(i, j, k, m) - q2_prime[m];
      }
   }
}

// ----------------------------------------------------------- quatgrad.m4 ---
// quatgrad_cell: 2d/quatgrad.m4:13-58, 3d/quatgrad.m4:11-73
void quatgrad_cell(const Box& b, int depth, const double* h, View* diff, View* grad)
{
   for (int m = 0; m < depth; m++)
      for (int a = 0; a < b.ndim; a++) {
         const double p5inv = 0.5 / h[a];
         FOR_CELLS(b, i, j, k)
         {
            grad[a](i, j, k, m) =
                (diff[a](i + E(a, 0), j + E(a, 1), k + E(a, 2), m) + diff[a](i, j, k, m)) *
                p5inv;
         }
      }
}

// quatgrad_cell_symm: 2d/quatgrad.m4:61-150, 3d/quatgrad.m4:76-190
void quatgrad_cell_symm(const Box& b, int depth, const double* h, View* diff, View* grad,
                        IView* iqrot)
{
   for (int a = 0; a < b.ndim; a++) {
      const double p5inv = 0.5 / h[a];
      FOR_CELLS(b, i, j, k)
      {
         double dtmp[4], dprime[4];
         const int ip = i + E(a, 0), jp = j + E(a, 1), kp = k + E(a, 2);
         if (depth > 1) {
            for (int m = 0; m < depth; m++) dtmp[m] = diff[a](ip, jp, kp, m);
            int iq = -1 * iqrot[a](ip, jp, kp);
            quatsymmrotate(dtmp, iq, dprime, depth);
         } else {
            dprime[0] = diff[a](ip, jp, kp, 0);
         }
         for (int m = 0; m < depth; m++)
            grad[a](i, j, k, m) = (dprime[m] + diff[a](i, j, k, m)) * p5inv;
      }
   }
}

// quatgrad_side: 2d/quatgrad.m4:153-207, 3d/quatgrad.m4:193-333
// grad[a] depth index = dir*depth + m (computeQDiffs.cc:46-60)
void quatgrad_side(const Box& b, int depth, const double* h, View* diff, View* grad)
{
   for (int m = 0; m < depth; m++)
      for (int a = 0; a < b.ndim; a++) {
         FOR_SIDES(b, a, i, j, k)
         {
            const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
            for (int t = 0; t < b.ndim; t++) {
               if (t == a) {
                  grad[a](i, j, k, t * depth + m) = (1.0 / h[a]) * diff[a](i, j, k, m);
               } else {
                  const double p25 = 0.25 * (1.0 / h[t]);
                  const int t0 = E(t, 0), t1 = E(t, 1), t2 = E(t, 2);
                  grad[a](i, j, k, t * depth + m) =
                      p25 * (diff[t](im + t0, jm + t1, km + t2, m) + diff[t](im, jm, km, m) +
                             diff[t](i + t0, j + t1, k + t2, m) + diff[t](i, j, k, m));
               }
            }
         }
      }
}

// quatgrad_side_symm: 2d/quatgrad.m4:281-400, 3d/quatgrad.m4:376-720
void quatgrad_side_symm(const Box& b, int depth, const double* h, View* diff, View* grad,
                        IView* iqrot)
{
   for (int a = 0; a < b.ndim; a++) {
      FOR_SIDES(b, a, i, j, k)
      {
         double d1[4], d1p[4], d2[4], d2p[4], d3[4], d4[4], d4p[4];
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         for (int t = 0; t < b.ndim; t++) {
            if (t == a) continue;
            const double p25 = 0.25 * (1.0 / h[t]);
            const int t0 = E(t, 0), t1 = E(t, 1), t2 = E(t, 2);
            if (depth > 1) {
               for (int m = 0; m < depth; m++) {
                  d1[m] = diff[t](i + t0, j + t1, k + t2, m);
                  d2[m] = diff[t](im + t0, jm + t1, km + t2, m);
                  d3[m] = diff[t](im, jm, km, m);
               }
               int iq = -1 * iqrot[t](i + t0, j + t1, k + t2);
               quatsymmrotate(d1, iq, d1p, depth);
               iq = -1 * iqrot[t](im + t0, jm + t1, km + t2);
               quatsymmrotate(d2, iq, d2p, depth);
               for (int m = 0; m < depth; m++) d4[m] = d2p[m] + d3[m];
               iq = iqrot[a](i, j, k);
               quatsymmrotate(d4, iq, d4p, depth);
            } else {
               d1p[0] = diff[t](i + t0, j + t1, k + t2, 0);
               d4p[0] = diff[t](im + t0, jm + t1, km + t2, 0) + diff[t](im, jm, km, 0);
            }
            for (int m = 0; m < depth; m++)
               grad[a](i, j, k, t * depth + m) =
                   p25 * (d4p[m] + d1p[m] + diff[t](i, j, k, m));
         }
         for (int m = 0; m < depth; m++)
            grad[a](i, j, k, a * depth + m) = (1.0 / h[a]) * diff[a](i, j, k, m);
      }
   }
}

// quatgrad_modulus: 2d/quatgrad.m4 (last routine), 3d/quatgrad.m4:793-833
void quatgrad_modulus(const Box& b, int depth, View* grad_cell, View grad_mod)
{
   FOR_CELLS(b, i, j, k)
   {
      double s = 0.0;
      for (int m = 0; m < depth; m++) {
         s = s + grad_cell[0](i, j, k, m) * grad_cell[0](i, j, k, m) +
             grad_cell[1](i, j, k, m) * grad_cell[1](i, j, k, m);
         if (b.ndim == 3) s = s + grad_cell[2](i, j, k, m) * grad_cell[2](i, j, k, m);
      }
      grad_mod(i, j, k) = sqrt(s);
   }
}

// quatgrad_modulus_from_sides_compact: 2d/quatgrad.m4, 3d/quatgrad.m4:725-788
void quatgrad_modulus_from_sides_compact(const Box& b, int depth, View* grad_side,
                                         View grad_mod)
{
   FOR_CELLS(b, i, j, k)
   {
      double s = 0.0;
      for (int a = 0; a < b.ndim; a++) {
         for (int m = 0; m < depth; m++) {
            const double g = grad_side[a](i, j, k, a * depth + m);
            s = s + g * g;
         }
         for (int m = 0; m < depth; m++) {
            const double g =
                grad_side[a](i + E(a, 0), j + E(a, 1), k + E(a, 2), a * depth + m);
            s = s + g * g;
         }
      }
      grad_mod(i, j, k) = sqrt(0.5 * s);
   }
}

// --------------------------------------------------------- quatfacops.m4 ---
// compute_face_coef{2,3}d: 2d/quatfacops.m4:14-123, 3d/quatfacops.m4:14-155
void compute_face_coef(const Box& b, int depth, double eps_q, View phi, View temp,
                       double misorientation_factor, View* gq, View* fc,
                       double gradient_floor, char floor_type, char interp1, char interp2,
                       char avg_type)
{
   const double floor_grad_norm2 = gradient_floor * gradient_floor;
   const double eps2 = eps_q * eps_q;
   const double max_grad_normi = 1.0 / gradient_floor;
   for (int a = 0; a < b.ndim; a++) {
      FOR_SIDES(b, a, i, j, k)
      {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         const double phia = average_func(phi(im, jm, km), phi(i, j, k), avg_type);
         const double tempa = 0.5 * (temp(im, jm, km) + temp(i, j, k));
         const double diff = misorientation_factor * tempa * interp_func(phia, interp1);
         const double hphi2 = interp_func(phia, interp2);
         double grad_norm2 = 0.0;
         for (int n = 0; n < b.ndim; n++)
            for (int m = 0; m < depth; m++) {
               const double g = gq[a](i, j, k, n * depth + m);
               grad_norm2 = grad_norm2 + g * g;
            }
         const double grad_normi =
             eval_grad_normi(grad_norm2, floor_type, floor_grad_norm2, max_grad_normi);
         fc[a](i, j, k) = -grad_normi * diff - eps2 * hphi2;
      }
   }
}

// compute_flux{2,3}d: 2d/quatfacops.m4:173-229, 3d/quatfacops.m4:227-311
void compute_flux(const Box& b, int depth, View* fc, View q, const double* h, View* f)
{
   for (int a = 0; a < b.ndim; a++) {
      const double hinv = 1.0 / h[a];
      for (int m = 0; m < depth; m++) FOR_SIDES(b, a, i, j, k)
         {
            f[a](i, j, k, m) = fc[a](i, j, k) * hinv *
                               (q(i, j, k, m) - q(i - E(a, 0), j - E(a, 1), k - E(a, 2), m));
         }
   }
}

// compute_flux{2,3}d_from_gradq: 2d/quatfacops.m4:232-281, 3d/quatfacops.m4:314-390
void compute_flux_from_gradq(const Box& b, int depth, View* fc, View* grad_side, View* f)
{
   for (int a = 0; a < b.ndim; a++)
      for (int m = 0; m < depth; m++) FOR_SIDES(b, a, i, j, k)
         {
            f[a](i, j, k, m) = fc[a](i, j, k) * grad_side[a](i, j, k, a * depth + m);
         }
}

// compute_lambda_flux{2,3}d: 2d/quatfacops.m4:683-735, 3d/quatfacops.m4:729-791
void compute_lambda_flux(const Box& b, int depth, View* f, View q, const double* h,
                         View lambda)
{
   double fac[3] = {0, 0, 0};
   for (int d = 0; d < b.ndim; d++) fac[d] = 0.5 / h[d];
   FOR_CELLS(b, i, j, k)
   {
      double lam = 0.0, sumq2 = 0.0;
      for (int m = 0; m < depth; m++) {
         double s = (f[0](i + 1, j, k, m) - f[0](i, j, k, m)) * fac[0] +
                    (f[1](i, j + 1, k, m) - f[1](i, j, k, m)) * fac[1];
         if (b.ndim == 3) s = s + (f[2](i, j, k + 1, m) - f[2](i, j, k, m)) * fac[2];
         lam = lam - q(i, j, k, m) * s;
         sumq2 = sumq2 + q(i, j, k, m) * q(i, j, k, m);
      }
      lambda(i, j, k) = lam / sumq2;
   }
}

// add_quat_proj_op{2,3}d: 2d/quatfacops.m4:542-601, 3d/quatfacops.m4:659-727
void add_quat_proj_op(const Box& b, int depth, View mobility, View* f, View q, View lambda,
                      const double* h, View rhs)
{
   double dinv[3] = {0, 0, 0};
   for (int d = 0; d < b.ndim; d++) dinv[d] = 1.0 / h[d];
   for (int m = 0; m < depth; m++) FOR_CELLS(b, i, j, k)
      {
         double divergence = (f[0](i + 1, j, k, m) - f[0](i, j, k, m)) * dinv[0] +
                             (f[1](i, j + 1, k, m) - f[1](i, j, k, m)) * dinv[1];
         if (b.ndim == 3)
            divergence = divergence + (f[2](i, j, k + 1, m) - f[2](i, j, k, m)) * dinv[2];
         rhs(i, j, k, m) =
             rhs(i, j, k, m) -
             mobility(i, j, k) * (divergence + 2.0 * q(i, j, k, m) * lambda(i, j, k));
      }
}

// add_quat_op{2,3}d: 2d/quatfacops.m4:487-539
void add_quat_op(const Box& b, int depth, View mobility, View* f, const double* h, View rhs)
{
   double dinv[3] = {0, 0, 0};
   for (int d = 0; d < b.ndim; d++) dinv[d] = 1.0 / h[d];
   for (int m = 0; m < depth; m++) FOR_CELLS(b, i, j, k)
      {
         double divergence = (f[0](i + 1, j, k, m) - f[0](i, j, k, m)) * dinv[0] +
                             (f[1](i, j + 1, k, m) - f[1](i, j, k, m)) * dinv[1];
         if (b.ndim == 3)
            divergence = divergence + (f[2](i, j, k + 1, m) - f[2](i, j, k, m)) * dinv[2];
         rhs(i, j, k, m) = rhs(i, j, k, m) - mobility(i, j, k) * divergence;
      }
}

// correctrhsquatforsymmetry: 2d/correctrhsquatforsymmetry.m4:15-140, 3d:15-166
void correctrhsquatforsymmetry(const Box& b, int depth, const double* dx, View* nonsymm_diff,
                               View* symm_diff, View rhs, View quat, View* facecoeff,
                               View mobility, IView* iqrot)
{
   double invdx2[3] = {0, 0, 0};
   for (int d = 0; d < b.ndim; d++) invdx2[d] = 1.0 / (dx[d] * dx[d]);
   FOR_CELLS(b, i, j, k)
   {
      double tmp[4], dtmp[4], dprime[3][4];
      for (int a = 0; a < b.ndim; a++) {
         const int ip = i + E(a, 0), jp = j + E(a, 1), kp = k + E(a, 2);
         if (depth > 1) {
            for (int m = 0; m < depth; m++) dtmp[m] = symm_diff[a](ip, jp, kp, m);
            int iq = -1 * iqrot[a](ip, jp, kp);
            quatsymmrotate(dtmp, iq, dprime[a], depth);
         } else {
            dprime[a][0] = symm_diff[a](ip, jp, kp, 0);
         }
      }
      for (int m = 0; m < depth; m++) {
         double t = 0.0;
         for (int a = 0; a < b.ndim; a++) {
            const int ip = i + E(a, 0), jp = j + E(a, 1), kp = k + E(a, 2);
            const double term =
                invdx2[a] *
                (facecoeff[a](ip, jp, kp) * (nonsymm_diff[a](ip, jp, kp, m) - dprime[a][m]) -
                 facecoeff[a](i, j, k) *
                     (nonsymm_diff[a](i, j, k, m) - symm_diff[a](i, j, k, m)));
            t = (a == 0) ? term : t + term;
         }
         tmp[m] = t;
      }
      if (depth > 1) {
         double beta = 0.0, lambda = 0.0;
         for (int m = 0; m < depth; m++) {
            beta = beta + quat(i, j, k, m) * quat(i, j, k, m);
            lambda = lambda + quat(i, j, k, m) * tmp[m];
         }
         lambda = lambda / beta;
         for (int m = 0; m < depth; m++)
            rhs(i, j, k, m) =
                rhs(i, j, k, m) + mobility(i, j, k) * (tmp[m] - lambda * quat(i, j, k, m));
      } else {
         rhs(i, j, k, 0) = rhs(i, j, k, 0) + mobility(i, j, k) * tmp[0];
      }
   }
}

// ----------------------------------------------------------- mobility.m4 ---
// quatmobility: 2d/mobility.m4, 3d/mobility.m4:13-98 (loops over box+ngmobility)
void quatmobility(const Box& b, View phase, View mobility, int ng, double scale_mobility,
                  double min_mobility, char func_type, double alt_scale_factor)
{
   int L[3], H[3];
   for (int d = 0; d < 3; d++) {
      int g = (d < b.ndim) ? ng : 0;
      L[d] = b.lo[d] - g;
      H[d] = b.hi[d] + g;
   }
   FOR_BOX(i, j, k, L[0], H[0], L[1], H[1], L[2], H[2])
   {
      double phi = phase(i, j, k);
      double qfunc;
      if (func_type == 'p' || func_type == 'P') {
         phi = fmax(0.0, fmin(1.0, phi));
         qfunc = phi * phi * phi * (10.0 - 15.0 * phi + 6.0 * phi * phi);
         qfunc = 1.0 - qfunc;
      } else if (func_type == 'e' || func_type == 'E') {
         const double c = alt_scale_factor;
         phi = fmax(0.0, fmin(1.0, phi));
         qfunc = (1.0 - exp(c * phi)) / (1.0 - exp(c));
         qfunc = 1.0 - qfunc;
      } else if (func_type == 'i' || func_type == 'I') {
         phi = fmax(1.e-6, fmin(1.0, phi));
         qfunc = fmax(0.0, (1.0 - phi) / (phi * phi));
         qfunc = fmin(qfunc, alt_scale_factor);
      } else {
         fprintf(stderr, "Error in quatmobility: unknown function type\n");
         abort();
      }
      mobility(i, j, k) = min_mobility + (scale_mobility - min_mobility) * qfunc;
   }
}

// ------------------------------------------------------ concentrationrhs ---
// add_cahnhilliarddoublewell_flux: 2d/concentrationrhs.m4:85-137 (+3d)
// scatter-add over box+1, so flux(i) = (0 - M/h mu(i-1)) + M/h mu(i)
void add_cahnhilliarddoublewell_flux(const Box& b, const double* dx, View conc,
                                     double mobility, double ca, double cb,
                                     double well_scale, double kappa, View* flux)
{
   double dinv[3] = {0, 0, 0}, dinv2[3] = {0, 0, 0};
   for (int d = 0; d < b.ndim; d++) {
      dinv[d] = 1.0 / dx[d];
      dinv2[d] = dinv[d] * dinv[d];
   }
   int L[3], H[3];
   for (int d = 0; d < 3; d++) {
      int g = (d < b.ndim) ? 1 : 0;
      L[d] = b.lo[d] - g;
      H[d] = b.hi[d] + g;
   }
   FOR_BOX_SERIAL(i, j, k, L[0], H[0], L[1], H[1], L[2], H[2])
   {
      double lap = dinv2[0] * (-2.0 * conc(i, j, k) + conc(i - 1, j, k) + conc(i + 1, j, k)) +
                   dinv2[1] * (-2.0 * conc(i, j, k) + conc(i, j - 1, k) + conc(i, j + 1, k));
      if (b.ndim == 3)
         lap = lap +
               dinv2[2] * (-2.0 * conc(i, j, k) + conc(i, j, k - 1) + conc(i, j, k + 1));
      const double c = conc(i, j, k);
      const double mu =
          2.0 * well_scale * (c - ca) * (cb - c) * (cb + ca - 2.0 * c) - kappa * lap;
      for (int a = 0; a < b.ndim; a++) {
         flux[a](i, j, k) = flux[a](i, j, k) + mobility * dinv[a] * mu;
         flux[a](i + E(a, 0), j + E(a, 1), k + E(a, 2)) =
             flux[a](i + E(a, 0), j + E(a, 1), k + E(a, 2)) - mobility * dinv[a] * mu;
      }
   }
}

// computerhsconcentration: 2d/concentrationrhs.m4:326-363, 3d:412-458
void computerhsconcentration(const Box& b, const double* dx, View* flux, double mobility,
                             View rhs)
{
   double dinv[3] = {0, 0, 0};
   for (int d = 0; d < b.ndim; d++) dinv[d] = 1.0 / dx[d];
   FOR_CELLS(b, i, j, k)
   {
      double s = dinv[0] * (flux[0](i + 1, j, k) - flux[0](i, j, k)) +
                 dinv[1] * (flux[1](i, j + 1, k) - flux[1](i, j, k));
      if (b.ndim == 3) s = s + dinv[2] * (flux[2](i, j, k + 1) - flux[2](i, j, k));
      rhs(i, j, k) = mobility * s;
   }
}

// concentrationflux: 2d/concentrationrhs.m4:15-79, 3d/concentrationrhs.m4
void concentrationflux(const Box& b, const double* dx, View conc, View phase, View* diffconc,
                       View* dphi, View* flux)
{
   for (int a = 0; a < b.ndim; a++) {
      const double dinv = 1.0 / dx[a];
      FOR_SIDES(b, a, i, j, k)
      {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         flux[a](i, j, k) =
             dinv * (diffconc[a](i, j, k) * (conc(i, j, k) - conc(im, jm, km)) +
                     dphi[a](i, j, k) * (phase(i, j, k) - phase(im, jm, km)));
      }
   }
}

// add_flux: 2d/flux.m4, 3d/flux.m4:15-96
void add_flux(const Box& b, const double* dx, View conc, int ncomp, View* diffconc,
              View* flux)
{
   for (int ic = 0; ic < ncomp; ic++)
      for (int jc = 0; jc < ncomp; jc++) {
         const int ijc = ic + jc * ncomp;
         for (int a = 0; a < b.ndim; a++) {
            const double dinv = 1.0 / dx[a];
            FOR_SIDES(b, a, i, j, k)
            {
               flux[a](i, j, k, ic) =
                   flux[a](i, j, k, ic) +
                   dinv * (diffconc[a](i, j, k, ijc) *
                           (conc(i, j, k, jc) -
                            conc(i - E(a, 0), j - E(a, 1), k - E(a, 2), jc)));
            }
         }
      }
}

// concentration_pfmdiffusion: 2d/concentrationdiffusion.m4, 3d:14-127
void concentration_pfmdiffusion(const Box& b, View phi, View* diff, View temp, double d_liquid,
                                double q0_liquid, double d_solid_A, double q0_solid_A,
                                double gas_constant_R, char interp_type, char avg_type)
{
   const double q0_liquid_invR = q0_liquid / gas_constant_R;
   const double q0_solidA_invR = q0_solid_A / gas_constant_R;
   for (int a = 0; a < b.ndim; a++) {
      FOR_SIDES(b, a, i, j, k)
      {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         const double vphi = average_func(phi(im, jm, km), phi(i, j, k), avg_type);
         const double hphi = interp_func(vphi, interp_type);
         const double invT = 2.0 / (temp(im, jm, km) + temp(i, j, k));
         const double diff_liquid = d_liquid * exp(-q0_liquid_invR * invT);
         const double diff_solidA = d_solid_A * exp(-q0_solidA_invR * invT);
         diff[a](i, j, k) = (1.0 - hphi) * diff_liquid + hphi * diff_solidA;
      }
   }
}

}  // namespace oracle
