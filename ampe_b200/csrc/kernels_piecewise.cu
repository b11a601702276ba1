// One CUDA kernel per Fortran routine / per-patch C++ loop of AMPE's RHS path, on
// SAMRAI-layout device arrays with ghost widths (include/ampe_b200_kernels.h).  This is
// the piecewise boundary: a Strategy class of the reference can be re-pointed at these
// symbols one call at a time.  The hot path proper is the fused kernel (rhs_tile.cuh, rhs_march.cuh);
// these kernels keep the reference's unfused pass structure (one thread per loop point,
// intermediates in global memory) and the same operation order.
#include <cuda_runtime.h>

#include <string>

#include "../../include/ampe_b200_kernels.h"
#include "calphad.cuh"
#include "params.h"
#include "pointwise.cuh"

int ampe_set_err(int code, const std::string& msg);
int ampe_derive_params(const ampe_rhs_config& c, ampe::Params& p);

namespace {
using namespace ampe;

#include "box_view.cuh"

__device__ __forceinline__ void rot(const double* q, int iq, double* qp, int qlen, const double* qr,
                                    const int* conj)
{
   if (qlen == 4) {
      if (iq < 0) iq = conj[-iq - 1];
      if (iq == 1) {
         for (int m = 0; m < 4; m++) qp[m] = q[m];
      } else {
         quatmult4(q, qr + 4 * (iq - 1), qp);
      }
   } else if (qlen == 2) {
      if (iq < 0) iq = (iq == -2) ? 4 : ((iq == -4) ? 2 : -iq);
      const double r0 = (iq == 1) ? 1.0 : ((iq == 3) ? -1.0 : 0.0);
      const double r1 = (iq == 2) ? 1.0 : ((iq == 4) ? -1.0 : 0.0);
      if (iq == 1) {
         qp[0] = q[0];
         qp[1] = q[1];
      } else {
         qp[0] = q[0] * r0 - q[1] * r1;
         qp[1] = q[0] * r1 + q[1] * r0;
      }
   } else {
      qp[0] = q[0];
   }
}

// 48 cubic rotations (setqr, quat.f:165-286), uploaded once per process
double* g_qr = nullptr;
int* g_conj = nullptr;
static int ensure_qr()
{
   if (g_qr) return AMPE_OK;
   static const int raw[48][4] = {
       {1, 0, 0, 0},    {0, 1, 0, 0},    {0, 0, 1, 0},    {0, 0, 0, 1},    {-1, 0, 0, 0},
       {0, -1, 0, 0},   {0, 0, -1, 0},   {0, 0, 0, -1},   {1, 1, 0, 0},    {1, 0, 1, 0},
       {1, 0, 0, 1},    {0, 1, 1, 0},    {0, 1, 0, 1},    {0, 0, 1, 1},    {-1, 1, 0, 0},
       {-1, 0, 1, 0},   {-1, 0, 0, 1},   {0, -1, 1, 0},   {0, -1, 0, 1},   {0, 0, -1, 1},
       {1, -1, 0, 0},   {1, 0, -1, 0},   {1, 0, 0, -1},   {0, 1, -1, 0},   {0, 1, 0, -1},
       {0, 0, 1, -1},   {-1, -1, 0, 0},  {-1, 0, -1, 0},  {-1, 0, 0, -1},  {0, -1, -1, 0},
       {0, -1, 0, -1},  {0, 0, -1, -1},  {1, 1, 1, 1},    {-1, 1, 1, 1},   {1, -1, 1, 1},
       {1, 1, -1, 1},   {1, 1, 1, -1},   {-1, -1, 1, 1},  {-1, 1, -1, 1},  {-1, 1, 1, -1},
       {1, -1, -1, 1},  {1, -1, 1, -1},  {1, 1, -1, -1},  {1, -1, -1, -1}, {-1, 1, -1, -1},
       {-1, -1, 1, -1}, {-1, -1, -1, 1}, {-1, -1, -1, -1}};
   static const int conj[48] = {1,  6,  7,  8,  5,  2,  3,  4,  21, 22, 23, 30, 31, 32, 27, 28,
                                29, 24, 25, 26, 9,  10, 11, 18, 19, 20, 15, 16, 17, 12, 13, 14,
                                44, 48, 43, 42, 41, 45, 46, 47, 37, 36, 35, 33, 38, 39, 40, 34};
   double qr[48][4];
   for (int n = 0; n < 48; n++) {
      double q[4] = {(double)raw[n][0], (double)raw[n][1], (double)raw[n][2], (double)raw[n][3]};
      const double m = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      const double minv = (m < 1.e-15) ? 0.0 : 1.0 / m;
      for (int k = 0; k < 4; k++) qr[n][k] = q[k] * minv;
   }
   if (cudaMalloc(&g_qr, sizeof(qr)) != cudaSuccess || cudaMalloc(&g_conj, sizeof(conj)) != cudaSuccess)
      return ampe_set_err(AMPE_ECUDA, "cudaMalloc(rotation table)");
   cudaMemcpy(g_qr, qr, sizeof(qr), cudaMemcpyHostToDevice);
   cudaMemcpy(g_conj, conj, sizeof(conj), cudaMemcpyHostToDevice);
   return AMPE_OK;
}
}  // namespace

extern "C" {

// ---- quatrhs.m4 --------------------------------------------------------------------------
int ampe_k_gradient_flux(int ndim, const int* ifirst, const int* ilast, const double* dx,
                         double epsilon, const double* phase, int ngphase, double* const* flux,
                         int ngflux, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(phase, b, -1, ngphase);
   const DV3 fl = sides(flux, b, ngflux);
   const double epsilon2 = epsilon * epsilon;
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const double dinv = epsilon2 / dx[a];
      const DV f = fl.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         f(i, j, k) = (ph(i, j, k) - ph(i - E(a, 0), j - E(a, 1), k - E(a, 2))) * dinv;
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_compute_flux_isotropic(int ndim, const int* ifirst, const int* ilast, const double* dx,
                                  double epsilon, const double* phase, int ngphase,
                                  double* const* flux, int ngflux, void* stream)
{
   if (ndim != 2)
      return ampe_set_err(AMPE_EINVAL, "compute_flux_isotropic: incomplete in 3D (reference stops)");
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(phase, b, -1, ngphase);
   const DV3 fl = sides(flux, b, ngflux);
   const double epsilon2 = epsilon * epsilon;
   for (int a = 0; a < 2; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const double dinv = (1.0 / 12.0) * epsilon2 / dx[a];
      const DV f = fl.a[a];
      const int t = 1 - a;
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int im = i - E(a, 0), jm = j - E(a, 1);
         f(i, j) = dinv * ((ph(i - E(t, 0), j - E(t, 1)) - ph(im - E(t, 0), jm - E(t, 1))) +
                           (ph(i, j) - ph(im, jm)) * 10.0 +
                           (ph(i + E(t, 0), j + E(t, 1)) - ph(im + E(t, 0), jm + E(t, 1))));
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_anisotropic_gradient_flux(int ndim, const int* ifirst, const int* ilast,
                                     const double* dx, double epsilon, double nu, int knumber,
                                     const double* phase, int ngphase, const double* quat, int ngq,
                                     int qlen, double* const* flux, int ngflux, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(phase, b, -1, ngphase);
   const CV q = view(quat, b, -1, ngq);
   const DV3 fl = sides(flux, b, ngflux);
   double dinv[3] = {0, 0, 0};
   for (int d = 0; d < ndim; d++) dinv[d] = 1.0 / dx[d];
   const double di0 = dinv[0], di1 = dinv[1], di2 = dinv[2];
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV f = fl.a[a];
      int rc;
      if (ndim == 2) {
         // 2d/quatrhs.m4:154-256, libm evaluation exactly as written
         rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
            const int t = 1 - a;
            const int im = i - E(a, 0), jm = j - E(a, 1);
            const double dn = (ph(i, j) - ph(im, jm)) * (a == 0 ? di0 : di1);
            const double dt = 0.25 *
                              (ph(im + E(t, 0), jm + E(t, 1)) - ph(im - E(t, 0), jm - E(t, 1)) +
                               ph(i + E(t, 0), j + E(t, 1)) - ph(i - E(t, 0), j - E(t, 1))) *
                              (t == 0 ? di0 : di1);
            const double dphidx = (a == 0) ? dn : dt, dphidy = (a == 0) ? dt : dn;
            double theta;
            if (fabs(dphidx) > (double)1.e-12f)
               theta = atan(dphidy / dphidx);
            else
               theta = 0.5 * 3.141592653589793;
            double qa = 0.5 * (q(im, jm, 0, 0) + q(i, j, 0, 0));
            if (qa > 1.0) qa = 1.0;
            if (qa < -1.0) qa = -1.0;
            const double phi = (qlen == 4) ? 2.0 * acos(qa) : acos(qa);
            const double epstheta = epsilon * (1.0 + nu * cos(knumber * (theta - phi)));
            const double depsdtheta = -knumber * epsilon * nu * sin(knumber * (theta - phi));
            f(i, j) = (a == 0) ? epstheta * epstheta * dphidx - epstheta * depsdtheta * dphidy
                               : epstheta * epstheta * dphidy + epstheta * depsdtheta * dphidx;
         });
      } else {
         // 3d/quatrhs.m4:149-349 (nu = eps4)
         const double eps4 = nu;
         const double factor = 4. * eps4 / (1. - 3. * eps4);
         rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
            const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
            const double di[3] = {di0, di1, di2};
            double g[3];
            for (int t = 0; t < 3; t++) {
               if (t == a) {
                  g[t] = (ph(i, j, k) - ph(im, jm, km)) * di[t];
               } else {
                  const int t0 = E(t, 0), t1 = E(t, 1), t2 = E(t, 2);
                  g[t] = 0.25 *
                         (ph(im + t0, jm + t1, km + t2) - ph(im - t0, jm - t1, km - t2) +
                          ph(i + t0, j + t1, k + t2) - ph(i - t0, j - t1, k - t2)) *
                         di[t];
               }
            }
            const double gphi2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
            double dgamma[4], n4;
            if (fabs(gphi2) > (double)1.e-12f) {
               const double nni = 1. / sqrt(gphi2);
               const double n[4] = {0., g[0] * nni, g[1] * nni, g[2] * nni};
               double qq[4], qp[4], qtmp[4], np[4], dg[4];
               for (int m = 0; m < 4; m++) qq[m] = 0.5 * (q(im, jm, km, m) + q(i, j, k, m));
               qp[0] = qq[0], qp[1] = -qq[1], qp[2] = -qq[2], qp[3] = -qq[3];
               quatmult4(n, qp, qtmp);
               quatmult4(qq, qtmp, np);
               const double a2 = np[1] * np[1], a3 = np[2] * np[2], a4 = np[3] * np[3];
               n4 = a2 * a2 + a3 * a3 + a4 * a4;
               dg[0] = 0.;
               dg[1] = np[1] * (np[1] * np[1] - n4);
               dg[2] = np[2] * (np[2] * np[2] - n4);
               dg[3] = np[3] * (np[3] * np[3] - n4);
               quatmult4(dg, qq, qtmp);
               quatmult4(qp, qtmp, dgamma);
            } else {
               dgamma[0] = 0., dgamma[1] = 0.;
               dgamma[2] = (a == 2) ? 1. : 0.;
               dgamma[3] = (a == 2) ? 0. : 1.;
               n4 = 0.;
            }
            const double gamma = epsilon * (1. - 3. * eps4) * (1. + factor * n4);
            f(i, j, k) = gamma * gamma * g[a] + 16. * epsilon * gamma * eps4 * sqrt(gphi2) * dgamma[a + 1];
         });
      }
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_computerhspbg(int ndim, const int* ifirst, const int* ilast, const double* dx,
                         double misorientation_factor, double epsilonq, double* const* flux,
                         int ngflux, const double* temp, int ngtemp, double phi_well_scale,
                         double eta_well_scale, const double* phi, int ngphi, const double* eta, int ngeta,
                         const double* orient_grad_mod, int ngogm, double* rhs, int ngrhs,
                         const char* phi_well_type, const char* eta_well_type, const char* phi_interp_type,
                         const char* orient_interp_type1, const char* orient_interp_type2, int with_orient,
                         int three_phase, void* stream)
{
   // full argument list of COMPUTERHSPBG (QuatFort.h:90-110, 2d/quatrhs.m4:264-405): character arguments by
   // pointer as Fortran takes them, the eta block only dereferenced when three_phase != 0
   if (!phi_well_type || !orient_interp_type1 || !orient_interp_type2)
      return ampe_set_err(AMPE_EINVAL, "computerhspbg: null character argument");
   const char pwt = phi_well_type[0], oi1 = orient_interp_type1[0], oi2 = orient_interp_type2[0];
   if (pwt != 'd' && pwt != 's')
      return ampe_set_err(AMPE_EINVAL, "Error in deriv_well_func: type unknown");
   char ewt = 'd', pit = 'p';
   if (three_phase != 0) {
      if (!eta || !eta_well_type || !phi_interp_type)
         return ampe_set_err(AMPE_EINVAL, "computerhspbg: three_phase needs eta, eta_well_type and phi_interp_type");
      ewt = eta_well_type[0], pit = phi_interp_type[0];
      if (ewt != 'd' && ewt != 's') return ampe_set_err(AMPE_EINVAL, "Error in well_func: type unknown");
   }
   const Box b = mkbox(ndim, ifirst, ilast);
   const DV3 fl = sides(flux, b, ngflux);
   const CV T = view(temp, b, -1, ngtemp), ph = view(phi, b, -1, ngphi);
   const CV ogm = view(orient_grad_mod, b, -1, ngogm);
   const CV et = view(three_phase != 0 ? eta : phi, b, -1, three_phase != 0 ? ngeta : ngphi);
   const DV r = view(rhs, b, -1, ngrhs);
   double di[3] = {0, 0, 0};
   for (int d = 0; d < ndim; d++) di[d] = 1.0 / dx[d];
   const double d0 = di[0], d1 = di[1], d2 = di[2];
   const double epsilonq2 = 0.5 * epsilonq * epsilonq;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      double diff_term = (fl.a[0](i + 1, j, k) - fl.a[0](i, j, k)) * d0;
      diff_term = diff_term + (fl.a[1](i, j + 1, k) - fl.a[1](i, j, k)) * d1;
      if (ndim == 3) diff_term = diff_term + (fl.a[2](i, j, k + 1) - fl.a[2](i, j, k)) * d2;
      double v = diff_term;
      v = v - phi_well_scale * deriv_well_func(ph(i, j, k), pwt);
      if (three_phase != 0) {
         // eta energy well (2d/quatrhs.m4:356-372)
         const double h_prime = deriv_interp_func(ph(i, j, k), pit);
         v = v - eta_well_scale * h_prime * well_func(et(i, j, k), ewt);
      }
      if (with_orient != 0) {
         const double p1 = deriv_interp_func(ph(i, j, k), oi1);
         const double p2 = deriv_interp_func(ph(i, j, k), oi2);
         v = v - misorientation_factor * T(i, j, k) * p1 * ogm(i, j, k) -
             p2 * epsilonq2 * ogm(i, j, k) * ogm(i, j, k);
      }
      r(i, j, k) = v;
   });
}

int ampe_k_phaserhs_fenergy(int ndim, const int* ifirst, const int* ilast, const double* fl,
                            const double* fa, const double* phi, int ngphi, double* rhs, int ngrhs,
                            char energy_interp_type, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV l = view(fl, b, -1, 0), a = view(fa, b, -1, 0), ph = view(phi, b, -1, ngphi);
   const DV r = view(rhs, b, -1, ngrhs);
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      const double hp = deriv_interp_func(ph(i, j, k), energy_interp_type);
      r(i, j, k) = r(i, j, k) + hp * (l(i, j, k) - a(i, j, k));
   });
}

int ampe_k_computerhstemp(int ndim, const int* ifirst, const int* ilast, const double* dx,
                          double thermal_diffusivity, double latent_heat, const double* temp,
                          int ngtemp, const double* cp, int ngcp, int with_phase,
                          const double* phi_rhs, int ngphi_rhs, double* rhs, int ngrhs,
                          void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV T = view(temp, b, -1, ngtemp), c = view(cp, b, -1, ngcp);
   const CV pr = view(phi_rhs, b, -1, ngphi_rhs);
   const DV r = view(rhs, b, -1, ngrhs);
   double di2[3] = {0, 0, 0};
   for (int d = 0; d < ndim; d++) di2[d] = 1.0 / (dx[d] * dx[d]);
   const double d0 = di2[0], d1 = di2[1], d2 = di2[2];
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      // laplacian (2d/laplacian.m4:37-52)
      const double tx = (T(i - 1, j, k) - 2.0 * T(i, j, k) + T(i + 1, j, k));
      const double ty = (T(i, j - 1, k) - 2.0 * T(i, j, k) + T(i, j + 1, k));
      double dt = tx * d0 + ty * d1;
      if (ndim == 3) dt = dt + (T(i, j, k - 1) - 2.0 * T(i, j, k) + T(i, j, k + 1)) * d2;
      double v = thermal_diffusivity * dt;
      if (with_phase != 0) v = v + (latent_heat / c(i, j, k)) * pr(i, j, k);
      r(i, j, k) = v;
   });
}

int ampe_k_computerhsbiaswell(int ndim, const int* ifirst, const int* ilast, const double* phi,
                              int ngphi, const double* temp, int ngtemp, double alpha, double gamma,
                              const double* te, int ngte, double* rhs, int ngrhs, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(phi, b, -1, ngphi), T = view(temp, b, -1, ngtemp), e = view(te, b, -1, ngte);
   const DV r = view(rhs, b, -1, ngrhs);
   const double coeff = alpha / (double)(4.f * atanf(1.f));  // pi = 4.*atan(1.) is REAL*4
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      const double m = coeff * atan(gamma * (e(i, j, k) - T(i, j, k)));
      r(i, j, k) = r(i, j, k) + m * ph(i, j, k) * (1.0 - ph(i, j, k));
   });
}

// COMPUTERHSDELTATEMPERATURE (QuatFort.h:203; 2d/quatrhs.m4:893-940, 3d/quatrhs.m4:1040-1087)
int ampe_k_computerhsdeltatemperature(int ndim, const int* ifirst, const int* ilast, const double* phi, int ngphi,
                                      const double* temp, int ngtemp, double tm, double latentheat, double* rhs,
                                      int ngrhs, const char* energy_interp_type, void* stream)
{
   if (!energy_interp_type || !(tm > 0.0)) return ampe_set_err(AMPE_EINVAL, "computerhsdeltatemperature: bad argument");
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(phi, b, -1, ngphi), T = view(temp, b, -1, ngtemp);
   const DV r = view(rhs, b, -1, ngrhs);
   const double alpha = latentheat / tm;
   const double woff = (double)(0.25f / 6.f);  // REAL*4 expression in the 3D routine
   const char it = energy_interp_type[0];
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      double wtemp;
      if (ndim == 2)
         wtemp = 0.75 * T(i, j, k) + 0.0625 * (T(i - 1, j, k) + T(i, j - 1, k) + T(i + 1, j, k) + T(i, j + 1, k));
      else
         wtemp = 0.75 * T(i, j, k) + woff * (T(i - 1, j, k) + T(i, j - 1, k) + T(i + 1, j, k) + T(i, j + 1, k) +
                                             T(i, j, k - 1) + T(i, j, k + 1));
      const double m = alpha * (tm - wtemp);
      r(i, j, k) = r(i, j, k) + m * deriv_interp_func(ph(i, j, k), it);
   });
}

// ---- quatdiffs.m4 / quatgrad.m4 -------------------------------------------------------------
int ampe_k_quatdiffs(int ndim, const int* lo, const int* hi, int depth, const double* q, int ngq,
                     double* const* diff, int ngdiff, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const CV qq = view(q, b, -1, ngq);
   const DV3 d = sides(diff, b, ngdiff);
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 1, L, H);
      const DV da = d.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         for (int m = 0; m < depth; m++)
            da(i, j, k, m) = qq(i, j, k, m) - qq(i - E(a, 0), j - E(a, 1), k - E(a, 2), m);
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_quatdiffs_symm(int ndim, const int* lo, const int* hi, int depth, const double* q,
                          int ngq, double* const* diff, int ngdiff, const int* const* iqrot,
                          int ngiq, void* stream)
{
   int rc = ensure_qr();
   if (rc) return rc;
   const Box b = mkbox(ndim, lo, hi);
   const CV qq = view(q, b, -1, ngq);
   const DV3 d = sides(diff, b, ngdiff);
   const IV3 iq = isides(iqrot, b, ngiq);
   const double* qr = g_qr;
   const int* conj = g_conj;
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 1, L, H);
      const DV da = d.a[a];
      const IV ia = iq.a[a];
      rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         double q2[4], q2p[4];
         for (int m = 0; m < depth; m++) q2[m] = qq(i - E(a, 0), j - E(a, 1), k - E(a, 2), m);
         rot(q2, ia(i, j, k), q2p, depth, qr, conj);
         for (int m = 0; m < depth; m++) da(i, j, k, m) = qq(i, j, k, m) - q2p[m];
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_quatgrad_cell(int ndim, const int* lo, const int* hi, int depth, const double* h,
                         double* const* diff, int ngdiff, double* const* grad, int nggrad,
                         void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 d = sides(diff, b, ngdiff), g = cells3(grad, b, nggrad);
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   for (int a = 0; a < ndim; a++) {
      const double p5 = 0.5 / h[a];
      const DV da = d.a[a], ga = g.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         for (int m = 0; m < depth; m++)
            ga(i, j, k, m) = (da(i + E(a, 0), j + E(a, 1), k + E(a, 2), m) + da(i, j, k, m)) * p5;
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_quatgrad_cell_symm(int ndim, const int* lo, const int* hi, int depth, const double* h,
                              double* const* diff, int ngdiff, double* const* grad, int nggrad,
                              const int* const* iqrot, int ngiq, void* stream)
{
   int rc = ensure_qr();
   if (rc) return rc;
   const Box b = mkbox(ndim, lo, hi);
   const DV3 d = sides(diff, b, ngdiff), g = cells3(grad, b, nggrad);
   const IV3 iq = isides(iqrot, b, ngiq);
   const double* qr = g_qr;
   const int* conj = g_conj;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   for (int a = 0; a < ndim; a++) {
      const double p5 = 0.5 / h[a];
      const DV da = d.a[a], ga = g.a[a];
      const IV ia = iq.a[a];
      rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int ip = i + E(a, 0), jp = j + E(a, 1), kp = k + E(a, 2);
         double dt[4], dp[4];
         for (int m = 0; m < depth; m++) dt[m] = da(ip, jp, kp, m);
         if (depth > 1)
            rot(dt, -ia(ip, jp, kp), dp, depth, qr, conj);
         else
            dp[0] = dt[0];
         for (int m = 0; m < depth; m++) ga(i, j, k, m) = (dp[m] + da(i, j, k, m)) * p5;
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_quatgrad_side(int ndim, const int* lo, const int* hi, int depth, const double* h,
                         double* const* diff, int ngdiff, double* const* grad, int nggrad,
                         void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 d = sides(diff, b, ngdiff), g = sides(grad, b, nggrad);
   const double hi0 = 1.0 / h[0], hi1 = 1.0 / h[1], hi2 = ndim == 3 ? 1.0 / h[2] : 0.0;
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV ga = g.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const double hinv[3] = {hi0, hi1, hi2};
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         for (int m = 0; m < depth; m++)
            for (int t = 0; t < ndim; t++) {
               if (t == a) {
                  ga(i, j, k, t * depth + m) = hinv[a] * d.a[a](i, j, k, m);
               } else {
                  const double p25 = 0.25 * hinv[t];
                  const int t0 = E(t, 0), t1 = E(t, 1), t2 = E(t, 2);
                  const DV dt = d.a[t];
                  ga(i, j, k, t * depth + m) =
                      p25 * (dt(im + t0, jm + t1, km + t2, m) + dt(im, jm, km, m) +
                             dt(i + t0, j + t1, k + t2, m) + dt(i, j, k, m));
               }
            }
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_quatgrad_side_symm(int ndim, const int* lo, const int* hi, int depth, const double* h,
                              double* const* diff, int ngdiff, double* const* grad, int nggrad,
                              const int* const* iqrot, int ngiq, void* stream)
{
   int rc = ensure_qr();
   if (rc) return rc;
   const Box b = mkbox(ndim, lo, hi);
   const DV3 d = sides(diff, b, ngdiff), g = sides(grad, b, nggrad);
   const IV3 iq = isides(iqrot, b, ngiq);
   const double* qr = g_qr;
   const int* conj = g_conj;
   const double hi0 = 1.0 / h[0], hi1 = 1.0 / h[1], hi2 = ndim == 3 ? 1.0 / h[2] : 0.0;
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV ga = g.a[a];
      rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const double hinv[3] = {hi0, hi1, hi2};
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         for (int t = 0; t < ndim; t++) {
            if (t == a) continue;
            const double p25 = 0.25 * hinv[t];
            const int t0 = E(t, 0), t1 = E(t, 1), t2 = E(t, 2);
            const DV dt = d.a[t];
            double d1[4], d1p[4], d2[4], d2p[4], d3[4], d4[4], d4p[4];
            if (depth > 1) {
               for (int m = 0; m < depth; m++) {
                  d1[m] = dt(i + t0, j + t1, k + t2, m);
                  d2[m] = dt(im + t0, jm + t1, km + t2, m);
                  d3[m] = dt(im, jm, km, m);
               }
               rot(d1, -iq.a[t](i + t0, j + t1, k + t2), d1p, depth, qr, conj);
               rot(d2, -iq.a[t](im + t0, jm + t1, km + t2), d2p, depth, qr, conj);
               for (int m = 0; m < depth; m++) d4[m] = d2p[m] + d3[m];
               rot(d4, iq.a[a](i, j, k), d4p, depth, qr, conj);
            } else {
               d1p[0] = dt(i + t0, j + t1, k + t2, 0);
               d4p[0] = dt(im + t0, jm + t1, km + t2, 0) + dt(im, jm, km, 0);
            }
            for (int m = 0; m < depth; m++)
               ga(i, j, k, t * depth + m) = p25 * (d4p[m] + d1p[m] + dt(i, j, k, m));
         }
         for (int m = 0; m < depth; m++) ga(i, j, k, a * depth + m) = hinv[a] * d.a[a](i, j, k, m);
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_quatgrad_modulus(int ndim, const int* lo, const int* hi, int depth,
                            double* const* grad_cell, int nggq, double* grad_mod, int ngm,
                            void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 g = cells3(grad_cell, b, nggq);
   const DV gm = view(grad_mod, b, -1, ngm);
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      double s = 0.0;
      for (int m = 0; m < depth; m++) {
         s = s + g.a[0](i, j, k, m) * g.a[0](i, j, k, m) + g.a[1](i, j, k, m) * g.a[1](i, j, k, m);
         if (ndim == 3) s = s + g.a[2](i, j, k, m) * g.a[2](i, j, k, m);
      }
      gm(i, j, k) = sqrt(s);
   });
}

int ampe_k_quatgrad_modulus_from_sides_compact(int ndim, const int* lo, const int* hi, int depth,
                                               double* const* grad_side, int nggq,
                                               double* grad_mod, int ngm, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 g = sides(grad_side, b, nggq);
   const DV gm = view(grad_mod, b, -1, ngm);
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      double s = 0.0;
      for (int a = 0; a < ndim; a++) {
         for (int m = 0; m < depth; m++) {
            const double v = g.a[a](i, j, k, a * depth + m);
            s = s + v * v;
         }
         for (int m = 0; m < depth; m++) {
            const double v = g.a[a](i + E(a, 0), j + E(a, 1), k + E(a, 2), a * depth + m);
            s = s + v * v;
         }
      }
      gm(i, j, k) = sqrt(0.5 * s);
   });
}

// ---- quatfacops.m4 --------------------------------------------------------------------------
int ampe_k_compute_face_coef(int ndim, const int* lo, const int* hi, int depth, double eps_q,
                             const double* phi, int ngp, const double* temp, int ngt,
                             double misorientation_factor, double* const* gq, int nggq,
                             double* const* fc, int ngf, double gradient_floor, char floor_type,
                             char interp_type1, char interp_type2, char avg_type, void* stream)
{
   if (floor_type != 'm' && floor_type != 't' && floor_type != 's')
      return ampe_set_err(AMPE_EINVAL, "Error in eval_grad_normi: floor_type unknown");
   if (avg_type != 'a' && avg_type != 'h')
      return ampe_set_err(AMPE_EINVAL, "Error in average_func: type unknown");
   const Box b = mkbox(ndim, lo, hi);
   const CV ph = view(phi, b, -1, ngp), T = view(temp, b, -1, ngt);
   const DV3 g = sides(gq, b, nggq), f = sides(fc, b, ngf);
   const double floor2 = gradient_floor * gradient_floor, eps2 = eps_q * eps_q;
   const double maxn = 1.0 / gradient_floor;
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV ga = g.a[a], fa = f.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         const double phia = average_func(ph(im, jm, km), ph(i, j, k), avg_type);
         const double tempa = 0.5 * (T(im, jm, km) + T(i, j, k));
         const double diff = misorientation_factor * tempa * interp_func(phia, interp_type1);
         const double hphi2 = interp_func(phia, interp_type2);
         double g2 = 0.0;
         for (int n = 0; n < ndim; n++)
            for (int m = 0; m < depth; m++) {
               const double v = ga(i, j, k, n * depth + m);
               g2 = g2 + v * v;
            }
         const double normi = eval_grad_normi(g2, floor_type, floor2, maxn);
         fa(i, j, k) = -normi * diff - eps2 * hphi2;
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_compute_flux(int ndim, const int* lo, const int* hi, int depth, double* const* fc,
                        int ngfc, const double* q, int ngq, const double* h, double* const* flux,
                        int ngflux, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 c = sides(fc, b, ngfc), f = sides(flux, b, ngflux);
   const CV qq = view(q, b, -1, ngq);
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const double hinv = 1.0 / h[a];
      const DV ca = c.a[a], fa = f.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         for (int m = 0; m < depth; m++)
            fa(i, j, k, m) = ca(i, j, k) * hinv *
                             (qq(i, j, k, m) - qq(i - E(a, 0), j - E(a, 1), k - E(a, 2), m));
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_compute_flux_from_gradq(int ndim, const int* lo, const int* hi, int depth,
                                   double* const* fc, int ngfc, double* const* grad_side,
                                   double* const* flux, int ngflux, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 c = sides(fc, b, ngfc), g = sides(grad_side, b, 0), f = sides(flux, b, ngflux);
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV ca = c.a[a], ga = g.a[a], fa = f.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         for (int m = 0; m < depth; m++) fa(i, j, k, m) = ca(i, j, k) * ga(i, j, k, a * depth + m);
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_compute_lambda_flux(int ndim, const int* lo, const int* hi, int depth,
                               double* const* flux, int ngflux, const double* q, int ngq,
                               const double* h, double* lambda, int nglambda, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 f = sides(flux, b, ngflux);
   const CV qq = view(q, b, -1, ngq);
   const DV lam = view(lambda, b, -1, nglambda);
   const double x = 0.5 / h[0], y = 0.5 / h[1], z = ndim == 3 ? 0.5 / h[2] : 0.0;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      double l = 0.0, sumq2 = 0.0;
      for (int m = 0; m < depth; m++) {
         double s = (f.a[0](i + 1, j, k, m) - f.a[0](i, j, k, m)) * x +
                    (f.a[1](i, j + 1, k, m) - f.a[1](i, j, k, m)) * y;
         if (ndim == 3) s = s + (f.a[2](i, j, k + 1, m) - f.a[2](i, j, k, m)) * z;
         l = l - qq(i, j, k, m) * s;
         sumq2 = sumq2 + qq(i, j, k, m) * qq(i, j, k, m);
      }
      lam(i, j, k) = l / sumq2;
   });
}

int ampe_k_add_quat_proj_op(int ndim, const int* lo, const int* hi, int depth,
                            const double* mobility, int ngmob, double* const* flux, int ngflux,
                            const double* q, int ngq, const double* lambda, int nglambda,
                            const double* h, double* rhs, int ngrhs, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 f = sides(flux, b, ngflux);
   const CV mob = view(mobility, b, -1, ngmob), qq = view(q, b, -1, ngq);
   const CV lam = view(lambda, b, -1, nglambda);
   const DV r = view(rhs, b, -1, ngrhs);
   const double x = 1.0 / h[0], y = 1.0 / h[1], z = ndim == 3 ? 1.0 / h[2] : 0.0;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      for (int m = 0; m < depth; m++) {
         double dv = (f.a[0](i + 1, j, k, m) - f.a[0](i, j, k, m)) * x +
                     (f.a[1](i, j + 1, k, m) - f.a[1](i, j, k, m)) * y;
         if (ndim == 3) dv = dv + (f.a[2](i, j, k + 1, m) - f.a[2](i, j, k, m)) * z;
         r(i, j, k, m) = r(i, j, k, m) - mob(i, j, k) * (dv + 2.0 * qq(i, j, k, m) * lam(i, j, k));
      }
   });
}

int ampe_k_add_quat_op(int ndim, const int* lo, const int* hi, int depth, const double* mobility,
                       int ngmob, double* const* flux, int ngflux, const double* h, double* rhs,
                       int ngrhs, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV3 f = sides(flux, b, ngflux);
   const CV mob = view(mobility, b, -1, ngmob);
   const DV r = view(rhs, b, -1, ngrhs);
   const double x = 1.0 / h[0], y = 1.0 / h[1], z = ndim == 3 ? 1.0 / h[2] : 0.0;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      for (int m = 0; m < depth; m++) {
         double dv = (f.a[0](i + 1, j, k, m) - f.a[0](i, j, k, m)) * x +
                     (f.a[1](i, j + 1, k, m) - f.a[1](i, j, k, m)) * y;
         if (ndim == 3) dv = dv + (f.a[2](i, j, k + 1, m) - f.a[2](i, j, k, m)) * z;
         r(i, j, k, m) = r(i, j, k, m) - mob(i, j, k) * dv;
      }
   });
}

int ampe_k_correctrhsquatforsymmetry(int ndim, const int* lo, const int* hi, int depth,
                                     const double* dx, double* const* nonsymm_diff,
                                     double* const* symm_diff, int ngdiff, double* rhs, int ngrhs,
                                     const double* quat, int ngq, double* const* facecoeff,
                                     int ngfacecoeff, const double* mobility, int ngmob,
                                     const int* const* iqrot, int ngiq, void* stream)
{
   int rc = ensure_qr();
   if (rc) return rc;
   const Box b = mkbox(ndim, lo, hi);
   const DV3 nsd = sides(nonsymm_diff, b, ngdiff), sd = sides(symm_diff, b, ngdiff);
   const DV3 fc = sides(facecoeff, b, ngfacecoeff);
   const IV3 iq = isides(iqrot, b, ngiq);
   const DV r = view(rhs, b, -1, ngrhs);
   const CV qq = view(quat, b, -1, ngq), mob = view(mobility, b, -1, ngmob);
   const double* qr = g_qr;
   const int* conj = g_conj;
   const double i0 = 1.0 / (dx[0] * dx[0]), i1 = 1.0 / (dx[1] * dx[1]);
   const double i2 = ndim == 3 ? 1.0 / (dx[2] * dx[2]) : 0.0;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      const double inv2[3] = {i0, i1, i2};
      double dpr[3][4], tmp[4], dt[4];
      for (int a = 0; a < ndim; a++) {
         const int ip = i + E(a, 0), jp = j + E(a, 1), kp = k + E(a, 2);
         for (int m = 0; m < depth; m++) dt[m] = sd.a[a](ip, jp, kp, m);
         if (depth > 1)
            rot(dt, -iq.a[a](ip, jp, kp), dpr[a], depth, qr, conj);
         else
            dpr[a][0] = dt[0];
      }
      for (int m = 0; m < depth; m++) {
         double t = 0.0;
         for (int a = 0; a < ndim; a++) {
            const int ip = i + E(a, 0), jp = j + E(a, 1), kp = k + E(a, 2);
            const double term =
                inv2[a] * (fc.a[a](ip, jp, kp) * (nsd.a[a](ip, jp, kp, m) - dpr[a][m]) -
                           fc.a[a](i, j, k) * (nsd.a[a](i, j, k, m) - sd.a[a](i, j, k, m)));
            t = (a == 0) ? term : t + term;
         }
         tmp[m] = t;
      }
      if (depth > 1) {
         double beta = 0.0, lambda = 0.0;
         for (int m = 0; m < depth; m++) {
            beta = beta + qq(i, j, k, m) * qq(i, j, k, m);
            lambda = lambda + qq(i, j, k, m) * tmp[m];
         }
         lambda = lambda / beta;
         for (int m = 0; m < depth; m++)
            r(i, j, k, m) = r(i, j, k, m) + mob(i, j, k) * (tmp[m] - lambda * qq(i, j, k, m));
      } else {
         r(i, j, k, 0) = r(i, j, k, 0) + mob(i, j, k) * tmp[0];
      }
   });
}

int ampe_k_quatmobility(int ndim, const int* ifirst, const int* ilast, const double* phase,
                        int ngphase, double* mobility, int ngmobility, double scale_mobility,
                        double min_mobility, char func_type, double alt_scale_factor, void* stream)
{
   const char f = func_type;
   if (f != 'p' && f != 'P' && f != 'e' && f != 'E' && f != 'i' && f != 'I')
      return ampe_set_err(AMPE_EINVAL, "Error in quatmobility: unknown function type");
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(phase, b, -1, ngphase);
   const DV mob = view(mobility, b, -1, ngmobility);
   int L[3], H[3];
   cell_bounds(b, ngmobility, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      mob(i, j, k) = quat_mobility(ph(i, j, k), func_type, scale_mobility, min_mobility,
                                   alt_scale_factor);
   });
}

// ---- composition ------------------------------------------------------------------------------
int ampe_k_concentrationflux(int ndim, const int* ifirst, const int* ilast, const double* dx,
                             const double* conc, int ngconc, const double* phi, int ngphi,
                             double* const* diffconc, int ngdiffconc, double* const* dphicoupl,
                             int ngdphicoupl, double* const* flux, int ngflux, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV c = view(conc, b, -1, ngconc), ph = view(phi, b, -1, ngphi);
   const DV3 D = sides(diffconc, b, ngdiffconc), P = sides(dphicoupl, b, ngdphicoupl);
   const DV3 f = sides(flux, b, ngflux);
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const double dinv = 1.0 / dx[a];
      const DV Da = D.a[a], Pa = P.a[a], fa = f.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         fa(i, j, k) = dinv * (Da(i, j, k) * (c(i, j, k) - c(im, jm, km)) +
                               Pa(i, j, k) * (ph(i, j, k) - ph(im, jm, km)));
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_add_cahnhilliarddoublewell_flux(int ndim, const int* ifirst, const int* ilast,
                                           const double* dx, const double* conc, int ngconc,
                                           double mobility, double ca, double cb,
                                           double well_scale, double kappa, double* const* flux,
                                           int ngflux, void* stream)
{
   // the reference scatters +-M/h mu(cell) onto the faces of cells box+1; here every face of
   // the box gathers its two contributions in the reference's summation order
   // (flux += 0 - M/h mu(i-1), then + M/h mu(i)): faces between cells of box+1 are complete
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV c = view(conc, b, -1, ngconc);
   const DV3 f = sides(flux, b, ngflux);
   double di[3] = {0, 0, 0};
   for (int d = 0; d < ndim; d++) di[d] = 1.0 / dx[d];
   const double d0 = di[0], d1 = di[1], d2 = di[2];
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV fa = f.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const double dinv[3] = {d0, d1, d2};
         auto mu = [&](int ii, int jj, int kk) {
            double lap = dinv[0] * dinv[0] * (-2.0 * c(ii, jj, kk) + c(ii - 1, jj, kk) + c(ii + 1, jj, kk)) +
                         dinv[1] * dinv[1] * (-2.0 * c(ii, jj, kk) + c(ii, jj - 1, kk) + c(ii, jj + 1, kk));
            if (ndim == 3)
               lap = lap + dinv[2] * dinv[2] *
                               (-2.0 * c(ii, jj, kk) + c(ii, jj, kk - 1) + c(ii, jj, kk + 1));
            const double cc = c(ii, jj, kk);
            return 2.0 * well_scale * (cc - ca) * (cb - cc) * (cb + ca - 2.0 * cc) - kappa * lap;
         };
         const double mm = mu(i - E(a, 0), j - E(a, 1), k - E(a, 2));
         const double mc = mu(i, j, k);
         fa(i, j, k) = (fa(i, j, k) - mobility * dinv[a] * mm) + mobility * dinv[a] * mc;
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_add_flux(int ndim, const int* ifirst, const int* ilast, const double* dx,
                    const double* conc, int ngconc, int ncomp, double* const* diffconc, int ngdiff,
                    double* const* flux, int ngflux, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV c = view(conc, b, -1, ngconc);
   const DV3 D = sides(diffconc, b, ngdiff), f = sides(flux, b, ngflux);
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const double dinv = 1.0 / dx[a];
      const DV Da = D.a[a], fa = f.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         for (int ic = 0; ic < ncomp; ic++)
            for (int jc = 0; jc < ncomp; jc++) {
               const int ijc = ic + jc * ncomp;
               fa(i, j, k, ic) = fa(i, j, k, ic) +
                                 dinv * (Da(i, j, k, ijc) *
                                         (c(i, j, k, jc) - c(i - E(a, 0), j - E(a, 1), k - E(a, 2), jc)));
            }
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_concentration_pfmdiffusion(int ndim, const int* ifirst, const int* ilast,
                                      const double* phi, int ngphi, double* const* diff, int ngdiff,
                                      const double* temp, int ngtemp, double d_liquid,
                                      double q0_liquid, double d_solid_A, double q0_solid_A,
                                      double gas_constant_R, char interp_type, char avg_type,
                                      void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(phi, b, -1, ngphi), T = view(temp, b, -1, ngtemp);
   const DV3 D = sides(diff, b, ngdiff);
   const double ql = q0_liquid / gas_constant_R, qs = q0_solid_A / gas_constant_R;
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV Da = D.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         const double vphi = average_func(ph(im, jm, km), ph(i, j, k), avg_type);
         const double hphi = interp_func(vphi, interp_type);
         const double invT = 2.0 / (T(im, jm, km) + T(i, j, k));
         const double dl = d_liquid * exp(-ql * invT);
         const double ds = d_solid_A * exp(-qs * invT);
         Da(i, j, k) = (1.0 - hphi) * dl + hphi * ds;
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_computerhsconcentration(int ndim, const int* ifirst, const int* ilast, const double* dx,
                                   double* const* flux, int ngflux, double mobility, double* rhs,
                                   int ngrhs, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const DV3 f = sides(flux, b, ngflux);
   const DV r = view(rhs, b, -1, ngrhs);
   const double x = 1.0 / dx[0], y = 1.0 / dx[1], z = ndim == 3 ? 1.0 / dx[2] : 0.0;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      double s = x * (f.a[0](i + 1, j, k) - f.a[0](i, j, k)) + y * (f.a[1](i, j + 1, k) - f.a[1](i, j, k));
      if (ndim == 3) s = s + z * (f.a[2](i, j, k + 1) - f.a[2](i, j, k));
      r(i, j, k) = mobility * s;
   });
}

// ---- per-patch loops of the C++ strategies -----------------------------------------------------
static int nfail_counter(int** dev)
{
   static int* d = nullptr;
   if (!d) {
      if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return ampe_set_err(AMPE_ECUDA, "cudaMalloc");
      cudaMemset(d, 0, sizeof(int));
   }
   *dev = d;
   return AMPE_OK;
}

int ampe_k_compute_phase_concentrations(const ampe_rhs_config* cfg, const int* ifirst,
                                        const int* ilast, const double* phi, int ngphi,
                                        const double* conc, int ngconc, const double* cl_ref,
                                        const double* ca_ref, double* cl, double* ca, int ngc,
                                        void* stream)
{
   Params p;
   int rc = ampe_derive_params(*cfg, p);
   if (rc) return rc;
   const Box b = mkbox(cfg->ndim, ifirst, ilast);
   const CV ph = view(phi, b, -1, ngphi), c = view(conc, b, -1, ngconc);
   const CV lr = view(cl_ref, b, -1, ngc), ar = view(ca_ref, b, -1, ngc);
   const DV l = view(cl, b, -1, ngc), a = view(ca, b, -1, ngc);
   int* nfail;
   rc = nfail_counter(&nfail);
   if (rc) return rc;
   int L[3], H[3];
   cell_bounds(b, ngc, L, H);
   const bool calphad = cfg->free_energy == AMPE_FE_CALPHAD;
   rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      const double hphi = interp_func(ph(i, j, k), p.conc_interp);
      double x0, x1;
      if (calphad) {
         x0 = lr(i, j, k);
         x1 = ar(i, j, k);
         KksFinal fin;
         if (kks_newton(p.ct, c(i, j, k), hphi, x0, x1, p.newton_tol, p.newton_max_its,
                        p.newton_alpha, fin) < 0)
            atomicAdd(nfail, 1);
      } else {
         const double h = clamp01(hphi), cc = c(i, j, k);
         x0 = (cc - h * (p.quad_ceq[1] - p.quad_rla * p.quad_ceq[0])) / ((1.0 - h) + h * p.quad_rla);
         x1 = (cc - (1.0 - h) * (p.quad_ceq[0] - p.quad_ral * p.quad_ceq[1])) /
              ((1.0 - h) * p.quad_ral + h);
      }
      l(i, j, k) = x0;
      a(i, j, k) = x1;
   });
   if (rc) return rc;
   int h = 0;
   cudaMemcpyAsync(&h, nfail, sizeof(int), cudaMemcpyDeviceToHost, ST(stream));
   cudaStreamSynchronize(ST(stream));
   cudaMemsetAsync(nfail, 0, sizeof(int), ST(stream));
   return h;
}

int ampe_k_compute_free_energy(const ampe_rhs_config* cfg, const int* ifirst, const int* ilast,
                               const double* c_i, int ngc, double* f, int phase, void* stream)
{
   Params p;
   int rc = ampe_derive_params(*cfg, p);
   if (rc) return rc;
   const Box b = mkbox(cfg->ndim, ifirst, ilast);
   const CV c = view(c_i, b, -1, ngc);
   const DV fo = view(f, b, -1, 0);
   const bool calphad = cfg->free_energy == AMPE_FE_CALPHAD;
   const double inv_vm = phase == 0 ? p.inv_vm_l : p.inv_vm_a;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      const double ci = c(i, j, k);
      double v = calphad ? calphad_f(p.ct, ci, phase)
                         : p.quad_A[phase] * (ci - p.quad_ceq[phase]) * (ci - p.quad_ceq[phase]);
      v *= inv_vm;
      fo(i, j, k) = v;
   });
}

int ampe_k_add_driving_force(const ampe_rhs_config* cfg, const int* ifirst, const int* ilast,
                             const double* phi, int ngphi, const double* fl, const double* fa,
                             const double* cl, const double* ca, int ngc, double* rhs, int ngrhs,
                             void* stream)
{
   Params p;
   int rc = ampe_derive_params(*cfg, p);
   if (rc) return rc;
   const Box b = mkbox(cfg->ndim, ifirst, ilast);
   const CV ph = view(phi, b, -1, ngphi), l = view(fl, b, -1, 0), a = view(fa, b, -1, 0);
   const CV c_l = view(cl, b, -1, ngc), c_a = view(ca, b, -1, ngc);
   const DV r = view(rhs, b, -1, ngrhs);
   const bool calphad = cfg->free_energy == AMPE_FE_CALPHAD;
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      const double hp = deriv_interp_func(ph(i, j, k), p.energy_interp);
      double mu;
      if (calphad) {
         mu = calphad_mu(p.ct, c_a(i, j, k), 1);
         mu *= p.inv_vm_a;
      } else {
         mu = (2. * p.quad_A[0] * (c_l(i, j, k) - p.quad_ceq[0])) * p.inv_vm_l;
      }
      r(i, j, k) += hp * ((l(i, j, k) - a(i, j, k)) - mu * (c_l(i, j, k) - c_a(i, j, k)));
   });
}

int ampe_k_set_ebs_diffusion(const ampe_rhs_config* cfg, const int* ifirst, const int* ilast,
                             const double* phi, int ngphi, const double* cl, const double* ca,
                             int ngc, double* const* diff_l, double* const* diff_a, void* stream)
{
   Params p;
   int rc = ampe_derive_params(*cfg, p);
   if (rc) return rc;
   const Box b = mkbox(cfg->ndim, ifirst, ilast);
   const CV ph = view(phi, b, -1, ngphi), l = view(cl, b, -1, ngc), a = view(ca, b, -1, ngc);
   const DV3 Dl = sides(diff_l, b, 0), Da = sides(diff_a, b, 0);
   for (int ax = 0; ax < cfg->ndim; ax++) {
      int L[3], H[3];
      side_bounds(b, ax, 0, L, H);
      const DV dl = Dl.a[ax], da = Da.a[ax];
      const bool tbased = cfg->free_energy != AMPE_FE_CALPHAD;
      rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int im = i - E(ax, 0), jm = j - E(ax, 1), km = k - E(ax, 2);
         if (tbased) {
            // diffusion_type "temperature_dependent": concentration_pfmdiffusion_of_temperature
            // (2d/concentrationdiffusion.m4:341-430) at the uniform temperature
            const double vphi = average_func(ph(im, jm, km), ph(i, j, k), p.avg_func);
            const double hphi = interp_func(vphi, p.diffusion_interp);
            dl(i, j, k) = (1.0 - hphi) * p.D_liquid;
            da(i, j, k) = hphi * p.D_solid;
            return;
         }
         const double c_l = 0.5 * (l(i, j, k) + l(im, jm, km));
         const double c_a = 0.5 * (a(i, j, k) + a(im, jm, km));
         const double vl = diffusion_mobility(p.ct, 0, c_l) * calphad_d2f(p.ct, c_l, 0);
         const double va = diffusion_mobility(p.ct, 1, c_a) * calphad_d2f(p.ct, c_a, 1);
         const double phia = average_func(ph(i, j, k), ph(im, jm, km), p.conc_avg_func);
         const double hphi = interp_func(phia, p.diffusion_interp);
         dl(i, j, k) = (1. - hphi) * vl;
         da(i, j, k) = hphi * va;
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_set_kks_phase_diffusion(const ampe_rhs_config* cfg, const int* ifirst, const int* ilast,
                                   const double* phi, int ngphi, const double* cl, const double* ca,
                                   int ngc, double* const* d0, double* const* dphi, void* stream)
{
   const Box b = mkbox(cfg->ndim, ifirst, ilast);
   const CV ph = view(phi, b, -1, ngphi), l = view(cl, b, -1, ngc), a = view(ca, b, -1, ngc);
   const DV3 D0 = sides(d0, b, 0), DP = sides(dphi, b, 0);
   const char interp = cfg->energy_interp, avg = cfg->conc_avg_func;
   for (int ax = 0; ax < cfg->ndim; ax++) {
      int L[3], H[3];
      side_bounds(b, ax, 0, L, H);
      const DV d = D0.a[ax], dp = DP.a[ax];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int im = i - E(ax, 0), jm = j - E(ax, 1), km = k - E(ax, 2);
         const double phia = average_func(ph(i, j, k), ph(im, jm, km), avg);
         const double c_l = 0.5 * (l(i, j, k) + l(im, jm, km));
         const double c_a = 0.5 * (a(i, j, k) + a(im, jm, km));
         dp(i, j, k) = d(i, j, k) * deriv_interp_func(phia, interp) * (c_l - c_a);
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_fill_periodic(int ndim, const int* ifirst, const int* ilast, int depth,
                         const double* src, double* dst, int ng, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV s = view(src, b, -1, 0);
   const DV d = view(dst, b, -1, ng);
   int L[3], H[3];
   cell_bounds(b, ng, L, H);
   const int n0 = b.hi[0] - b.lo[0] + 1, n1 = b.hi[1] - b.lo[1] + 1, n2 = b.hi[2] - b.lo[2] + 1;
   const int l0 = b.lo[0], l1 = b.lo[1], l2 = b.lo[2];
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      const int is = l0 + ((i - l0) % n0 + n0) % n0, js = l1 + ((j - l1) % n1 + n1) % n1;
      const int ks = l2 + ((k - l2) % n2 + n2) % n2;
      for (int m = 0; m < depth; m++) d(i, j, k, m) = s(is, js, ks, m);
   });
}

int ampe_k_cell_multiply(int ndim, const int* ifirst, const int* ilast, const double* a, int nga,
                         const double* b, int ngb, double* dst, int ngdst, void* stream)
{
   const Box bx = mkbox(ndim, ifirst, ilast);
   const CV A = view(a, bx, -1, nga), B = view(b, bx, -1, ngb);
   const DV D = view(dst, bx, -1, ngdst);
   int L[3], H[3];
   cell_bounds(bx, 0, L, H);
   return for_box(L, H, ST(stream),
                  [=] __device__(int i, int j, int k) { D(i, j, k) = A(i, j, k) * B(i, j, k); });
}

int ampe_k_fill_periodic_int(int ndim, const int* ifirst, const int* ilast, int axis,
                             const int* src, int* dst, int ng, void* stream)
{
   const Box b = mkbox(ndim, ifirst, ilast);
   const IV s = view(src, b, -1, 0);  // ghost-0, one value per lower face of each cell
   V<int> d = view(dst, b, axis, ng);
   int L[3], H[3];
   side_bounds(b, axis, ng, L, H);
   L[axis] -= ng;
   H[axis] += ng;
   const int n0 = b.hi[0] - b.lo[0] + 1, n1 = b.hi[1] - b.lo[1] + 1, n2 = b.hi[2] - b.lo[2] + 1;
   const int l0 = b.lo[0], l1 = b.lo[1], l2 = b.lo[2];
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      const int is = l0 + ((i - l0) % n0 + n0) % n0, js = l1 + ((j - l1) % n1 + n1) % n1;
      const int ks = l2 + ((k - l2) % n2 + n2) % n2;
      d(i, j, k) = s(is, js, ks);
   });
}

}  // extern "C"
