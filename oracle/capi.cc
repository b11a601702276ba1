// TEST INFRASTRUCTURE ONLY (see oracle.h).  C entry points for ctypes (tests/,
// smoke(), bench.py's cpu_baseline / --impl reference legs).
#include "oracle.h"
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

extern "C" {

void* oracle_create(const ampe_rhs_config* cfg) { return create(*cfg); }
void oracle_destroy(void* c) { destroy((Ctx*)c); }
void oracle_set_ref(void* c, const double* cl, const double* ca) { set_ref((Ctx*)c, cl, ca); }
void oracle_set_rotations(void* c, const int* const* iq) { set_rotations((Ctx*)c, iq); }
int oracle_eval(void* c, double t, const ampe_rhs_fields* y, const ampe_rhs_fields* ydot,
                int fd_flag)
{
   return eval((Ctx*)c, t, y, ydot, fd_flag);
}
void oracle_get_phase_concentrations(void* c, double* cl, double* ca)
{
   get_phase_concentrations((Ctx*)c, cl, ca);
}
int oracle_energy(void* c, const ampe_rhs_fields* y, double* out) { return energy((Ctx*)c, y, out); }
int oracle_abi_sizeof_config() { return (int)sizeof(ampe_rhs_config); }
int oracle_num_threads()
{
#ifdef _OPENMP
   return omp_get_max_threads();
#else
   return 1;
#endif
}

// torchrun exports OMP_NUM_THREADS=1; the CPU baseline asks for the host's cores explicitly
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
   if (n > 0) omp_set_num_threads(n);
#else
   (void)n;
#endif
}

// ---- pointwise (functions.f / quat.f) ----
double oracle_interp_func(double phi, char t) { return interp_func(phi, t); }
double oracle_deriv_interp_func(double phi, char t) { return deriv_interp_func(phi, t); }
double oracle_second_deriv_interp_func(double phi, char t)
{
   return second_deriv_interp_func(phi, t);
}
double oracle_well_func(double phi, char t) { return well_func(phi, t); }
double oracle_deriv_well_func(double phi, char t) { return deriv_well_func(phi, t); }
double oracle_average_func(double a, double b, char t) { return average_func(a, b, t); }
double oracle_interp_ratio_func(double phi, char a, char b)
{
   return interp_ratio_func(phi, a, b);
}
double oracle_compl_interp_ratio_func(double phi, char a, char b)
{
   return compl_interp_ratio_func(phi, a, b);
}
double oracle_eval_grad_normi(double g2, char t, double f2, double mx)
{
   return eval_grad_normi(g2, t, f2, mx);
}
void oracle_quatsymmrotate(const double* q, int iq, double* qp, int qlen)
{
   quatsymmrotate(q, iq, qp, qlen);
}
// quatfindsymm (quat.f:9-37): returns the rotation index (iq is in/out in the reference)
int oracle_quatfindsymm(const double* q1, const double* q2, int iq, double* q2p, int qlen)
{
   quatfindsymm(q1, q2, &iq, q2p, qlen);
   return iq;
}
void oracle_qr_table4(double* out) { memcpy(out, qr_table4(), 48 * 4 * sizeof(double)); }

// ---- kernel-level wrappers (SAMRAI layouts) for the reference's KATs ----
static Box mkbox(int ndim, const int* lo, const int* hi)
{
   Box b;
   b.ndim = ndim;
   for (int d = 0; d < 3; d++) {
      b.lo[d] = d < ndim ? lo[d] : 0;
      b.hi[d] = d < ndim ? hi[d] : 0;
   }
   return b;
}
// QUAT_SYMM_ROTATION (QuatFort.h:319), QUAT_FUNDAMENTAL (:331), PROJECT{2,3}D (:883, :1022)
void oracle_k_quat_symm_rotation(int ndim, const int* lo, const int* hi, double* q, int ngq, int depth,
                                 int* const* rot, int ngrot)
{
   Box b = mkbox(ndim, lo, hi);
   IView r[3];
   for (int a = 0; a < ndim; a++) r[a] = make_iview(rot[a], b, a, ngrot);
   quat_symm_rotation(b, make_view(q, b, -1, ngq, depth), depth, r);
}
void oracle_k_quat_fundamental(int ndim, const int* lo, const int* hi, double* q, int ngq, int depth)
{
   Box b = mkbox(ndim, lo, hi);
   quat_fundamental(b, make_view(q, b, -1, ngq, depth), depth);
}
void oracle_k_project(int ndim, const int* lo, const int* hi, int depth, double* q, int ngq,
                      double* corr, int ngc, double* err, int nge)
{
   Box b = mkbox(ndim, lo, hi);
   project(b, depth, make_view(q, b, -1, ngq, depth), make_view(corr, b, -1, ngc, depth),
           make_view(err, b, -1, nge, depth));
}
// QUATDIFFS (QuatFort.h), tests/testGradQ.cc
void oracle_k_quatdiffs(int ndim, const int* lo, const int* hi, int depth, double* q, int ngq,
                        double* const* diff, int ngdiff)
{
   Box b = mkbox(ndim, lo, hi);
   View d[3];
   for (int a = 0; a < ndim; a++) d[a] = make_view(diff[a], b, a, ngdiff, depth);
   quatdiffs(b, depth, make_view(q, b, -1, ngq, depth), d);
}
void oracle_k_quatgrad_cell(int ndim, const int* lo, const int* hi, int depth, const double* h,
                            double* const* diff, int ngdiff, double* const* grad, int nggrad)
{
   Box b = mkbox(ndim, lo, hi);
   View d[3], g[3];
   for (int a = 0; a < ndim; a++) {
      d[a] = make_view(diff[a], b, a, ngdiff, depth);
      g[a] = make_view(grad[a], b, -1, nggrad, depth);
   }
   quatgrad_cell(b, depth, h, d, g);
}
// grad[a]: side array of axis a with depth ndim*depth (dir-major)
void oracle_k_quatgrad_side(int ndim, const int* lo, const int* hi, int depth, const double* h,
                            double* const* diff, int ngdiff, double* const* grad, int nggrad)
{
   Box b = mkbox(ndim, lo, hi);
   View d[3], g[3];
   for (int a = 0; a < ndim; a++) {
      d[a] = make_view(diff[a], b, a, ngdiff, depth);
      g[a] = make_view(grad[a], b, a, nggrad, ndim * depth);
   }
   quatgrad_side(b, depth, h, d, g);
}
// ADD_FLUX (ConcFort.h:139), tests/testFlux.cc pattern
void oracle_k_add_flux(int ndim, const int* lo, const int* hi, const double* dx, double* conc,
                       int ngconc, int ncomp, double* const* diffconc, int ngdiff,
                       double* const* flux, int ngflux)
{
   Box b = mkbox(ndim, lo, hi);
   View d[3], f[3];
   for (int a = 0; a < ndim; a++) {
      d[a] = make_view(diffconc[a], b, a, ngdiff, ncomp * ncomp);
      f[a] = make_view(flux[a], b, a, ngflux, ncomp);
   }
   add_flux(b, dx, make_view(conc, b, -1, ngconc, ncomp), ncomp, d, f);
}

// ---- Thermo4PFM stand-in ----
double oracle_calphad_free_energy(const ampe_calphad_binary* db, double T, double c, int pi)
{
   return calphad_free_energy(*db, T, c, pi);
}
double oracle_calphad_deriv_free_energy(const ampe_calphad_binary* db, double T, double c,
                                        int pi)
{
   return calphad_deriv_free_energy(*db, T, c, pi);
}
double oracle_calphad_second_deriv_free_energy(const ampe_calphad_binary* db, double T,
                                               double c, int pi)
{
   return calphad_second_deriv_free_energy(*db, T, c, pi);
}
int oracle_calphad_phase_concentrations(const ampe_calphad_binary* db, double T, double c0,
                                        double hphi, double* x, double tol, int max_its,
                                        double alpha)
{
   return calphad_phase_concentrations(*db, T, c0, hphi, x, tol, max_its, alpha);
}
int oracle_calphad_ceq(const ampe_calphad_binary* db, double T, double* ceq, double tol,
                       int max_its, double alpha)
{
   return calphad_ceq(*db, T, ceq, tol, max_its, alpha);
}
double oracle_calphad_fmix(double l0, double l1, double l2, double l3, double c)
{
   return calphad_fmix(l0, l1, l2, l3, c);
}
double oracle_calphad_fmix_deriv(double l0, double l1, double l2, double l3, double c)
{
   return calphad_fmix_deriv(l0, l1, l2, l3, c);
}
double oracle_calphad_fmix_deriv2(double l0, double l1, double l2, double l3, double c)
{
   return calphad_fmix_deriv2(l0, l1, l2, l3, c);
}
double oracle_xlogx(double x) { return xlogx(x); }
double oracle_xlogx_deriv(double x) { return xlogx_deriv(x); }
double oracle_xlogx_deriv2(double x) { return xlogx_deriv2(x); }
double oracle_calphad_diffusion_mobility(const ampe_calphad_binary* db, int phase, double c0,
                                         double T)
{
   return calphad_diffusion_mobility_binary(*db, phase, c0, T);
}
}
