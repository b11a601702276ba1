// TEST INFRASTRUCTURE ONLY (see oracle.h).  C entry points for ctypes (tests/,
// smoke(), bench.py's cpu_baseline / --impl reference legs).
#include "ctx.h"
#include "oracle.h"
#include "precond.h"
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
int oracle_scalar_diagnostics(void* c, const ampe_rhs_fields* y, double* out)
{
   return scalar_diagnostics((Ctx*)c, y, out);
}
// ---- block preconditioners (precond.cc) --------------------------------------------------------
// ncycles > 0: oracle_integrate_implicit runs right-preconditioned GMRES with that many V-cycles
void oracle_set_preconditioner(void* c, int ncycles, int has_dquatdphi, int left)
{
   ((Ctx*)c)->precond_cycles = ncycles;
   ((Ctx*)c)->precond_dquatdphi = has_dquatdphi != 0;
   ((Ctx*)c)->precond_left = left != 0;
}
void oracle_precond_stats(void* c, double* out2)
{
   out2[0] = ((Ctx*)c)->precond_stats[0];
   out2[1] = ((Ctx*)c)->precond_stats[1];
}
// after an fd_flag = 0 evaluation at the state the coefficients are frozen at
int oracle_precond_setup(void* c, double gamma, int ncycles, int has_dquatdphi)
{
   return precond_setup((Ctx*)c, gamma, ncycles, has_dquatdphi != 0);
}
// QuatFACOps::multiplyDQuatDPhiBlock: out (depth qlen, ghost 0) = [dF_q/dphi] z_phase
int oracle_precond_dquatdphi(void* c, const double* z_phase, double* out)
{
   return precond_dquatdphi((Ctx*)c, z_phase, out);
}
int oracle_precond_solve(void* c, const ampe_rhs_fields* r, const ampe_rhs_fields* z)
{
   return precond_solve((Ctx*)c, r, z);
}
// restated reference operator of one block: 0 phase, 1 quaternion component, 2 composition, 3 temperature
int oracle_precond_apply(void* c, int block, const double* u, double* out)
{
   return precond_apply((Ctx*)c, block, u, out);
}
// the multigrid of one block of the context (borrowed handle for the oracle_mg_* calls below)
void* oracle_precond_block(void* c, int block) { return precond_block((Ctx*)c, block); }

// host loop over the product's per-cell multigrid functions, same calls as ampe_mg_* with HOST arrays
void* oracle_mg_create(int ndim, const int* n, const double* dx, int with_s, int ncomp)
{
   return new HostMG(ndim, n, dx, with_s != 0, ncomp);
}
int oracle_mg_num_components(void* g) { return ((HostMG*)g)->numComponents(); }
void oracle_mg_destroy(void* g) { delete (HostMG*)g; }
int oracle_mg_set_elliptic(void* g, const double* m, int ngm, double m_const, const double* c, int ngc,
                           double c_const, const double* const* d, const double* const* d2, int ngd,
                           double d_scale, double d_const)
{
   try {
      ((HostMG*)g)->setElliptic(m, ngm, m_const, c, ngc, c_const, d, d2, ngd, d_scale, d_const);
      return 0;
   } catch (...) {
      return -1;
   }
}
int oracle_mg_set_quat(void* g, double gamma, const double* mobility, int ngm, const double* const* fc, int ngfc)
{
   try {
      ((HostMG*)g)->setQuat(gamma, mobility, ngm, fc, ngfc);
      return 0;
   } catch (...) {
      return -1;
   }
}
int oracle_mg_solve(void* g, const double* rhs, double* soln, int ncycles, int symmetrized)
{
   try {
      ((HostMG*)g)->solve(rhs, soln, ncycles, symmetrized != 0);
      return 0;
   } catch (...) {
      return -1;
   }
}
void oracle_mg_apply(void* g, const double* u, double* out) { ((HostMG*)g)->apply(u, out); }
void oracle_mg_set_sweeps(void* g, int pre, int post, int coarse) { ((HostMG*)g)->setSweeps(pre, post, coarse); }
void oracle_mg_set_zero_slope(void* g, const int* zero_slope) { ((HostMG*)g)->setZeroSlope(zero_slope); }
int oracle_mg_set_fused(void* g, int on, long long min_cells)
{
   ((HostMG*)g)->setFused(on != 0, min_cells);
   return ((HostMG*)g)->fusedLevels();
}
int oracle_mg_num_levels(void* g) { return ((HostMG*)g)->numLevels(); }
void oracle_mg_level_extents(void* g, int level, int* n_out)
{
   const int* n = ((HostMG*)g)->levelExtents(level);
   n_out[0] = n[0], n_out[1] = n[1], n_out[2] = n[2];
}
int oracle_mg_copy_level(void* g, int level, int which, double* out)
{
   HostMG* mg = (HostMG*)g;
   const double* src = mg->levelArray(level, which);
   if (!src) return -1;
   const int* n = mg->levelExtents(level);
   memcpy(out, src, sizeof(double) * (size_t)n[0] * n[1] * n[2]);
   return 0;
}
// the restated reference operators on caller-supplied SAMRAI-layout arrays over the box [0, n-1]:
// m, c ghost ngm / ngc cell arrays, d side arrays ghost 0, u / out ghost 0
void oracle_k_elliptic_apply(int ndim, const int* n, const double* dx, double* m, int ngm, double* c, int ngc,
                             double* const* d, const double* u, double* out)
{
   Box b;
   b.ndim = ndim;
   for (int a = 0; a < 3; a++) b.lo[a] = 0, b.hi[a] = a < ndim ? n[a] - 1 : 0;
   View dv[3];
   for (int a = 0; a < ndim; a++) dv[a] = make_view(d[a], b, a, 0, 1);
   elliptic_apply(b, dx, make_view(m, b, -1, ngm, 1), make_view(c, b, -1, ngc, 1), dv, u, out);
}
void oracle_k_quat_stencil_apply(int ndim, const int* n, const double* dx, double gamma, double* sqrt_m, int ngm,
                                 double* const* fc, const double* w, double* out)
{
   Box b;
   b.ndim = ndim;
   for (int a = 0; a < 3; a++) b.lo[a] = 0, b.hi[a] = a < ndim ? n[a] - 1 : 0;
   View fv[3];
   for (int a = 0; a < ndim; a++) fv[a] = make_view(fc[a], b, a, 0, 1);
   quat_stencil_apply(b, dx, gamma, make_view(sqrt_m, b, -1, ngm, 1), fv, w, out);
}
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
// piecewise entry points of quatrhs.m4 for the GPU tests of the ampe_k_* kernels
void oracle_k_anisotropic_gradient_flux(int ndim, const int* lo, const int* hi, const double* dx, double epsilon,
                                        double nu, int knumber, double* phase, int ngphase, double* quat, int ngq,
                                        int qlen, double* const* flux, int ngflux)
{
   Box b = mkbox(ndim, lo, hi);
   View f[3];
   for (int a = 0; a < ndim; a++) f[a] = make_view(flux[a], b, a, ngflux, 1);
   anisotropic_gradient_flux(b, dx, epsilon, nu, knumber, make_view(phase, b, -1, ngphase, 1),
                             make_view(quat, b, -1, ngq, qlen), qlen, f);
}
void oracle_k_computerhspbg(int ndim, const int* lo, const int* hi, const double* dx, double misorientation_factor,
                            double epsilonq, double* const* flux, int ngflux, double* temp, int ngtemp,
                            double phi_well_scale, double eta_well_scale, double* phi, int ngphi, double* eta,
                            int ngeta, double* ogm, int ngogm, double* rhs, int ngrhs, char phi_well_type,
                            char eta_well_type, char phi_interp_type, char oi1, char oi2, int with_orient,
                            int three_phase)
{
   Box b = mkbox(ndim, lo, hi);
   View f[3];
   for (int a = 0; a < ndim; a++) f[a] = make_view(flux[a], b, a, ngflux, 1);
   computerhspbg(b, dx, misorientation_factor, epsilonq, f, make_view(temp, b, -1, ngtemp, 1), phi_well_scale,
                 eta_well_scale, make_view(phi, b, -1, ngphi, 1), eta ? make_view(eta, b, -1, ngeta, 1) : View(),
                 ogm ? make_view(ogm, b, -1, ngogm, 1) : View(), make_view(rhs, b, -1, ngrhs, 1), phi_well_type,
                 eta_well_type, phi_interp_type, oi1, oi2, with_orient, three_phase);
}
}
