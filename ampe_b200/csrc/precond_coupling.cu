// The dquat/dphi coupling block of the preconditioner (precond_has_dquatdphi, QuatIntegrator.cc:114,
// 468-470; QuatFACOps::multiplyDQuatDPhiBlock, QuatFACOps.cc:1892-1956) as piecewise kernels on
// SAMRAI-layout device arrays, one per Fortran routine (include/ampe_b200_kernels.h):
//   ampe_k_quatdiffusionderiv           QUATDIFFUSIONDERIV          {2d,3d}/quatdiffusion.m4:11-150
//   ampe_k_quatmobilityderiv            QUATMOBILITYDERIV           {2d,3d}/mobility.m4:97-170
//   ampe_k_compute_dquatdphi_face_coef  COMPUTE_DQUATDPHI_FACE_COEF {2d,3d}/quatfacops.m4:124-170
//   ampe_k_multicomponent_multiply      MULTICOMPONENT_MULTIPLY     {2d,3d}/quatfacops.m4:1022-1049
//   ampe_k_take_square_root             TAKE_SQUARE_ROOT            {2d,3d}/quatfacops.m4:999-1020
//   ampe_k_cell_axpy                    HierarchyCellDataOpsReal::axpy (QuatIntegrator.cc:3612)
// One-off, bandwidth-trivial box kernels (one thread per loop point), same operation order as the
// Fortran; the flux and operator passes of the block re-use ampe_k_compute_flux / ampe_k_add_quat_op.
#include <cuda_runtime.h>

#include <string>

#include "../../include/ampe_b200_kernels.h"
#include "pointwise.cuh"

int ampe_set_err(int code, const std::string& msg);

namespace {
using namespace ampe;

#include "box_view.cuh"

// deriv_average_func, functions.f:371-402
__device__ __forceinline__ double deriv_average_func(double avg_phi, double next_phi, char type)
{
   if (type == 'a') return 0.5;
   if (avg_phi < 1.0e-16) return 0.0;
   return 0.5 * next_phi * next_phi / (avg_phi * avg_phi);
}

}  // namespace

extern "C" {

int ampe_k_quatdiffusionderiv(int ndim, const int* ifirst, const int* ilast, double misorientation_factor,
                              const double* temperature, int tghosts, const double* var, int ngvar, int depth,
                              double* const* gradq, int nggradq, double* const* diff, int ngdiff,
                              double gradient_floor, char smooth_floor_type, char interp_type, char avg_type,
                              void* stream)
{
   if (smooth_floor_type != 'm' && smooth_floor_type != 't' && smooth_floor_type != 's')
      return ampe_set_err(AMPE_EINVAL, "Error in eval_grad_normi: floor_type unknown");
   if (avg_type != 'a' && avg_type != 'h') return ampe_set_err(AMPE_EINVAL, "Error in average_func: type unknown");
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(var, b, -1, ngvar), T = view(temperature, b, -1, tghosts);
   const DV3 g = sides(gradq, b, nggradq), df = sides(diff, b, ngdiff);
   const double floor2 = gradient_floor * gradient_floor, maxn = 1.0 / gradient_floor;
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV ga = g.a[a], da = df.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         const double vm = ph(im, jm, km), vp = ph(i, j, k);
         if (vm < 1.0e-16 || vp < 1.0e-16) {
            da(i, j, k, 0) = 0.0;
            da(i, j, k, 1) = 0.0;
            return;
         }
         const double phi = average_func(vm, vp, avg_type);
         const double t = 0.5 * (T(im, jm, km) + T(i, j, k));
         const double d_deriv = misorientation_factor * t * deriv_interp_func(phi, interp_type);
         double g2 = 0.0;
         for (int n = 0; n < ndim; n++)
            for (int m = 0; m < depth; m++) {
               const double v = ga(i, j, k, n * depth + m);
               g2 = g2 + v * v;
            }
         const double fac = eval_grad_normi(g2, smooth_floor_type, floor2, maxn) * d_deriv;
         da(i, j, k, 0) = fac * deriv_average_func(phi, vm, avg_type);
         da(i, j, k, 1) = fac * deriv_average_func(phi, vp, avg_type);
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_quatmobilityderiv(int ndim, const int* ifirst, const int* ilast, const double* phase, int ngphase,
                             double* dmobility, int ngmobility, double scale_mobility, double min_mobility,
                             char func_type, double alt_scale_factor, void* stream)
{
   const char f = func_type;
   if (f != 'p' && f != 'P' && f != 'e' && f != 'E' && f != 'i' && f != 'I')
      return ampe_set_err(AMPE_EINVAL, "Error in quatmobilityderiv: unknown function type");
   const Box b = mkbox(ndim, ifirst, ilast);
   const CV ph = view(phase, b, -1, ngphase);
   const DV dm = view(dmobility, b, -1, ngmobility);
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      double phi = ph(i, j, k), dqfunc;
      if (f == 'p' || f == 'P') {
         phi = fmax(0.0, fmin(1.0, phi));
         dqfunc = -30.0 * phi * phi * (1.0 - phi) * (1.0 - phi);
      } else if (f == 'e' || f == 'E') {
         const double c = alt_scale_factor;
         phi = fmax(0.0, fmin(1.0, phi));
         dqfunc = (c * exp(c * phi)) / (1. - exp(c));
      } else {
         phi = fmax(1.e-6, fmin(1.0, phi));
         dqfunc = (phi - 2.0) / (phi * phi * phi);
      }
      dm(i, j, k) = (scale_mobility - min_mobility) * dqfunc;
   });
}

int ampe_k_compute_dquatdphi_face_coef(int ndim, const int* lo, const int* hi, int depth, double* const* dprime,
                                       int ngdprime, const double* phi, int ngphi, double* const* face_coef,
                                       int ngfc, void* stream)
{
   (void)depth;  // the reference passes qlen; the routine does not use it
   const Box b = mkbox(ndim, lo, hi);
   const CV ph = view(phi, b, -1, ngphi);
   const DV3 dp = sides(dprime, b, ngdprime), fc = sides(face_coef, b, ngfc);
   for (int a = 0; a < ndim; a++) {
      int L[3], H[3];
      side_bounds(b, a, 0, L, H);
      const DV da = dp.a[a], fa = fc.a[a];
      int rc = for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
         const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
         fa(i, j, k) = -da(i, j, k, 0) * ph(im, jm, km) - da(i, j, k, 1) * ph(i, j, k);
      });
      if (rc) return rc;
   }
   return AMPE_OK;
}

int ampe_k_multicomponent_multiply(int ndim, const int* lo, const int* hi, const double* factor, int ngfactor,
                                   double* var, int ngvar, int vnc, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const CV f = view(factor, b, -1, ngfactor);
   const DV v = view(var, b, -1, ngvar);
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      for (int n = 0; n < vnc; n++) v(i, j, k, n) = v(i, j, k, n) * f(i, j, k);
   });
}

// the reference takes the root over the whole ghost box of the array
int ampe_k_take_square_root(int ndim, const int* lo, const int* hi, double* data, int ng, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const DV d = view(data, b, -1, ng);
   int L[3], H[3];
   cell_bounds(b, ng, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) { d(i, j, k) = sqrt(d(i, j, k)); });
}

// dst = alpha x + y on the box, per depth component
int ampe_k_cell_axpy(int ndim, const int* lo, const int* hi, int depth, double alpha, const double* x, int ngx,
                     const double* y, int ngy, double* dst, int ngdst, void* stream)
{
   const Box b = mkbox(ndim, lo, hi);
   const CV xv = view(x, b, -1, ngx), yv = view(y, b, -1, ngy);
   const DV dv = view(dst, b, -1, ngdst);
   int L[3], H[3];
   cell_bounds(b, 0, L, H);
   return for_box(L, H, ST(stream), [=] __device__(int i, int j, int k) {
      for (int m = 0; m < depth; m++) dv(i, j, k, m) = alpha * xv(i, j, k, m) + yv(i, j, k, m);
   });
}

}  // extern "C"
