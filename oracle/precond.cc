// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Block preconditioners of QuatIntegrator::CVSpgmrPrecondSet / CVSpgmrPrecondSolve
// (source/QuatIntegrator.cc:3300-3376, 3666-3771) on a single uniform periodic level.  Three parts:
//
//  (1) the OPERATORS, restated routine by routine from the reference with their own array layouts
//      (ghosted solution, side-centred fluxes): efo_compfluxvardc + efo_compresvarsca
//      (2d/ellipticfacops.m4:16-56, 346-393), set_j_ij + set_stencil (2d/quatlevelsolver.m4:9-118),
//      PhaseFACOps::setCOnPatchPrivate (PhaseFACOps.cc:100-186), ConcFACOps::setOperatorCoefficients
//      (ConcFACOps.cc:19-50), EBSCompositionRHSStrategy::setDiffusionCoeffForPreconditioner
//      (EBSCompositionRHSStrategy.cc:332-430).  These are the independent check of the product's
//      operator (ampe_mg_apply) and measure the residual its solve leaves.
//  (2) a host loop over the PRODUCT's per-cell multigrid arithmetic (ampe_b200/csrc/mg_cell.h, the
//      functions the CUDA kernels call) with the same cycle structure as mg.cu -- like stepper.cc
//      re-uses the product's integrator template.  The reference solves the level with hypre PFMG
//      (third party, absent from its tree): there is no reference solver to restate, so the solve is
//      judged by the residual measured with (1) and by what it does to the Krylov iteration.
//      PARITY UNPINNED for the solver itself.
//  (3) CVSpgmrPrecondSet / CVSpgmrPrecondSolve on the oracle context, used by stepper.cc.
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "../ampe_b200/csrc/mg_cell.h"
#include "ctx.h"
#include "oracle.h"
#include "precond.h"

namespace oracle {

static inline int wrapi(int i, int n)
{
   i %= n;
   return i < 0 ? i + n : i;
}

// ---- (1) operators -----------------------------------------------------------------------------
// A u = M div(D grad u) + C u on ghost-0 input, periodic: ghost fill, efo_compfluxvardc, then the
// operator part of efo_compresvarsca (residual = rhs - [this]).
void elliptic_apply(const Box& b, const double* dx, View m, View c, View* d, const double* u, double* out)
{
   const int D = b.ndim;
   const int n0 = b.hi[0] + 1, n1 = b.hi[1] + 1, n2 = b.hi[2] + 1;
   Field soln;
   soln.alloc(b, -1, 1, 1);
   const int g2 = D == 3 ? 1 : 0;
   for (int k = -g2; k < n2 + g2; k++)
      for (int j = -1; j < n1 + 1; j++)
         for (int i = -1; i < n0 + 1; i++)
            soln.v(i, j, k) = u[(size_t)wrapi(i, n0) + (size_t)n0 * (wrapi(j, n1) + (size_t)n1 * wrapi(k, n2))];
   SideField flux;
   flux.alloc(b, 0, 1);
   const double dxi = 1. / dx[0], dyi = 1. / dx[1], dzi = D == 3 ? 1. / dx[2] : 0.0;
   for (int k = 0; k < n2; k++)
      for (int j = 0; j < n1; j++)
         for (int i = 0; i <= n0; i++)
            flux.a[0].v(i, j, k) = dxi * d[0](i, j, k) * (soln.v(i, j, k) - soln.v(i - 1, j, k));
   for (int k = 0; k < n2; k++)
      for (int j = 0; j <= n1; j++)
         for (int i = 0; i < n0; i++)
            flux.a[1].v(i, j, k) = dyi * d[1](i, j, k) * (soln.v(i, j, k) - soln.v(i, j - 1, k));
   if (D == 3)
      for (int k = 0; k <= n2; k++)
         for (int j = 0; j < n1; j++)
            for (int i = 0; i < n0; i++)
               flux.a[2].v(i, j, k) = dzi * d[2](i, j, k) * (soln.v(i, j, k) - soln.v(i, j, k - 1));
   for (int k = 0; k < n2; k++)
      for (int j = 0; j < n1; j++)
         for (int i = 0; i < n0; i++) {
            double div = dxi * (flux.a[0].v(i + 1, j, k) - flux.a[0].v(i, j, k)) +
                         dyi * (flux.a[1].v(i, j + 1, k) - flux.a[1].v(i, j, k));
            if (D == 3) div = div + dzi * (flux.a[2].v(i, j, k + 1) - flux.a[2].v(i, j, k));
            out[(size_t)i + (size_t)n0 * (j + (size_t)n1 * k)] = m(i, j, k) * div + c(i, j, k) * soln.v(i, j, k);
         }
}

// the quaternion level matrix applied to one component: set_j_ij (off-diagonals fc / h^2, diagonal
// minus their sum), set_stencil (centre = gamma sqrt_m diag sqrt_m + 1, neighbour = gamma sqrt_m(i)
// offdiag sqrt_m(neighbour)), then the stencil product hypre would form
void quat_stencil_apply(const Box& b, const double* h, double gamma, View sqrt_m, View* fc, const double* w,
                        double* out)
{
   const int D = b.ndim;
   const int n0 = b.hi[0] + 1, n1 = b.hi[1] + 1, n2 = b.hi[2] + 1;
   SideField od;
   od.alloc(b, 0, 1);
   Field diag;
   diag.alloc(b, -1, 0, 1);
   for (int a = 0; a < D; a++) {
      const double fac = 1.0 / (h[a] * h[a]);
      for (int k = 0; k < n2 + (a == 2); k++)
         for (int j = 0; j < n1 + (a == 1); j++)
            for (int i = 0; i < n0 + (a == 0); i++) od.a[a].v(i, j, k) = fc[a](i, j, k) * fac;
   }
   for (int k = 0; k < n2; k++)
      for (int j = 0; j < n1; j++)
         for (int i = 0; i < n0; i++) {
            double dg = -od.a[0].v(i, j, k) - od.a[0].v(i + 1, j, k) - od.a[1].v(i, j, k) - od.a[1].v(i, j + 1, k);
            if (D == 3) dg = dg - od.a[2].v(i, j, k) - od.a[2].v(i, j, k + 1);
            diag.v(i, j, k) = dg;
         }
   auto W = [&](int i, int j, int k) {
      return w[(size_t)wrapi(i, n0) + (size_t)n0 * (wrapi(j, n1) + (size_t)n1 * wrapi(k, n2))];
   };
   for (int k = 0; k < n2; k++)
      for (int j = 0; j < n1; j++)
         for (int i = 0; i < n0; i++) {
            const double fac = gamma * sqrt_m(i, j, k);
            double acc = (fac * diag.v(i, j, k) * sqrt_m(i, j, k) + 1.0) * W(i, j, k);
            acc += fac * od.a[0].v(i, j, k) * sqrt_m(i - 1, j, k) * W(i - 1, j, k);
            acc += fac * od.a[0].v(i + 1, j, k) * sqrt_m(i + 1, j, k) * W(i + 1, j, k);
            acc += fac * od.a[1].v(i, j, k) * sqrt_m(i, j - 1, k) * W(i, j - 1, k);
            acc += fac * od.a[1].v(i, j + 1, k) * sqrt_m(i, j + 1, k) * W(i, j + 1, k);
            if (D == 3) {
               acc += fac * od.a[2].v(i, j, k) * sqrt_m(i, j, k - 1) * W(i, j, k - 1);
               acc += fac * od.a[2].v(i, j, k + 1) * sqrt_m(i, j, k + 1) * W(i, j, k + 1);
            }
            out[(size_t)i + (size_t)n0 * (j + (size_t)n1 * k)] = acc;
         }
}

// ---- (1b) the dquat/dphi coupling block (precond_has_dquatdphi, QuatIntegrator.cc:114, 468-470) ---
// quatdiffusionderiv ({2d,3d}/quatdiffusion.m4:11-150): derivative of the orientation diffusivity
// 2HT p'(phi_face) / |grad q| with respect to the two cells of a face; depth 1 = lower cell, 2 = upper
void quatdiffusionderiv(const Box& b, double misorientation_factor, View temperature, View var, int depth,
                        View* gradq, View* diff, double gradient_floor, char smooth_floor_type, char interp_type,
                        char avg_type)
{
   const double jlf_threshold = 1.0e-16;
   const double floor_grad_norm2 = gradient_floor * gradient_floor;
   const double max_grad_normi = 1.0 / gradient_floor;
   for (int a = 0; a < b.ndim; a++) {
      const int e0 = a == 0, e1 = a == 1, e2 = a == 2;
      for (int k = b.lo[2]; k <= b.hi[2] + e2; k++)
         for (int j = b.lo[1]; j <= b.hi[1] + e1; j++)
            for (int i = b.lo[0]; i <= b.hi[0] + e0; i++) {
               const double vm = var(i - e0, j - e1, k - e2), vp = var(i, j, k);
               if (vm < jlf_threshold || vp < jlf_threshold) {
                  diff[a](i, j, k, 0) = 0.0;
                  diff[a](i, j, k, 1) = 0.0;
                  continue;
               }
               const double phi = average_func(vm, vp, avg_type);
               const double t = 0.5 * (temperature(i - e0, j - e1, k - e2) + temperature(i, j, k));
               const double d_deriv = misorientation_factor * t * deriv_interp_func(phi, interp_type);
               double grad_norm2 = 0.0;
               for (int n = 0; n < b.ndim; n++)
                  for (int m = 0; m < depth; m++) {
                     const double g = gradq[a](i, j, k, n * depth + m);
                     grad_norm2 = grad_norm2 + g * g;
                  }
               const double grad_normi = eval_grad_normi(grad_norm2, smooth_floor_type, floor_grad_norm2, max_grad_normi);
               const double fac = grad_normi * d_deriv;
               diff[a](i, j, k, 0) = fac * deriv_average_func(phi, vm, avg_type);
               diff[a](i, j, k, 1) = fac * deriv_average_func(phi, vp, avg_type);
            }
   }
}

// quatmobilityderiv ({2d,3d}/mobility.m4:97-170)
void quatmobilityderiv(const Box& b, View phase, View dmobility, double scale_mobility, double min_mobility,
                       char func_type, double alt_scale_factor)
{
   for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
         for (int i = b.lo[0]; i <= b.hi[0]; i++) {
            double phi = phase(i, j, k), dqfunc;
            if (func_type == 'p' || func_type == 'P') {
               phi = fmax(0.0, fmin(1.0, phi));
               dqfunc = -30.0 * phi * phi * (1.0 - phi) * (1.0 - phi);
            } else if (func_type == 'e' || func_type == 'E') {
               const double c = alt_scale_factor;
               phi = fmax(0.0, fmin(1.0, phi));
               dqfunc = (c * exp(c * phi)) / (1. - exp(c));
            } else if (func_type == 'i' || func_type == 'I') {
               phi = fmax(1.e-6, fmin(1.0, phi));
               dqfunc = (phi - 2.0) / (phi * phi * phi);
            } else {
               throw std::runtime_error("Error in quatmobilityderiv: unknown function type");
            }
            dmobility(i, j, k) = (scale_mobility - min_mobility) * dqfunc;
         }
}

// compute_dquatdphi_face_coef ({2d,3d}/quatfacops.m4:124-170): fc = -D'_lower phi_lower - D'_upper phi_upper
void compute_dquatdphi_face_coef(const Box& b, View* dprime, View phi, View* fc)
{
   for (int a = 0; a < b.ndim; a++) {
      const int e0 = a == 0, e1 = a == 1, e2 = a == 2;
      for (int k = b.lo[2]; k <= b.hi[2] + e2; k++)
         for (int j = b.lo[1]; j <= b.hi[1] + e1; j++)
            for (int i = b.lo[0]; i <= b.hi[0] + e0; i++)
               fc[a](i, j, k) = -dprime[a](i, j, k, 0) * phi(i - e0, j - e1, k - e2) - dprime[a](i, j, k, 1) * phi(i, j, k);
   }
}

// multicomponent_multiply ({2d,3d}/quatfacops.m4:1022-1049)
void multicomponent_multiply(const Box& b, View factor, View var, int vnc)
{
   for (int n = 0; n < vnc; n++)
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
         for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) var(i, j, k, n) = var(i, j, k, n) * factor(i, j, k);
}

// ---- (2) host loop over the product's per-cell multigrid arithmetic ------------------------------
using ampe_mg_cell::Level;

HostMG::HostMG(int ndim, const int* n, const double* dx, bool with_s, int ncomp)
    : d_ndim(ndim), d_nc(ncomp < 1 ? 1 : ncomp), d_with_s(with_s)
{
   int cur[3] = {1, 1, 1};
   for (int d = 0; d < 3; d++) {
      d_n[d] = d < ndim ? n[d] : 1;
      cur[d] = d_n[d];
      d_inv_h2[d] = d < ndim ? 1.0 / (dx[d] * dx[d]) : 0.0;
   }
   for (int l = 0; l < 16; l++) {
      const size_t nc = (size_t)cur[0] * cur[1] * cur[2];
      d_store.emplace_back(nc * ((with_s ? 1 : 0) + 3 * d_nc), 0.0);  // s, then u, f, r with d_nc components each
      d_coef.emplace_back();                                  // c, m, d0..d2: allocated when stored as arrays
      Level L;
      L.ndim = ndim;
      for (int d = 0; d < 3; d++) L.n[d] = cur[d];
      d_levels.push_back(L);
      bool even = true;
      for (int d = 0; d < ndim; d++) even = even && (cur[d] % 2 == 0);
      d_two_colour.push_back(even);
      bool can = even;
      for (int d = 0; d < ndim; d++) can = can && (cur[d] / 2 >= 2);
      if (!can) break;
      for (int d = 0; d < ndim; d++) cur[d] /= 2;
   }
   for (size_t l = 0; l < d_levels.size(); l++) {  // pointers after the vector of vectors stopped growing
      Level& L = d_levels[l];
      const size_t nc = (size_t)L.n[0] * L.n[1] * L.n[2];
      double* p = d_store[l].data();
      L.c = L.m = L.s = nullptr;
      L.c_const = L.m_const = 1.0;
      if (with_s) L.s = p, p += nc;
      for (int d = 0; d < 3; d++) L.d[d] = nullptr, L.d_const[d] = 0.0;
      L.nc = d_nc, L.cs = (long long)nc;
      L.clamp[0] = L.clamp[1] = L.clamp[2] = 0;
      L.u = p, p += nc * d_nc;
      L.f = p, p += nc * d_nc;
      L.r = p, p += nc * d_nc;
   }
}

// same rule as configure() in ampe_b200/csrc/mg.cu: a coefficient that is a constant of the block is not
// stored; C and M constants are level independent, a constant D / h^2 is divided by 4 per level
void HostMG::configure(bool var_c, double c_const, bool var_m, double m_const, bool var_d, double d_const)
{
   double scale = 1.0;
   for (size_t l = 0; l < d_levels.size(); l++) {
      Level& L = d_levels[l];
      const size_t nc = (size_t)L.n[0] * L.n[1] * L.n[2];
      auto& cf = d_coef[l];
      auto arr = [&](int which) {
         if (cf[which].size() != nc) cf[which].assign(nc, 0.0);
         return cf[which].data();
      };
      L.c = var_c ? arr(0) : nullptr;
      L.m = var_m ? arr(1) : nullptr;
      L.c_const = c_const, L.m_const = m_const;
      for (int a = 0; a < 3; a++) {
         L.d[a] = (var_d && a < d_ndim) ? arr(2 + a) : nullptr;
         L.d_const[a] = a < d_ndim ? d_const * d_inv_h2[a] * scale : 0.0;
      }
      scale *= 0.25;
   }
}

#define MG_FOR_CELLS(L)                   \
   for (int k = 0; k < (L).n[2]; k++)     \
      for (int j = 0; j < (L).n[1]; j++)  \
         for (int i = 0; i < (L).n[0]; i++)

void HostMG::buildCoarse()
{
   for (size_t l = 0; l + 1 < d_levels.size(); l++) {
      const Level &F = d_levels[l], &Cl = d_levels[l + 1];
      MG_FOR_CELLS(Cl) ampe_mg_cell::mg_coarsen_cell(F, Cl, i, j, k);
   }
   d_set = true;
}

void HostMG::setElliptic(const double* m, int ngm, double m_const, const double* c, int ngc, double c_const,
                         const double* const* d, const double* const* d2, int ngd, double d_scale, double d_const)
{
   if (d_with_s) throw std::runtime_error("HostMG::setElliptic on a quaternion solver");
   const bool bc = d_zero_slope[0] || d_zero_slope[1] || d_zero_slope[2];
   configure(c != nullptr, c_const, m != nullptr, m_const, d != nullptr || bc, d_const);
   const Level& L = d_levels[0];
   if (m || c || d) {
      MG_FOR_CELLS(L) ampe_mg_cell::mg_set_elliptic_cell(L, m, ngm, c, ngc, d, d2, ngd, d_scale, d_inv_h2, i, j, k);
   }
   if (bc) {
      MG_FOR_CELLS(L) ampe_mg_cell::mg_boundary_faces_cell(L, d == nullptr ? 1 : 0, i, j, k);
   }
   buildCoarse();
}

void HostMG::setZeroSlope(const int* zero_slope)
{
   for (int d = 0; d < 3; d++) d_zero_slope[d] = (d < d_ndim && zero_slope[d]) ? 1 : 0;
   for (Level& L : d_levels)
      for (int d = 0; d < 3; d++) L.clamp[d] = d_zero_slope[d];
}

void HostMG::setQuat(double gamma, const double* mobility, int ngm, const double* const* face_coef, int ngfc)
{
   if (!d_with_s) throw std::runtime_error("HostMG::setQuat on a scalar solver");
   configure(false, 1.0, true, 0.0, true, 0.0);
   const Level& L = d_levels[0];
   MG_FOR_CELLS(L) ampe_mg_cell::mg_set_quat_cell(L, gamma, mobility, ngm, face_coef, ngfc, d_inv_h2, i, j, k);
   if (d_zero_slope[0] || d_zero_slope[1] || d_zero_slope[2]) {
      MG_FOR_CELLS(L) ampe_mg_cell::mg_boundary_faces_cell(L, 0, i, j, k);
   }
   buildCoarse();
}

void HostMG::setFused(bool on, long long min_cells)
{
   d_fused_restriction = on;
   d_tile.assign(d_levels.size(), ampe_mg_cell::TileShape{{0, 0, 0}});
   d_alt_u.resize(d_levels.size());
   for (size_t l = 0; on && l < d_levels.size(); l++) {
      const Level& L = d_levels[l];
      const long long nc = (long long)L.n[0] * L.n[1] * L.n[2];
      if (!d_two_colour[l] || nc <= min_cells) continue;
      auto pick = [](int n, int first) {
         for (int t = first; t >= 2; t /= 2)
            if (n % t == 0) return t;
         return 0;
      };
      ampe_mg_cell::TileShape T;
      T.t[0] = pick(L.n[0], d_ndim == 3 ? 32 : 64);
      T.t[1] = pick(L.n[1], d_ndim == 3 ? 8 : 16);
      T.t[2] = d_ndim == 3 ? pick(L.n[2], 8) : 1;
      if (T.t[0] < 8 || T.t[1] < 2 || T.t[2] < 1) continue;
      d_alt_u[l].assign((size_t)nc * d_nc, 0.0);
      d_tile[l] = T;
   }
}

int HostMG::fusedLevels() const
{
   int n = 0;
   for (const auto& T : d_tile) n += T.t[0] > 0;
   return n;
}

void HostMG::smooth(int l, int sweeps)
{
   Level& L = d_levels[l];
   for (int s = 0; s < sweeps; s++) {
      if (!d_tile.empty() && d_tile[l].t[0] > 0) {
         // the device runs one block per tile; here one "thread" per tile walks the same phases
         const ampe_mg_cell::TileShape T = d_tile[l];
         std::vector<double> tile((size_t)(T.t[0] + 4) * (T.t[1] + 4) * (L.ndim == 3 ? T.t[2] + 4 : 1));
         double* alt = d_alt_u[l].data();
         for (int o2 = 0; o2 < L.n[2]; o2 += T.t[2])
            for (int o1 = 0; o1 < L.n[1]; o1 += T.t[1])
               for (int o0 = 0; o0 < L.n[0]; o0 += T.t[0])
                  for (int m = 0; m < L.nc; m++)  // the device loops the components inside the block too
                     ampe_mg_cell::mg_rb_tile_pass(L, L.f + m * L.cs, L.u + m * L.cs, alt + m * L.cs, tile.data(), T, o0,
                                                   o1, o2, 0, 1);
         // ping-pong: the level's u lives in d_store; copy back instead of swapping owners
         std::memcpy(L.u, alt, sizeof(double) * d_alt_u[l].size());
         continue;
      }
      if (d_two_colour[l]) {
         for (int colour = 0; colour < 2; colour++)
            MG_FOR_CELLS(L)
         if (((i + j + k) & 1) == colour) ampe_mg_cell::mg_smooth_cell(L, i, j, k);
      } else {
         MG_FOR_CELLS(L) ampe_mg_cell::mg_residual_cell(L, i, j, k);
         MG_FOR_CELLS(L) ampe_mg_cell::mg_jacobi_cell(L, 0.8, i, j, k);
      }
   }
}

void HostMG::vcycle()
{
   const int nl = (int)d_levels.size();
   for (int l = 0; l + 1 < nl; l++) {
      smooth(l, d_pre);
      const Level &F = d_levels[l], &Cl = d_levels[l + 1];
      if (d_fused_restriction) {
         MG_FOR_CELLS(Cl) ampe_mg_cell::mg_restrict_residual_cell(F, Cl, i, j, k);
      } else {
         MG_FOR_CELLS(F) ampe_mg_cell::mg_residual_cell(F, i, j, k);
         MG_FOR_CELLS(Cl) ampe_mg_cell::mg_restrict_cell(F, Cl, i, j, k);
      }
   }
   smooth(nl - 1, d_coarse);
   for (int l = nl - 2; l >= 0; l--) {
      const Level &F = d_levels[l], &Cl = d_levels[l + 1];
      MG_FOR_CELLS(F) ampe_mg_cell::mg_prolong_cell(Cl, F, i, j, k);
      smooth(l, d_post);
   }
}

void HostMG::solve(const double* rhs, double* soln, int ncycles, bool symmetrized)
{
   if (!d_set) throw std::runtime_error("HostMG::solve: coefficients not set");
   const Level& L = d_levels[0];
   const size_t nc = (size_t)L.n[0] * L.n[1] * L.n[2];
   for (int m = 0; m < L.nc; m++)
      for (size_t o = 0; o < nc; o++) {
         L.f[o + m * nc] = (symmetrized && L.s) ? rhs[o + m * nc] / L.s[o] : rhs[o + m * nc];
         L.u[o + m * nc] = 0.0;
      }
   for (int c = 0; c < ncycles; c++) vcycle();
   for (int m = 0; m < L.nc; m++)
      for (size_t o = 0; o < nc; o++)
         soln[o + m * nc] = (symmetrized && L.s) ? L.u[o + m * nc] * L.s[o] : L.u[o + m * nc];
}

void HostMG::apply(const double* u, double* out) const
{
   const Level& L = d_levels[0];
   MG_FOR_CELLS(L) out[ampe_mg_cell::mg_index(L, i, j, k)] = ampe_mg_cell::mg_apply_cell(L, u, i, j, k);
}

const double* HostMG::levelArray(int level, int which) const
{
   const Level& L = d_levels.at(level);
   const double* p = which == 0 ? L.c : which == 1 ? L.m : which == 2 ? L.s : L.d[which - 3];
   if (p || which == 2 || (which >= 3 && which - 3 >= d_ndim)) return p;
   // a constant of the block: materialise it for the caller
   const size_t nc = (size_t)L.n[0] * L.n[1] * L.n[2];
   d_scratch.assign(nc, which == 0 ? L.c_const : which == 1 ? L.m_const : L.d_const[which - 3]);
   return d_scratch.data();
}

// ---- (3) CVSpgmrPrecondSet / CVSpgmrPrecondSolve on the oracle context ---------------------------
struct Precond {
   double gamma = 0.0;
   int ncycles = 2;
   std::unique_ptr<HostMG> phase, quat, conc, temp;
   // the operators' SAMRAI-layout coefficient arrays, kept for precond_apply
   Field phase_c;       // C of the phase block, ghost 0
   SideField conc_d;    // -gamma D_pfm of the composition block, ghost 0
   Field sqrt_m;        // sqrt of the quaternion mobility, ghost 1
   Field ones, conc_m;  // constants as fields for elliptic_apply
   SideField const_d;   // constant D as side field (phase / temperature)
   // dquat/dphi coupling block (precond_has_dquatdphi)
   bool has_dquatdphi = false;
   Field m_deriv;        // d(quat mobility)/d(phi), ghost 0
   SideField d_deriv;    // d(face diffusivity)/d(phi of the lower, upper cell), depth 2, ghost 0
   SideField fc_scratch, flux;  // face_coef_scratch, flux scratch (depth Q)
   Field phase_sol, quat_rhs;   // z_phase with ghosts (ghost 1); coupled right-hand side (depth Q)
};

static void views3(SideField& s, View* v, int ndim)
{
   for (int d = 0; d < ndim; d++) v[d] = s.a[d].v;
}

void precond_destroy(Ctx* c)
{
   delete (Precond*)c->precond;
   c->precond = nullptr;
}

// Requires the context's intermediates at the state y of the last fd_flag = 0 evaluation (the
// reference calls setCoefficients(t, y, true) here, QuatIntegrator.cc:3319; the integrator template
// calls this hook right after that evaluation).
int precond_setup(Ctx* c, double gamma, int ncycles, bool has_dquatdphi)
{
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   const int D = p.ndim;
   if (!c->precond) c->precond = new Precond;
   Precond& P = *(Precond*)c->precond;
   P.gamma = gamma;
   P.ncycles = ncycles;
   const int n[3] = {b.hi[0] + 1, b.hi[1] + 1, b.hi[2] + 1};
   if (p.with_phase) {
      // PhaseFACOps::setOperatorCoefficients (PhaseFACOps.cc:33-51): M = mobility field,
      // C = 1 + gamma M w g''(phi) (setCOnPatchPrivate), D = -gamma eps^2
      if (P.phase_c.data.empty()) P.phase_c.alloc(b, -1, 0, 1);
      for (int k = 0; k < n[2]; k++)
         for (int j = 0; j < n[1]; j++)
            for (int i = 0; i < n[0]; i++) {
               const double m = c->phase_mobility.v(i, j, k);
               const double phi = c->phase.v(i, j, k);
               const double g_phi_dbl_prime = second_deriv_well_func(phi, 'd');
               const double gamma_m = gamma * m;
               P.phase_c.v(i, j, k) = 1.0 + gamma_m * p.phi_well_scale * g_phi_dbl_prime;
            }
      if (!P.phase) {
         P.phase.reset(new HostMG(D, n, p.dx, false));
         P.phase->setZeroSlope(p.zero_slope);
      }
      P.phase->setElliptic(c->phase_mobility.data.data(), 1, 0.0, P.phase_c.data.data(), 0, 0.0, nullptr, nullptr, 0,
                           1.0, -gamma * p.epsilon_phase * p.epsilon_phase);
   }
   if (p.with_concentration && (p.conc_rhs_form == AMPE_CONC_KKS || p.conc_rhs_form == AMPE_CONC_EBS)) {
      // ConcFACOps::setOperatorCoefficients (ConcFACOps.cc:19-50): D = -gamma D_pfm, C = 1,
      // M = mobility; D_pfm = D_l + D_a for EBS (setDiffusionCoeffForPreconditioner), D0 for KKS
      if (P.conc_d.a[0].data.empty()) P.conc_d.alloc(b, 0, 1);
      const double* d1[3] = {nullptr, nullptr, nullptr};
      const double* d2[3] = {nullptr, nullptr, nullptr};
      const bool ebs = p.conc_rhs_form == AMPE_CONC_EBS;
      for (int a = 0; a < D; a++) {
         const std::vector<double>& x = ebs ? c->diff_l.a[a].data : c->diff0.a[a].data;
         d1[a] = x.data();
         if (ebs) d2[a] = c->diff_a.a[a].data.data();
         for (size_t o = 0; o < x.size(); o++) {
            double v = 0.0;
            v = v + x[o];
            if (ebs) v = v + c->diff_a.a[a].data[o];
            P.conc_d.a[a].data[o] = -gamma * v;
         }
      }
      if (!P.conc) {
         P.conc.reset(new HostMG(D, n, p.dx, false));
         P.conc->setZeroSlope(p.zero_slope);
      }
      P.conc->setElliptic(nullptr, 0, p.conc_mobility, nullptr, 0, 1.0, d1, ebs ? d2 : nullptr, 0, -gamma, 0.0);
   }
   if (p.with_unsteady_temperature) {
      // QuatIntegrator.cc:3340-3346: m = 1, c = 1, d = -gamma thermal_diffusivity
      if (!P.temp) {
         P.temp.reset(new HostMG(D, n, p.dx, false));
         P.temp->setZeroSlope(p.zero_slope);
      }
      P.temp->setElliptic(nullptr, 0, 1.0, nullptr, 0, 1.0, nullptr, nullptr, 0, 1.0, -gamma * p.thermal_diffusivity);
   }
   if (p.evolve_quat) {
      // QuatFACOps::setOperatorCoefficients (QuatFACOps.cc:735-818): face coefficients from
      // (phase, T, grad_q copy) -- the context's face_coef of the last evaluation --, sqrt(mobility)
      if (P.sqrt_m.data.empty()) P.sqrt_m.alloc(b, -1, 1, 1);
      for (size_t o = 0; o < c->quat_mobility.data.size(); o++) P.sqrt_m.data[o] = sqrt(c->quat_mobility.data[o]);
      const double* fc[3] = {nullptr, nullptr, nullptr};
      for (int a = 0; a < D; a++) fc[a] = c->face_coef.a[a].data.data();
      if (!P.quat) {
         P.quat.reset(new HostMG(D, n, p.dx, true, p.qlen));
         P.quat->setZeroSlope(p.zero_slope);
      }
      P.quat->setQuat(gamma, c->quat_mobility.data.data(), 1, fc, 0);
   }
   // setCoefficients with d_precond_has_dquatdphi (QuatIntegrator.cc:2978-2983, 3064-3070)
   P.has_dquatdphi = has_dquatdphi && p.with_phase && p.evolve_quat;
   if (P.has_dquatdphi) {
      const int Q = p.qlen;
      if (P.m_deriv.data.empty()) {
         P.m_deriv.alloc(b, -1, 0, 1);
         P.d_deriv.alloc(b, 0, 2);
         P.fc_scratch.alloc(b, 0, 1);
         P.flux.alloc(b, 0, Q);
         P.phase_sol.alloc(b, -1, 1, 1);
         P.quat_rhs.alloc(b, -1, 0, Q);
      }
      quatmobilityderiv(b, c->phase.v, P.m_deriv.v, p.quat_mobility, p.min_quat_mobility, p.quat_mobility_func,
                        p.quat_mobility_alt_scale);
      View gq[3], dd[3];
      views3(c->quat_grad_side_copy, gq, D);
      views3(P.d_deriv, dd, D);
      // DerivDiffusionCoeffForQuat (QuatIntegrator.cc:1954-1957): interp = orient_interp_func_type1
      quatdiffusionderiv(b, 2. * p.H_parameter, c->temp.v, c->phase.v, Q, gq, dd, p.quat_grad_floor, p.grad_floor_type,
                         p.orient_interp1, p.avg_func);
   }
   return 0;
}

// QuatFACOps::multiplyDQuatDPhiBlock(phase_id, out_id) (QuatFACOps.cc:1892-1956): out = dF_q/dphi z_phase with
// the frozen q: [mobility'(phi) div(fc grad q)] z_phase + sqrt_m div(fc' grad q), fc' = -D'_- z_- - D'_+ z_+
static void multiply_dquatdphi_block(Ctx* c, Precond& P, const double* z_phase, View out)
{
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   const int D = p.ndim, Q = p.qlen;
   const int n0 = b.hi[0] + 1, n1 = b.hi[1] + 1, n2 = b.hi[2] + 1;
   const int g2 = D == 3 ? 1 : 0;
   for (int k = -g2; k < n2 + g2; k++)
      for (int j = -1; j < n1 + 1; j++)
         for (int i = -1; i < n0 + 1; i++)
            P.phase_sol.v(i, j, k) = z_phase[(size_t)wrapi(i, n0) + (size_t)n0 * (wrapi(j, n1) + (size_t)n1 * wrapi(k, n2))];
   for (int m = 0; m < Q; m++)
      for (int k = 0; k < n2; k++)
         for (int j = 0; j < n1; j++)
            for (int i = 0; i < n0; i++) out(i, j, k, m) = 0.0;
   View fc[3], fcs[3], fl[3], dd[3];
   views3(c->face_coef, fc, D);
   views3(P.fc_scratch, fcs, D);
   views3(P.flux, fl, D);
   views3(P.d_deriv, dd, D);
   // accumulateOperatorOnLevel(d_m_deriv_id, d_face_coef_id, d_q_local_id, -1, out_id, ...)
   compute_flux(b, Q, fc, c->quat.v, p.dx, fl);
   add_quat_op(b, Q, P.m_deriv.v, fl, p.dx, out);
   multicomponent_multiply(b, P.phase_sol.v, out, Q);
   // computeDQuatDPhiFaceCoefs + accumulateOperatorOnLevel(d_sqrt_m_id, d_face_coef_scratch_id, ...)
   compute_dquatdphi_face_coef(b, dd, P.phase_sol.v, fcs);
   compute_flux(b, Q, fcs, c->quat.v, p.dx, fl);
   add_quat_op(b, Q, P.sqrt_m.v, fl, p.dx, out);
}

static size_t ncell_of(const Ctx* c)
{
   size_t n = 1;
   for (int d = 0; d < c->cfg.ndim; d++) n *= (size_t)c->cfg.n[d];
   return n;
}

// CVSpgmrPrecondSolve (QuatIntegrator.cc:3666-3771), block diagonal (precond_has_dquatdphi = false,
// QuatIntegrator.cc:468-470): z_block = A_block^-1 r_block; a block without a solver is copied
int precond_solve(Ctx* c, const ampe_rhs_fields* r, const ampe_rhs_fields* z)
{
   if (!c->precond) return -1;
   Precond& P = *(Precond*)c->precond;
   const ampe_rhs_config& p = c->cfg;
   const size_t nc = ncell_of(c);
   if (p.with_phase) P.phase->solve(r->phase, z->phase, P.ncycles, false);
   if (p.evolve_quat) {
      const double* rq = r->quat;
      if (P.has_dquatdphi) {
         // QuatPrecondSolve (QuatIntegrator.cc:3602-3612): rhs_q = r_q + gamma [dF_q/dphi] z_phase
         multiply_dquatdphi_block(c, P, z->phase, P.quat_rhs.v);
         for (size_t o = 0; o < nc * (size_t)p.qlen; o++) P.quat_rhs.data[o] = P.gamma * P.quat_rhs.data[o] + r->quat[o];
         rq = P.quat_rhs.data.data();
      }
      P.quat->solve(rq, z->quat, P.ncycles, true);  // all qlen components in one pass per sweep
   }
   if (p.with_unsteady_temperature) P.temp->solve(r->temperature, z->temperature, P.ncycles, false);
   if (p.with_concentration) {
      if (P.conc)
         P.conc->solve(r->conc, z->conc, P.ncycles, false);
      else if (z->conc != r->conc)
         memcpy(z->conc, r->conc, sizeof(double) * nc);
   }
   return 0;
}

// out = A_block u with the restated reference operators of part (1); block: 0 phase, 1 quaternion
// (one component, the system QuatSysSolver hands to the level solver: I + gamma sqrt_m L sqrt_m),
// 2 composition, 3 temperature
int precond_apply(Ctx* c, int block, const double* u, double* out)
{
   if (!c->precond) return -1;
   Precond& P = *(Precond*)c->precond;
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   const int D = p.ndim;
   auto const_side = [&](double v) {
      if (P.const_d.a[0].data.empty()) P.const_d.alloc(b, 0, 1);
      for (int a = 0; a < D; a++)
         for (auto& x : P.const_d.a[a].data) x = v;
   };
   auto const_cell = [&](Field& f, double v) {
      if (f.data.empty()) f.alloc(b, -1, 0, 1);
      for (auto& x : f.data) x = v;
   };
   View d[3];
   if (block == 0 && p.with_phase) {
      const_side(-P.gamma * p.epsilon_phase * p.epsilon_phase);
      views3(P.const_d, d, D);
      elliptic_apply(b, p.dx, c->phase_mobility.v, P.phase_c.v, d, u, out);
      return 0;
   }
   if (block == 1 && p.evolve_quat) {
      views3(c->face_coef, d, D);
      quat_stencil_apply(b, p.dx, P.gamma, P.sqrt_m.v, d, u, out);
      return 0;
   }
   if (block == 2 && P.conc) {
      const_cell(P.conc_m, p.conc_mobility);
      const_cell(P.ones, 1.0);
      views3(P.conc_d, d, D);
      elliptic_apply(b, p.dx, P.conc_m.v, P.ones.v, d, u, out);
      return 0;
   }
   if (block == 3 && p.with_unsteady_temperature) {
      const_side(-P.gamma * p.thermal_diffusivity);
      const_cell(P.ones, 1.0);
      views3(P.const_d, d, D);
      elliptic_apply(b, p.dx, P.ones.v, P.ones.v, d, u, out);
      return 0;
   }
   return -1;
}

int precond_dquatdphi(Ctx* c, const double* z_phase, double* out)
{
   if (!c->precond || !((Precond*)c->precond)->has_dquatdphi) return -1;
   Precond& P = *(Precond*)c->precond;
   multiply_dquatdphi_block(c, P, z_phase, P.quat_rhs.v);
   memcpy(out, P.quat_rhs.data.data(), sizeof(double) * P.quat_rhs.data.size());
   return 0;
}

HostMG* precond_block(Ctx* c, int block)
{
   if (!c->precond) return nullptr;
   Precond& P = *(Precond*)c->precond;
   return block == 0 ? P.phase.get() : block == 1 ? P.quat.get() : block == 2 ? P.conc.get() : P.temp.get();
}

}  // namespace oracle
