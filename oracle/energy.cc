// TEST INFRASTRUCTURE ONLY (see oracle.h).  CPU restatement of the scalar energy
// diagnostics: QuatModel::evaluateEnergy (source/QuatModel.cc:4888-4976) ->
// TwoPhasesEnergyEvaluationStrategy::evaluateEnergy (…Strategy.cc:60-260) -> the Fortran
// routines quatenergy / phi_interface_energy / interface_anisotropic_energy / bulkenergy
// (source/fortran/2d/quatenergy.m4:12-455, 3d/quatenergy.m4:166-520), same loop order,
// sequential accumulation.  weight = cell volume (single uniform level).
//
// The Cahn-Hilliard model (benchmarks/PFHub1a) has no energy evaluator in the reference
// (evaluateEnergy is skipped when with_phase() is false, QuatModel.cc:4953-4958); the PFHub 1a
// benchmark defines F = sum_cells [ w (c-ca)^2 (cb-c)^2 + kappa/2 |grad c|^2 ] dV, whose
// variational derivative is the mu of add_cahnhilliarddoublewell_flux
// (2d/concentrationrhs.m4:85-137).  That definition is restated here with the face-centred
// gradient the flux uses; it is "parity unpinned" (no reference code or golden value).
#include <cmath>
#include <cstring>

#include "ctx.h"

namespace oracle {

static inline int wrapi(int i, int n)
{
   i %= n;
   return i < 0 ? i + n : i;
}
static void fill_periodic_e(const Box& b, const double* src, Field& dst, int ng, int depth)
{
   const int n0 = b.hi[0] + 1, n1 = b.hi[1] + 1, n2 = b.hi[2] + 1;
   const int g2 = (b.ndim == 3) ? ng : 0;
   for (int m = 0; m < depth; m++)
      for (int k = -g2; k < n2 + g2; k++)
         for (int j = -ng; j < n1 + ng; j++)
            for (int i = -ng; i < n0 + ng; i++) {
               const int is = wrapi(i, n0), js = wrapi(j, n1), ks = wrapi(k, n2);
               dst.v(i, j, k, m) =
                   src[(size_t)is + (size_t)n0 * (js + (size_t)n1 * (ks + (size_t)n2 * m))];
            }
}

// out[0] total, [1] phase interface, [2] orientational, [3] q interface, [4] well, [5] bulk free
int energy(Ctx* c, const ampe_rhs_fields* y, double* out)
{
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   const int D = p.ndim, Q = p.qlen, ng = c->ng;
   for (int n = 0; n < 8; n++) out[n] = 0.0;
   double weight = 1.0;
   for (int d = 0; d < D; d++) weight *= p.dx[d];
   int status = 0;

   if (p.conc_rhs_form == AMPE_CONC_CAHN_HILLIARD && p.with_concentration) {
      fill_periodic_e(b, y->conc, c->conc, ng, 1);
      double tot = 0.0, well = 0.0, grad = 0.0;
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
         for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) {
               const double cc = c->conc.v(i, j, k);
               const double w = p.ch_well_scale * (cc - p.ch_ca) * (cc - p.ch_ca) * (p.ch_cb - cc) *
                                (p.ch_cb - cc);
               // |grad c|^2 as the average of the two face gradients per direction
               double g2 = 0.0;
               for (int d = 0; d < D; d++) {
                  const double gl = (cc - c->conc.v(i - (d == 0), j - (d == 1), k - (d == 2))) / p.dx[d];
                  const double gu = (c->conc.v(i + (d == 0), j + (d == 1), k + (d == 2)) - cc) / p.dx[d];
                  g2 = g2 + 0.5 * (gl * gl + gu * gu);
               }
               const double eg = 0.5 * p.ch_kappa * g2;
               well = well + w * weight;
               grad = grad + eg * weight;
               tot = tot + (w + eg) * weight;
            }
      out[0] = tot, out[1] = grad, out[4] = well;
      return 0;
   }
   if (!p.with_phase) return 0;

   // copyCurrentToScratch
   if (!p.with_unsteady_temperature)
      for (auto& v : c->temp.data) v = p.T_uniform;
   fill_periodic_e(b, y->phase, c->phase, ng, 1);
   if (Q > 0) fill_periodic_e(b, y->quat, c->quat, ng, Q);
   if (p.with_concentration) fill_periodic_e(b, y->conc, c->conc, ng, 1);
   if (p.with_unsteady_temperature) fill_periodic_e(b, y->temperature, c->temp, ng, 1);

   View gside[3];
   if (p.evolve_quat) {
      View diffs_symm[3];
      for (int d = 0; d < D; d++) {
         diffs_symm[d] = c->quat_diffs.a[d].v.at(0);
         gside[d] = c->quat_grad_side.a[d].v.at(0);
      }
      View diffs_nonsymm[3];
      for (int d = 0; d < D; d++) diffs_nonsymm[d] = c->quat_diffs.a[d].v.at(p.symmetry_aware ? Q : 0);
      quatdiffs(b, Q, c->quat.v, diffs_nonsymm);
      if (p.symmetry_aware) {
         quatdiffs_symm(b, Q, c->quat.v, diffs_symm, c->iqrot);
         quatgrad_side_symm(b, Q, p.dx, diffs_symm, gside, c->iqrot);
      } else {
         quatgrad_side(b, Q, p.dx, diffs_symm, gside);
      }
   }
   const bool kks = p.with_concentration &&
                    (p.conc_rhs_form == AMPE_CONC_KKS || p.conc_rhs_form == AMPE_CONC_EBS);
   if (kks) {
      if (compute_phase_concentrations(c) < 0) status = AMPE_ENEWTON;
      compute_free_energies(c);
   }

   const double mf = 2.0 * p.H_parameter;
   const double floor2 = (p.grad_floor_type == 's') ? p.quat_grad_floor * p.quat_grad_floor : 0.0;
   double total = 0.0, phi_e = 0.0, orient_e = 0.0, qint_e = 0.0, well_e = 0.0, free_e = 0.0;
   const View& phi = c->phase.v;

   // ---- phi interface energy ----
   if (p.epsilon_anisotropy > 0.0 && D == 3) return AMPE_EINVAL;  // 3D anisotropic energy: not on the path
   if (p.epsilon_anisotropy > 0.0) {
      // interface_anisotropic_energy (2d/quatenergy.m4:12-103)
      const double pi = 4.0 * atan(1.0);
      const double dxinv = 0.5 / p.dx[0], dyinv = 0.5 / p.dx[1];
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
         for (int i = b.lo[0]; i <= b.hi[0]; i++) {
            const double dphidx = (phi(i + 1, j) - phi(i - 1, j)) * dxinv;
            const double dphidy = (phi(i, j + 1) - phi(i, j - 1)) * dyinv;
            double theta;
            if (fabs(dphidx) > (double)1.e-12f)
               theta = atan(dphidy / dphidx);
            else
               theta = 0.5 * pi;
            double q = c->quat.v(i, j, 0, 0);
            if (q > 1.0) q = 1.0;
            if (q < -1.0) q = -1.0;
            const double refangle = (Q == 4) ? 2.0 * acos(q) : acos(q);
            const double epstheta =
                p.epsilon_phase * (1.0 + p.epsilon_anisotropy * cos(p.knumber * (theta - refangle)));
            const double diff_term = dphidx * dphidx + dphidy * dphidy;
            double e = 0.5 * epstheta * epstheta * diff_term;
            e = e * weight;
            phi_e = phi_e + e;
         }
   } else {
      // phi_interface_energy (2d/quatenergy.m4:108-181)
      const double e2 = 0.5 * p.epsilon_phase * p.epsilon_phase;
      double d2inv[3] = {0, 0, 0};
      for (int d = 0; d < D; d++) d2inv[d] = 1.0 / (p.dx[d] * p.dx[d]);
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
         for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) {
               const double tx = d2inv[0] * (-phi(i + 1, j, k) + 2.0 * phi(i, j, k) - phi(i - 1, j, k));
               const double ty = d2inv[1] * (-phi(i, j + 1, k) + 2.0 * phi(i, j, k) - phi(i, j - 1, k));
               double diff_term = tx + ty;
               if (D == 3)
                  diff_term = diff_term +
                              d2inv[2] * (-phi(i, j, k + 1) + 2.0 * phi(i, j, k) - phi(i, j, k - 1));
               double e = e2 * diff_term * phi(i, j, k);
               e = e * weight;
               phi_e = phi_e + e;
            }
   }
   total = total + phi_e;

   // ---- orientational + q interface energy ----
   if (mf > 0.0 && p.evolve_quat) {
      auto side_g2 = [&](int d, int i, int j, int k) {
         double o2 = 0.0;
         if (D == 2) {  // 2d: do n / do m;  3d: do m / do n
            for (int n = 0; n < D; n++)
               for (int m = 0; m < Q; m++) {
                  const double g = gside[d](i, j, k, m + Q * n);
                  o2 = o2 + g * g;
               }
         } else {
            for (int m = 0; m < Q; m++)
               for (int n = 0; n < D; n++) {
                  const double g = gside[d](i, j, k, m + Q * n);
                  o2 = o2 + g * g;
               }
         }
         return o2;
      };
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
         for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) {
               double e = 0.0;
               for (int d = 0; d < D; d++)
                  for (int up = 0; up < 2; up++) {
                     const int il = i - (d == 0) * (1 - up), jl = j - (d == 1) * (1 - up),
                               kl = k - (d == 2) * (1 - up);
                     const int iu = il + (d == 0), ju = jl + (d == 1), ku = kl + (d == 2);
                     // lower face: average_func(phi(i-1), phi(i)); upper: (phi(i), phi(i+1));
                     // the 3D z lower face is written (phi(i,j,k), phi(i,j,k-1)) in the reference
                     double aphi;
                     if (D == 3 && d == 2 && up == 0)
                        aphi = average_func(phi(iu, ju, ku), phi(il, jl, kl), p.avg_func);
                     else
                        aphi = average_func(phi(il, jl, kl), phi(iu, ju, ku), p.avg_func);
                     const double p_phi = interp_func(aphi, p.orient_interp1);
                     double o2 = side_g2(d, iu, ju, ku);
                     // 3d/quatenergy.m4:386-387: the upper y face takes the root BEFORE the floor
                     if (!(D == 3 && d == 1 && up == 1)) o2 = o2 + floor2;
                     e = e + sqrt(o2) * p_phi;
                  }
               e = e * c->temp.v(i, j, k);
               if (D == 2)
                  e = e * 0.25 * mf;
               else
                  e = e * mf / 6.0;
               e = e * weight;
               total = total + e;
               orient_e = orient_e + e;
            }
      const double epsilonq2 = 0.5 * p.epsilon_q * p.epsilon_q;
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
         for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) {
               const double p_phi = interp_func(phi(i, j, k), p.orient_interp2);
               double e = 0.0;
               for (int mn = 0; mn < Q * D; mn++) {
                     const int m = (D == 2) ? mn % Q : mn / D, n = (D == 2) ? mn / Q : mn % D;
                     double s = 0.0;
                     for (int d = 0; d < D; d++) {
                        const double gl = gside[d](i, j, k, m + Q * n);
                        const double gu = gside[d](i + (d == 0), j + (d == 1), k + (d == 2), m + Q * n);
                        s = (d == 0) ? (gl * gl + gu * gu) : s + gl * gl + gu * gu;
                     }
                     e = e + s;
                  }
               if (D == 2)
                  e = e * 0.25 * epsilonq2 * p_phi;
               else
                  e = e * p_phi * epsilonq2 / 6.0;
               e = e * weight;
               total = total + e;
               qint_e = qint_e + e;
            }
   }

   // ---- double well ----
   for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
         for (int i = b.lo[0]; i <= b.hi[0]; i++) {
            double e = p.phi_well_scale * well_func(phi(i, j, k), 'd');
            e = e * weight;
            total = total + e;
            well_e = well_e + e;
         }

   // ---- bulkenergy (2d/quatenergy.m4:403-455); f_l, f_a stay 0 for the bias double well
   if (kks) {
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
         for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) {
               const double h_phi = interp_func(phi(i, j, k), p.energy_interp);
               double e = (1.0 - h_phi) * c->f_l.v(i, j, k) + h_phi * c->f_a.v(i, j, k);
               e = e * weight;
               total = total + e;
               free_e = free_e + e;
            }
   }
   out[0] = total, out[1] = phi_e, out[2] = orient_e, out[3] = qint_e, out[4] = well_e,
   out[5] = free_e;
   return status;
}

// QuatModel::printScalarDiagnostics (QuatModel.cc:2543-2690) on one uniform level: the weights of
// HierarchyCellDataOpsReal::{L1Norm, integral} are the cell volume.  out[12]: see ampe_scalar_diagnostics.
int scalar_diagnostics(Ctx* c, const ampe_rhs_fields* y, double* out)
{
   const ampe_rhs_config& p = c->cfg;
   size_t ncell = 1;
   double weight = 1.0;
   for (int d = 0; d < p.ndim; d++) {
      ncell *= (size_t)p.n[d];
      weight = weight * p.dx[d];
   }
   for (int n = 0; n < 12; n++) out[n] = 0.0;
   const double vol = weight * (double)ncell;
   out[0] = vol;
   double vphi = vol;  // QuatModel.cc:2606
   double sum_phi = 0.0;
   if (p.with_phase) {
      // evaluateVolumeSolid (:5170): L1Norm(phase, weight)
      double l1 = 0.0;
      for (size_t i = 0; i < ncell; i++) {
         l1 = l1 + fabs(y->phase[i]) * weight;
         sum_phi = sum_phi + y->phase[i] * weight;
      }
      vphi = l1;
   }
   out[1] = vphi;
   out[2] = vphi / vol;
   if (p.with_concentration) {
      // evaluateIntegralConcentration (:5106), max, evaluateIntegralPhaseConcentration (:5130)
      double c0V0 = 0.0, cmax = y->conc[0], cphi = 0.0;
      for (size_t i = 0; i < ncell; i++) {
         c0V0 = c0V0 + y->conc[i] * weight;
         if (y->conc[i] > cmax) cmax = y->conc[i];
         const double prod = y->conc[i] * (p.with_phase ? y->phase[i] : 1.0);
         cphi = cphi + fabs(prod) * weight;
      }
      const double c0 = c0V0 / vol;
      out[3] = c0V0, out[4] = cmax, out[5] = cphi;
      out[6] = (cphi - c0 * vphi) / c0V0;  // :2655
   }
   if (p.with_unsteady_temperature) {
      double tmin = y->temperature[0], tmax = y->temperature[0], tint = 0.0, cenergy = 0.0;
      for (size_t i = 0; i < ncell; i++) {
         const double t = y->temperature[i];
         if (t < tmin) tmin = t;
         if (t > tmax) tmax = t;
         tint = tint + t * weight;
         cenergy = cenergy + (p.cp * t) * weight;  // computeThermalEnergy (:5373): multiply(cp, T), integral
      }
      out[7] = tmin, out[8] = tmax, out[9] = tint / vol;
      out[10] = (p.with_phase ? sum_phi * (-1. * p.latent_heat) : 0.0) + cenergy;
   } else {
      out[7] = out[8] = out[9] = p.T_uniform;
   }
   return 0;
}

}  // namespace oracle
