// TEST INFRASTRUCTURE ONLY (see oracle.h).  Restatement of the C++ per-cell /
// per-face loops of the reference's CALPHAD, Quadratic, EBS and KKS strategies.
#include "ctx.h"
#include <algorithm>
#include <cmath>

namespace oracle {

static inline int E(int a, int d) { return a == d ? 1 : 0; }

static Quadratic quad_params(const ampe_rhs_config& p)
{
   Quadratic q;
   q.Tref = p.quad_Tref;
   q.A[0] = p.quad_A_l;
   q.Ceq[0] = p.quad_Ceq_l;
   q.m[0] = p.quad_m_l;
   q.A[1] = p.quad_A_s;
   q.Ceq[1] = p.quad_Ceq_s;
   q.m[1] = p.quad_m_s;
   return q;
}

// CALPHADequilibriumPhaseConcentrationsStrategy.cc:162-454 (loop on the ghost box,
// init x = (c_l_ref, c_a_ref), hphi = interp_func(conc_interp, phi));
// QuadraticEquilibriumPhaseConcentrationsStrategy.cc:42-144 (closed form).
int compute_phase_concentrations(Ctx* c)
{
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   const int ng = c->ng;
   int L[3], H[3];
   for (int d = 0; d < 3; d++) {
      int g = (d < b.ndim) ? ng : 0;
      L[d] = b.lo[d] - g;
      H[d] = b.hi[d] + g;
   }
   int nfail = 0;
   const Quadratic quad = quad_params(p);
   ORACLE_PAR_RED(nfail)
   for (int k = L[2]; k <= H[2]; k++)
      for (int j = L[1]; j <= H[1]; j++)
         for (int i = L[0]; i <= H[0]; i++) {
            const double temp = c->temp.v(i, j, k);
            const double phi = c->phase.v(i, j, k);
            const double hphi = interp_func(phi, p.conc_interp);
            const double conc = c->conc.v(i, j, k);
            double x[2];
            if (p.free_energy == AMPE_FE_CALPHAD) {
               x[0] = c->cl_ref.v(i, j, k);
               x[1] = c->ca_ref.v(i, j, k);
               int st = calphad_phase_concentrations(p.calphad, temp, conc, hphi, x,
                                                     p.newton_tol, p.newton_max_its,
                                                     p.newton_alpha);
               if (st < 0) nfail++;
            } else {
               quadratic_phase_concentrations(quad, temp, conc, hphi, x);
            }
            c->cl.v(i, j, k) = x[0];
            c->ca.v(i, j, k) = x[1];
         }
   return nfail ? -nfail : 0;
}

// CALPHADFreeEnergyStrategyBinary.cc:251-327 / QuadraticFreeEnergyStrategy.cc:171-247:
// f_i = computeFreeEnergy(T, c_i, phase_i) * (1e-6 / V_m^i)
void compute_free_energies(Ctx* c)
{
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   const double inv_vm_l = 1.e-6 / p.vm_liquid;
   const double inv_vm_a = 1.e-6 / p.vm_solid;
   const Quadratic quad = quad_params(p);
   for (int pass = 0; pass < 2; pass++)
      ORACLE_PAR
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
         for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++) {
               const double t = c->temp.v(i, j, k);
               const double ci = pass == 0 ? c->cl.v(i, j, k) : c->ca.v(i, j, k);
               double f;
               if (p.free_energy == AMPE_FE_CALPHAD)
                  f = calphad_free_energy(p.calphad, t, ci, pass);
               else
                  f = quadratic_free_energy(quad, t, ci, pass);
               f *= (pass == 0 ? inv_vm_l : inv_vm_a);
               (pass == 0 ? c->f_l : c->f_a).v(i, j, k) = f;
            }
}

// CALPHADFreeEnergyStrategyBinary.cc:521-683 (mu = dfa/dc(c_a) /V_m^A)
// QuadraticFreeEnergyStrategy.cc:393-530  (mu = dfl/dc(c_l) /V_m^L)
void add_driving_force(Ctx* c)
{
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   const Quadratic quad = quad_params(p);
   ORACLE_PAR
   for (int k = b.lo[2]; k <= b.hi[2]; k++)
      for (int j = b.lo[1]; j <= b.hi[1]; j++)
         for (int i = b.lo[0]; i <= b.hi[0]; i++) {
            const double t = c->temp.v(i, j, k);
            const double phi = c->phase.v(i, j, k);
            const double f_l = c->f_l.v(i, j, k);
            const double f_a = c->f_a.v(i, j, k);
            const double c_l = c->cl.v(i, j, k);
            const double c_a = c->ca.v(i, j, k);
            const double hphi_prime = deriv_interp_func(phi, p.energy_interp);
            if (p.free_energy == AMPE_FE_CALPHAD) {
               double mu = calphad_deriv_free_energy(p.calphad, t, c_a, 1);
               mu *= 1.e-6 / p.vm_solid;
               const double heta = 0.0, f_b = 0.0, c_b = 0.0;
               c->rhs_phase.v(i, j, k) +=
                   hphi_prime * ((f_l - (1.0 - heta) * f_a - heta * f_b) -
                                 mu * (c_l - (1.0 - heta) * c_a - heta * c_b));
            } else {
               const double deriv = quadratic_deriv_free_energy(quad, t, c_l, 0);
               const double mu = deriv * (1.e-6 / p.vm_liquid);
               c->rhs_phase.v(i, j, k) += hphi_prime * ((f_l - f_a) - mu * (c_l - c_a));
            }
         }
}

// QuatIntegrator::setDiffusionCoeffForConcentration (QuatIntegrator.cc:2388-2432)
void set_diffusion_coeff_for_concentration(Ctx* c)
{
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   if (p.conc_rhs_form == AMPE_CONC_EBS) {
      // MobilityCompositionDiffusionStrategy::setDiffCoeffInEachPhaseOnPatch (:197-407)
      // then setPFMDiffOnPatch (:409-611); stored already multiplied by (1-h) / h.
      for (int a = 0; a < b.ndim; a++)
         ORACLE_PAR
         for (int k = b.lo[2]; k <= b.hi[2] + E(a, 2); k++)
            for (int j = b.lo[1]; j <= b.hi[1] + E(a, 1); j++)
               for (int i = b.lo[0]; i <= b.hi[0] + E(a, 0); i++) {
                  const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
                  if (p.free_energy != AMPE_FE_CALPHAD) {
                     // diffusion_type "temperature_dependent": TbasedCompositionDiffusionStrategy::setDiffusion ->
                     // concentration_pfmdiffusion_of_temperature (2d/concentrationdiffusion.m4:341-430): Arrhenius
                     // diffusivity of each phase weighted with the phase fraction at the face; avg_func_type and
                     // diffusion_interp_func_type (CompositionDiffusionStrategyFactory.h:59-72)
                     const double vphi = average_func(c->phase.v(im, jm, km), c->phase.v(i, j, k), p.avg_func);
                     const double hphi = interp_func(vphi, p.diffusion_interp);
                     const double invT = 2.0 / (c->temp.v(im, jm, km) + c->temp.v(i, j, k));
                     const double diff_liquid = p.D_liquid * exp(-(p.Q0_liquid / GASCONSTANT_R_JPKPMOL) * invT);
                     const double diff_solid = p.D_solid * exp(-(p.Q0_solid / GASCONSTANT_R_JPKPMOL) * invT);
                     c->diff_l.a[a].v(i, j, k) = (1.0 - hphi) * diff_liquid;
                     c->diff_a.a[a].v(i, j, k) = hphi * diff_solid;
                     continue;
                  }
                  const double temp = 0.5 * (c->temp.v(i, j, k) + c->temp.v(im, jm, km));
                  const double c_l = 0.5 * (c->cl.v(i, j, k) + c->cl.v(im, jm, km));
                  const double c_a = 0.5 * (c->ca.v(i, j, k) + c->ca.v(im, jm, km));
                  // computeLocalDiffusionMatrix{L,A}: d2f [J/mol] * mobility
                  const double d2fl = calphad_second_deriv_free_energy(p.calphad, temp, c_l, 0);
                  const double mobl = calphad_diffusion_mobility_binary(p.calphad, 0, c_l, temp);
                  const double dl = mobl * d2fl;
                  const double d2fa = calphad_second_deriv_free_energy(p.calphad, temp, c_a, 1);
                  const double moba = calphad_diffusion_mobility_binary(p.calphad, 1, c_a, temp);
                  const double da = moba * d2fa;
                  const double phi =
                      average_func(c->phase.v(i, j, k), c->phase.v(im, jm, km), p.conc_avg_func);
                  const double hphi = interp_func(phi, p.diffusion_interp);
                  c->diff_l.a[a].v(i, j, k) = (1. - hphi) * dl;
                  c->diff_a.a[a].v(i, j, k) = hphi * da;
               }
   } else if (p.conc_rhs_form == AMPE_CONC_KKS) {
      // KKSCompositionRHSStrategy::setDiffusionCoeff (:72-94)
      View d0[3], dphi[3];
      for (int a = 0; a < b.ndim; a++) {
         d0[a] = c->diff0.a[a].v;
         dphi[a] = c->dphi.a[a].v;
      }
      concentration_pfmdiffusion(b, c->phase.v, d0, c->temp.v, p.D_liquid, p.Q0_liquid,
                                 p.D_solid, p.Q0_solid, GASCONSTANT_R_JPKPMOL, p.energy_interp,
                                 p.conc_avg_func);
      // setDiffCoeffForPhaseOnPatch (:210-373)
      for (int a = 0; a < b.ndim; a++)
         ORACLE_PAR
         for (int k = b.lo[2]; k <= b.hi[2] + E(a, 2); k++)
            for (int j = b.lo[1]; j <= b.hi[1] + E(a, 1); j++)
               for (int i = b.lo[0]; i <= b.hi[0] + E(a, 0); i++) {
                  const int im = i - E(a, 0), jm = j - E(a, 1), km = k - E(a, 2);
                  const double phi =
                      average_func(c->phase.v(i, j, k), c->phase.v(im, jm, km), p.conc_avg_func);
                  const double c_l = 0.5 * (c->cl.v(i, j, k) + c->cl.v(im, jm, km));
                  const double c_a = 0.5 * (c->ca.v(i, j, k) + c->ca.v(im, jm, km));
                  const double hphi_prime = deriv_interp_func(phi, p.energy_interp);
                  dphi[a](i, j, k) = d0[a](i, j, k) * hphi_prime * (c_l - c_a);
               }
   }
}

// EBSCompositionRHSStrategy::computeFluxOnPatch (EBSCompositionRHSStrategy.cc:174-318)
// KKSCompositionRHSStrategy::computeFluxOnPatch (KKSCompositionRHSStrategy.cc:377-441)
void compute_conc_flux_kks_ebs(Ctx* c)
{
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   View flux[3];
   for (int a = 0; a < b.ndim; a++) flux[a] = c->conc_flux.a[a].v;
   if (p.conc_rhs_form == AMPE_CONC_EBS) {
      View dl[3], da[3];
      for (int a = 0; a < b.ndim; a++) {
         std::fill(c->conc_flux.a[a].data.begin(), c->conc_flux.a[a].data.end(), 0.0);
         dl[a] = c->diff_l.a[a].v;
         da[a] = c->diff_a.a[a].v;
      }
      add_flux(b, p.dx, c->cl.v, 1, dl, flux);
      add_flux(b, p.dx, c->ca.v, 1, da, flux);
   } else {
      View d0[3], dphi[3];
      for (int a = 0; a < b.ndim; a++) {
         d0[a] = c->diff0.a[a].v;
         dphi[a] = c->dphi.a[a].v;
      }
      concentrationflux(b, p.dx, c->conc.v, c->phase.v, d0, dphi, flux);
   }
}

}  // namespace oracle
