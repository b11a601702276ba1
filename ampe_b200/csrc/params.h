// Device-side parameter record of one RHS evaluation, derived on the host from
// ampe_rhs_config.  Every derived constant is formed with exactly the expression
// the reference kernel uses for it (cited), so that host precomputation does not
// change a single bit of the result.
#pragma once
#include "../../include/ampe_b200.h"

namespace ampe {

// T-dependent CALPHAD coefficients.  T is spatially uniform in every CALPHAD
// configuration (ScalarTemperatureStrategy), so they are evaluated once per
// evaluation on the host (Thermo4PFM computeTdependentParameters is called per
// cell in the reference).
struct CalphadT {
   double fA[2], fB[2];  // species Gibbs energies per phase [L, A]
   double L[2][4];       // Redlich-Kister L_k(T) per phase
   double RT, RTinv;     // R*T, 1/(R*T)
   // mobility Q(T) = a0 + R T ln a1 : [species][phase]
   double qA[2][2], qB[2][2], qAB[2][2][4];
};

struct Params {
   int ndim, qlen;
   int n[3];      // local interior cells
   int ng;        // ghost width of the state (1; 2 for Cahn-Hilliard)
   int with_phase, with_conc, with_T, evolve_quat;
   int flux_type, conc_form, free_energy, symm, modulus_from_cells;
   int clamp[3];   // 1: zero-slope physical boundary in this direction (ghost = adjacent interior cell), 0: periodic
   int libm_trig;  // 1: evaluate the anisotropy with atan/acos/sincos exactly as written in the reference
   char energy_interp, conc_interp, diffusion_interp, orient_interp1, orient_interp2;
   char avg_func, conc_avg_func, grad_floor_type, quat_mobility_func;

   double h[3];
   double dinv[3];       // 1/h
   double p5inv[3];      // 0.5/h                      quatgrad_cell, compute_lambda_flux
   double p25inv[3];     // 0.25*(1/h)                 quatgrad_side
   double dinv2[3];      // 1/(h*h)                    laplacian, correctrhsquatforsymmetry
   double eps2_dinv[3];  // (eps*eps)/h                gradient_flux
   double iso_dinv[2];   // (1/12)*eps*eps/h           compute_flux_isotropic
   double ch_dinv2[3];   // (1/h)*(1/h)                add_cahnhilliarddoublewell_flux
   double ch_mdinv[3];   // ch_mobility*(1/h)

   double epsilon_phase, nu;
   int knumber;
   double phi_well_scale, phi_mobility;
   double misorientation_factor;  // 2*H (PhaseRHSStrategyWithQ.cc:250, QuatFaceCoeff.cc:93)
   double epsilonq2_half;         // 0.5*eps_q*eps_q   computerhspbg
   double epsq2;                  // eps_q*eps_q       compute_face_coef
   double floor2, max_normi;      // floor**2, 1/floor
   double quat_mobility, min_quat_mobility, quat_mobility_alt;
   double T_uniform;
   double thermal_diffusivity, latent_heat, cp, meltingT;
   double latent_over_cp;          // latent_heat/cp   computerhstemp
   double bias_coeff, bias_gamma;  // alpha/pi_f32, gamma   computerhsbiaswell
   double deltaT_alpha;            // latent_heat/Tm        computerhsdeltatemperature
   double conc_mobility;
   double ch_ca, ch_cb, ch_well_scale, ch_kappa;
   // quadratic
   double quad_A[2], quad_ceq[2];  // ceq(T) = Ceq + (T-Tref)*m   (uniform T)
   double quad_rla, quad_ral;      // A_l/A_a, A_a/A_l
   double q0_liquid_invR, q0_solid_invR, D_liquid, D_solid;
   double inv_vm_l, inv_vm_a;      // 1e-6/V_m
   int newton_max_its;
   double newton_tol, newton_alpha;
   CalphadT ct;
};

}  // namespace ampe
