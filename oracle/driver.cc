// TEST INFRASTRUCTURE ONLY (see oracle.h).  CPU restatement of
// QuatIntegrator::evaluateRHSFunction (source/QuatIntegrator.cc:3134-3295) on a
// single uniform periodic level: same order of passes, every intermediate
// materialised in an array between passes exactly like the reference.
#include "ctx.h"
#include "precond.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace oracle {

static void alloc_cell(Field& f, const Box& b, int ng, int depth) { f.alloc(b, -1, ng, depth); }

Ctx* create(const ampe_rhs_config& cfg)
{
   Ctx* c = new Ctx;
   c->cfg = cfg;
   c->T0 = cfg.T_uniform;
   Box& b = c->box;
   b.ndim = cfg.ndim;
   for (int d = 0; d < 3; d++) {
      b.lo[d] = 0;
      b.hi[d] = (d < cfg.ndim) ? cfg.n[d] - 1 : 0;
   }
   c->ng = (cfg.conc_rhs_form == AMPE_CONC_CAHN_HILLIARD) ? 2 : 1;
   const int ng = c->ng, Q = cfg.qlen, D = cfg.ndim;
   if (cfg.with_phase) {
      alloc_cell(c->phase, b, ng, 1);
      alloc_cell(c->phase_mobility, b, 1, 1);
      c->phase_flux.alloc(b, 0, 1);
      alloc_cell(c->rhs_phase, b, 0, 1);
   }
   alloc_cell(c->temp, b, ng, 1);
   if (Q > 0) alloc_cell(c->quat, b, ng, Q);
   if (cfg.evolve_quat) {
      c->quat_diffs.alloc(b, 1, cfg.symmetry_aware ? 2 * Q : Q);
      c->quat_grad_side.alloc(b, 0, D * Q);
      c->quat_grad_side_copy.alloc(b, 0, D * Q);
      for (int d = 0; d < D; d++) alloc_cell(c->quat_grad_cell[d], b, 0, Q);
      alloc_cell(c->quat_grad_modulus, b, 0, 1);
      alloc_cell(c->quat_mobility, b, 1, 1);
      c->face_coef.alloc(b, 0, 1);
      c->quat_flux.alloc(b, 0, Q);
      alloc_cell(c->lambda, b, 0, 1);
      alloc_cell(c->rhs_quat, b, 0, Q);
   }
   if (cfg.symmetry_aware) {
      for (int d = 0; d < D; d++) {
         c->iqrot_data[d].assign(view_size(b, d, 1, 1), 1);
         c->iqrot[d] = make_iview(c->iqrot_data[d].data(), b, d, 1);
      }
   }
   if (cfg.with_concentration) {
      alloc_cell(c->conc, b, ng, 1);
      c->conc_flux.alloc(b, cfg.conc_rhs_form == AMPE_CONC_CAHN_HILLIARD ? 1 : 0, 1);
      alloc_cell(c->rhs_conc, b, 0, 1);
      if (cfg.conc_rhs_form == AMPE_CONC_KKS || cfg.conc_rhs_form == AMPE_CONC_EBS) {
         alloc_cell(c->cl, b, ng, 1);
         alloc_cell(c->ca, b, ng, 1);
         alloc_cell(c->cl_ref, b, ng, 1);
         alloc_cell(c->ca_ref, b, ng, 1);
         alloc_cell(c->f_l, b, 0, 1);
         alloc_cell(c->f_a, b, 0, 1);
      }
      if (cfg.conc_rhs_form == AMPE_CONC_EBS) {
         c->diff_l.alloc(b, 0, 1);
         c->diff_a.alloc(b, 0, 1);
      }
      if (cfg.conc_rhs_form == AMPE_CONC_KKS) {
         c->diff0.alloc(b, 0, 1);
         c->dphi.alloc(b, 0, 1);
      }
   }
   if (cfg.with_unsteady_temperature) {
      alloc_cell(c->cp, b, 0, 1);
      alloc_cell(c->rhs_temp, b, 0, 1);
   }
   if (cfg.free_energy == AMPE_FE_BIASWELL) alloc_cell(c->te, b, 0, 1);
   return c;
}

void destroy(Ctx* c)
{
   precond_destroy(c);
   delete c;
}

static inline int wrap(int i, int n)
{
   i %= n;
   return i < 0 ? i + n : i;
}
// ghost index -> source index: periodic image, or (zero-slope physical boundary: boundary_N = "slope", "0" in
// every BoundaryConditions block, CartesianRobinBcHelper with a = 0, b = 1, g = 0, whose type-1 and corner
// formulas both reduce to the adjacent interior value) the nearest interior cell
static inline int ghost_src(int i, int n, int zero_slope)
{
   if (zero_slope) return i < 0 ? 0 : (i >= n ? n - 1 : i);
   return wrap(i, n);
}

// fillScratch (QuatIntegrator.cc:2873-2955) on a fully periodic single patch:
// copy y -> scratch interior, ghosts = periodic images (incl. edges/corners).
static const int* g_zero_slope = nullptr;  // of the context being evaluated (set by eval / set_ref)
static void fill_periodic(const Box& b, const double* src, Field& dst, int ng, int depth)
{
   static const int none[3] = {0, 0, 0};
   const int* zs = g_zero_slope ? g_zero_slope : none;
   const int n0 = b.hi[0] + 1, n1 = b.hi[1] + 1, n2 = b.hi[2] + 1;
   const int g2 = (b.ndim == 3) ? ng : 0;
   for (int m = 0; m < depth; m++)
      ORACLE_PAR
      for (int k = -g2; k < n2 + g2; k++)
         for (int j = -ng; j < n1 + ng; j++)
            for (int i = -ng; i < n0 + ng; i++) {
               const int is = ghost_src(i, n0, zs[0]), js = ghost_src(j, n1, zs[1]), ks = ghost_src(k, n2, zs[2]);
               dst.v(i, j, k, m) =
                   src[(size_t)is + (size_t)n0 * (js + (size_t)n1 * (ks + (size_t)n2 * m))];
            }
}
static void fill_periodic_int(const Box& b, int axis, const int* src, IView dst, int ng)
{
   // side data: lower face of cell (i,j,k); periodic image of the face index
   const int n0 = b.hi[0] + 1, n1 = b.hi[1] + 1, n2 = b.hi[2] + 1;
   const int g2 = (b.ndim == 3) ? ng : 0;
   for (int k = -g2; k < n2 + g2 + (axis == 2); k++)
      for (int j = -ng; j < n1 + ng + (axis == 1); j++)
         for (int i = -ng; i < n0 + ng + (axis == 0); i++) {
            const int is = wrap(i, n0), js = wrap(j, n1), ks = wrap(k, n2);
            dst(i, j, k) = src[(size_t)is + (size_t)n0 * (js + (size_t)n1 * ks)];
         }
}
static void copy_out(const Box& b, const Field& src, double* dst, int depth)
{
   const int n0 = b.hi[0] + 1, n1 = b.hi[1] + 1, n2 = b.hi[2] + 1;
   for (int m = 0; m < depth; m++)
      ORACLE_PAR
      for (int k = 0; k < n2; k++)
         for (int j = 0; j < n1; j++)
            for (int i = 0; i < n0; i++)
               dst[(size_t)i + (size_t)n0 * (j + (size_t)n1 * (k + (size_t)n2 * m))] =
                   src.v(i, j, k, m);
}

void set_ref(Ctx* c, const double* cl_ref, const double* ca_ref)
{
   // resetRefPhaseConcentrations (QuatModel.cc:5218-5231): whole-array copy,
   // ghosts included (here ghosts = periodic images of the given arrays)
   g_zero_slope = c->cfg.zero_slope;
   if (cl_ref && ca_ref) {
      fill_periodic(c->box, cl_ref, c->cl_ref, c->ng, 1);
      fill_periodic(c->box, ca_ref, c->ca_ref, c->ng, 1);
   } else {
      c->cl_ref.data = c->cl.data;
      c->ca_ref.data = c->ca.data;
   }
   c->have_ref = true;
}

void set_rotations(Ctx* c, const int* const* iqrot)
{
   for (int d = 0; d < c->cfg.ndim; d++)
      fill_periodic_int(c->box, d, iqrot[d], c->iqrot[d], 1);
}

void get_phase_concentrations(Ctx* c, double* cl, double* ca)
{
   copy_out(c->box, c->cl, cl, 1);
   copy_out(c->box, c->ca, ca, 1);
}

static void views(SideField& s, View* v, int ndim, int m0 = 0)
{
   for (int d = 0; d < ndim; d++) v[d] = s.a[d].v.at(m0);
}


int eval(Ctx* c, double time, const ampe_rhs_fields* y, const ampe_rhs_fields* ydot,
         int fd_flag)
{
   g_zero_slope = c->cfg.zero_slope;
   // setTemperatureField (QuatIntegrator.cc:3191) -> ScalarTemperatureStrategy::getCurrentTemperature
   // (ScalarTemperatureStrategy.cc:57-75): the uniform temperature of this evaluation
   if (c->cfg.dtemperaturedt != 0.0) {
      double t = c->T0 + c->cfg.dtemperaturedt * time;
      if (c->cfg.dtemperaturedt < 0. && t < c->cfg.target_temperature)
         t = c->cfg.target_temperature;
      else if (c->cfg.target_temperature > 0.0 && c->cfg.dtemperaturedt > 0. && t > c->cfg.target_temperature)
         t = c->cfg.target_temperature;
      c->cfg.T_uniform = t;
   }
   const ampe_rhs_config& p = c->cfg;
   const Box& b = c->box;
   const int D = p.ndim, Q = p.qlen, ng = c->ng;
   int status = 0;

   // QuatIntegrator.cc:3189
   const bool recompute_quat_sidegrad = (fd_flag == 0) || !p.lag_quat_sidegrad;

   // setTemperatureField (ScalarTemperatureStrategy.cc:58-83): uniform T unless evolved
   if (!p.with_unsteady_temperature)
      for (auto& v : c->temp.data) v = p.T_uniform;

   // ---- setCoefficients (QuatIntegrator.cc:2994-3083) ----------------------
   // fillScratch
   if (p.with_phase) fill_periodic(b, y->phase, c->phase, ng, 1);
   if (Q > 0) fill_periodic(b, y->quat, c->quat, ng, Q);
   if (p.with_concentration) fill_periodic(b, y->conc, c->conc, ng, 1);
   if (p.with_unsteady_temperature) fill_periodic(b, y->temperature, c->temp, ng, 1);

   // computeQuatGradients (QuatIntegrator.cc:2782-2827)
   if (p.evolve_quat) {
      View diffs_symm[3], diffs_nonsymm[3], gside[3], gcell[3];
      views(c->quat_diffs, diffs_symm, D, 0);
      views(c->quat_diffs, diffs_nonsymm, D, p.symmetry_aware ? Q : 0);
      views(c->quat_grad_side, gside, D);
      for (int d = 0; d < D; d++) gcell[d] = c->quat_grad_cell[d].v;
      // computeQDiffs (computeQDiffs.cc:225-285)
      quatdiffs(b, Q, c->quat.v, diffs_nonsymm);
      if (p.symmetry_aware) quatdiffs_symm(b, Q, c->quat.v, diffs_symm, c->iqrot);
      if (p.symmetry_aware) {
         quatgrad_cell_symm(b, Q, p.dx, diffs_symm, gcell, c->iqrot);
         quatgrad_side_symm(b, Q, p.dx, diffs_symm, gside, c->iqrot);
      } else {
         quatgrad_cell(b, Q, p.dx, diffs_symm, gcell);
         quatgrad_side(b, Q, p.dx, diffs_symm, gside);
      }
      if (recompute_quat_sidegrad)
         for (int d = 0; d < D; d++)
            c->quat_grad_side_copy.a[d].data = c->quat_grad_side.a[d].data;
      if (p.quat_grad_modulus_from_cells)
         quatgrad_modulus(b, Q, gcell, c->quat_grad_modulus.v);
      else
         quatgrad_modulus_from_sides_compact(b, Q, gside, c->quat_grad_modulus.v);
   }

   // computePhaseConcentrations (QuatIntegrator.cc:3085-3125)
   if (p.with_concentration &&
       (p.conc_rhs_form == AMPE_CONC_KKS || p.conc_rhs_form == AMPE_CONC_EBS)) {
      if (compute_phase_concentrations(c) < 0) status = AMPE_ENEWTON;
   }

   // computeMobilities (QuatIntegrator.cc:2959-2990)
   if (p.with_phase)  // computeUniformPhaseMobility (QuatModel.cc:4277-4291)
      for (auto& v : c->phase_mobility.data) v = p.phi_mobility;
   if (p.evolve_quat)
      quatmobility(b, c->phase.v, c->quat_mobility.v, 1, p.quat_mobility,
                   p.min_quat_mobility, p.quat_mobility_func, p.quat_mobility_alt_scale);

   // ---- evaluatePhaseRHS -> PhaseRHSStrategyWithQ::evaluateRHS -------------
   if (p.with_phase) {
      View flux[3];
      views(c->phase_flux, flux, D);
      // PhaseFluxStrategy*::computeFluxes
      if (p.phase_flux_type == AMPE_FLUX_ANISOTROPIC)
         anisotropic_gradient_flux(b, p.dx, p.epsilon_phase, p.epsilon_anisotropy, p.knumber,
                                   c->phase.v, c->quat.v, Q, flux);
      else if (p.phase_flux_type == AMPE_FLUX_ISOTROPIC)
         compute_flux_isotropic(b, p.dx, p.epsilon_phase, c->phase.v, flux);
      else
         gradient_flux(b, p.dx, p.epsilon_phase, c->phase.v, flux);

      if (p.free_energy == AMPE_FE_CALPHAD || p.free_energy == AMPE_FE_QUADRATIC)
         compute_free_energies(c);  // computeFreeEnergyLiquid / SolidA

      // PhaseRHSStrategyWithQ.cc:243-263
      // (two-phase models: ptr_eta = NULL, eta_well_scale unused, three_phase = 0)
      computerhspbg(b, p.dx, 2.0 * p.H_parameter, p.epsilon_q, flux, c->temp.v,
                    p.phi_well_scale, 0.0, c->phase.v, View(), c->quat_grad_modulus.v,
                    c->rhs_phase.v, 'd', 'd', p.energy_interp, p.orient_interp1, p.orient_interp2,
                    p.evolve_quat ? 1 : 0, 0);

      // addDrivingForce
      if (p.free_energy == AMPE_FE_BIASWELL) {
         // ConstantMeltingTemperatureStrategy.cc:15-25 + BiasDoubleWellUTRC...:32-77
         for (auto& v : c->te.data) v = p.meltingT;
         computerhsbiaswell(b, c->phase.v, c->temp.v, p.bias_well_alpha, p.bias_well_gamma,
                            c->te.v, c->rhs_phase.v);
      } else if (p.free_energy == AMPE_FE_DELTAT) {
         // DeltaTemperatureFreeEnergyStrategy::addDrivingForce (DeltaTemperatureFreeEnergyStrategy.cc:96-153)
         computerhsdeltatemperature(b, c->phase.v, c->temp.v, p.meltingT, p.latent_heat, c->rhs_phase.v,
                                    p.energy_interp);
      } else if (p.free_energy == AMPE_FE_CALPHAD || p.free_energy == AMPE_FE_QUADRATIC) {
         add_driving_force(c);
      }
      // multiply by mobility (PhaseRHSStrategyWithQ.cc:297)
      ORACLE_PAR
      for (int k = b.lo[2]; k <= b.hi[2]; k++)
         for (int j = b.lo[1]; j <= b.hi[1]; j++)
            for (int i = b.lo[0]; i <= b.hi[0]; i++)
               c->rhs_phase.v(i, j, k) = c->rhs_phase.v(i, j, k) * c->phase_mobility.v(i, j, k);
      copy_out(b, c->rhs_phase, ydot->phase, 1);
   }

   // ---- evaluateQuatRHS (QuatIntegrator.cc:2691-2780) ----------------------
   if (p.evolve_quat) {
      View gside[3], gcopy[3], fc[3], f[3];
      views(c->quat_grad_side, gside, D);
      views(c->quat_grad_side_copy, gcopy, D);
      views(c->face_coef, fc, D);
      views(c->quat_flux, f, D);
      // QuatFACOps::evaluateRHS (QuatFACOps.cc:1861-1889): rhs = 0
      for (auto& v : c->rhs_quat.data) v = 0.0;
      compute_face_coef(b, Q, p.epsilon_q, c->phase.v, c->temp.v, 2. * p.H_parameter, gcopy,
                        fc, p.quat_grad_floor, p.grad_floor_type, p.orient_interp1,
                        p.orient_interp2, p.avg_func);
      // accumulateOperatorOnLevel with gq_id = grad_q_id (the literal `true`
      // at QuatIntegrator.cc:2709): flux = fc * normal side gradient
      compute_flux_from_gradq(b, Q, fc, gside, f);
      if (Q != 1) {
         // rotation_index_id = -1 (d_use_gradq_for_flux false): non-symm variants
         compute_lambda_flux(b, Q, f, c->quat.v, p.dx, c->lambda.v);
         add_quat_proj_op(b, Q, c->quat_mobility.v, f, c->quat.v, c->lambda.v, p.dx,
                          c->rhs_quat.v);
      } else {
         add_quat_op(b, Q, c->quat_mobility.v, f, p.dx, c->rhs_quat.v);
      }
      // correctRhsForSymmetry (QuatIntegrator.cc:2742-2745, 3967-4073)
      if (p.symmetry_aware) {
         View sd[3], nsd[3];
         views(c->quat_diffs, sd, D, 0);
         views(c->quat_diffs, nsd, D, Q);
         correctrhsquatforsymmetry(b, Q, p.dx, nsd, sd, c->rhs_quat.v, c->quat.v, fc,
                                   c->quat_mobility.v, c->iqrot);
      }
      copy_out(b, c->rhs_quat, ydot->quat, Q);
   }

   // ---- concentration ------------------------------------------------------
   if (p.with_concentration) {
      View cflux[3];
      views(c->conc_flux, cflux, D);
      if (p.conc_rhs_form == AMPE_CONC_CAHN_HILLIARD) {
         // CahnHilliardDoubleWell::computeFluxOnPatch (CahnHilliardDoubleWell.cc:67-109)
         for (int d = 0; d < D; d++)
            for (auto& v : c->conc_flux.a[d].data) v = 0.0;
         add_cahnhilliarddoublewell_flux(b, p.dx, c->conc.v, p.ch_mobility, p.ch_ca, p.ch_cb,
                                         p.ch_well_scale, p.ch_kappa, cflux);
      } else {
         if (recompute_quat_sidegrad) set_diffusion_coeff_for_concentration(c);
         compute_conc_flux_kks_ebs(c);
      }
      // QuatIntegrator.cc:2657-2668
      computerhsconcentration(b, p.dx, cflux, p.conc_mobility, c->rhs_conc.v);
      copy_out(b, c->rhs_conc, ydot->conc, 1);
   }

   // ---- temperature (SimpleTemperatureRHSStrategy.cc:31-95) ----------------
   if (p.with_unsteady_temperature) {
      for (auto& v : c->cp.data) v = p.cp;
      computerhstemp(b, p.dx, p.thermal_diffusivity, p.latent_heat, c->temp.v, c->cp.v,
                     p.with_phase ? 1 : 0, c->rhs_phase.v, c->rhs_temp.v);
      copy_out(b, c->rhs_temp, ydot->temperature, 1);
   }
   return status;
}

}  // namespace oracle
