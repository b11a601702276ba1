// Host-side mirror of QuatIntegrator's RHS methods (reference: source/QuatIntegrator.cc
// RegisterVariables 787-985, fillScratch 2873-2955, computeQuatGradients 2782-2827,
// setCoefficients 2994-3083, evaluateRHSFunction 3134-3295) on one uniform periodic level.
//
//   use_fused = true  : evaluateRHSFunction -> ampe_rhs_eval (the fused sm_100a path)
//   use_fused = false : the reference's own sequence of Strategy calls, each one a piecewise
//                       CUDA kernel on SAMRAI-layout PatchData (drop-in check of every Strategy)
#pragma once
#include <map>
#include <vector>

#include "ImplicitIntegrator.h"
#include "ampe_host.h"

namespace ampe_host {

class QuatIntegrator;

// Device vector backend of ImplicitIntegrator: the solution vector is an ampe_rhs_fields of device
// arrays; every operation is a C-ABI call into libampe_b200.so (the N_Vector operations CVODE would
// issue through Sundials_SAMRAIVector, samrai/Sundials_SAMRAIVector.cc).
class DeviceVectorOps
{
 public:
   typedef ampe_rhs_fields Vec;
   DeviceVectorOps(ampe_rhs_ctx* ctx, const ampe_rhs_config& cfg, QuatIntegrator* owner = nullptr)
       : d_ctx(ctx), d_cfg(cfg), d_owner(owner)
   {
      d_ncell = 1;
      for (int d = 0; d < cfg.ndim; d++) d_ncell *= (size_t)cfg.n[d];
      d_length = 0;
      if (cfg.with_phase) d_length += (long long)d_ncell;
      if (cfg.evolve_quat) d_length += (long long)d_ncell * cfg.qlen;
      if (cfg.with_concentration) d_length += (long long)d_ncell;
      if (cfg.with_unsteady_temperature) d_length += (long long)d_ncell;
   }
   Vec clone(const Vec& y)
   {
      Vec v = {nullptr, nullptr, nullptr, nullptr};
      dup(y.phase, &v.phase, 1);
      dup(y.quat, &v.quat, d_cfg.qlen);
      dup(y.conc, &v.conc, 1);
      dup(y.temperature, &v.temperature, 1);
      return v;
   }
   void release(Vec& v)
   {
      cudaFree(v.phase);
      cudaFree(v.quat);
      cudaFree(v.conc);
      cudaFree(v.temperature);
      v = Vec{nullptr, nullptr, nullptr, nullptr};
   }
   void linearSum(double a, const Vec& x, double b, const Vec& y, Vec& z)
   {
      check(ampe_vec_linear_sum(d_ctx, a, &x, b, &y, &z, nullptr), "linearSum");
   }
   void scale(double a, const Vec& x, Vec& z) { check(ampe_vec_scale(d_ctx, a, &x, &z, nullptr), "scale"); }
   // on slab ranks the sums run over every rank's cells: the owner's all-reduce hook is what AMPE's
   // Sundials_SAMRAIVector gets from SAMRAI_MPI::sumReduction (every rank sees the same bits, so every rank takes the
   // same step-size and convergence decisions)
   inline double wdot(const Vec& x, const Vec& y, const Vec& w);
   inline long long length() const;
   void errorWeights(const Vec& y, double rtol, double atol, Vec& w)
   {
      check(ampe_vec_error_weights(d_ctx, &y, rtol, atol, &w, nullptr), "errorWeights");
   }
   inline int rhs(double t, const Vec& y, Vec& ydot, int fd_flag);
   void applyProjection(double, const Vec& y, Vec& corr, Vec& err)
   {
      check(ampe_apply_projection(d_ctx, &y, &corr, &err, nullptr), "applyProjection");
   }
   // what QuatModel::Advance does after the integrator returns: normalizeQuat (QuatModel.cc:4222-4262)
   // and resetRefPhaseConcentrations (QuatModel.cc:5218-5231)
   inline void postStep(Vec& y);
   // CVSpgmrPrecondSet / CVSpgmrPrecondSolve of the owning QuatIntegrator (defined below the class)
   inline bool preconditioned() const;
   inline int precondSetup(double t, const Vec& y, double gamma);
   inline void precondSolve(const Vec& r, Vec& z);

 private:
   void dup(const double* src, double** dst, int depth)
   {
      if (!src || depth < 1) return;
      const size_t nb = d_ncell * (size_t)depth * sizeof(double);
      cuda_check(cudaMalloc(dst, nb), "clone");
      cuda_check(cudaMemcpy(*dst, src, nb, cudaMemcpyDeviceToDevice), "clone");
   }
   ampe_rhs_ctx* d_ctx;
   ampe_rhs_config d_cfg;
   QuatIntegrator* d_owner;
   size_t d_ncell;
   long long d_length;
   mutable long long d_global_length = -1;  // slab ranks: unknowns of all ranks (one reduction, then cached)
};

class QuatIntegrator
{
 public:
   QuatIntegrator(const ampe_rhs_config& cfg, bool use_fused) : d_cfg(cfg), d_use_fused(use_fused)
   {
      Box box;
      box.ndim = cfg.ndim;
      for (int d = 0; d < cfg.ndim; d++) box.upper[d] = cfg.n[d] - 1;
      d_hierarchy.reset(new PatchHierarchy);
      d_hierarchy->level.reset(new PatchLevel);
      d_patch.reset(new Patch(box, cfg.dx));
      d_hierarchy->level->patches.push_back(d_patch);
      d_ncell = 1;
      for (int d = 0; d < cfg.ndim; d++) d_ncell *= (size_t)cfg.n[d];
      check(ampe_rhs_create(&d_cfg, &d_ctx), "ampe_rhs_create");
      RegisterVariables();
      buildStrategies();
   }
   ~QuatIntegrator()
   {
      if (d_halo) ampe_halo_destroy(d_halo);
      ampe_rhs_destroy(d_ctx);
   }

   // ---- slab ranks (cfg.nranks > 1): the ghost planes along the slab axis come from the neighbours through the
   // C-ABI exchange (what fillScratch gets from RefineSchedule::fillData over MPI, QuatIntegrator.cc:2873-2955).
   // The application ships one handle per neighbour at set-up (MPI_Sendrecv in AMPE) and supplies the sum
   // reduction of the vector operations (SAMRAI_MPI::sumReduction).
   void haloExport(void* handle)
   {
      if (!d_halo) check(ampe_halo_create(d_ctx, d_cfg.rank, d_cfg.nranks, &d_halo), "ampe_halo_create");
      check(ampe_halo_export(d_halo, handle), "ampe_halo_export");
   }
   void haloConnect(const void* handle_prev, const void* handle_next)
   {
      if (!d_halo) throw std::runtime_error("haloConnect: haloExport first");
      check(ampe_halo_connect(d_halo, handle_prev, handle_next), "ampe_halo_connect");
   }
   typedef double (*SumReduction)(double local, void* user);
   void setSumReduction(SumReduction fn, void* user) { d_sum_reduction = fn, d_sum_user = user; }
   ampe_halo* halo() const { return d_halo; }
   double sumReduction(double v) const
   {
      if (d_cfg.nranks > 1 && !d_sum_reduction) throw std::runtime_error("slab ranks: setSumReduction first");
      return d_sum_reduction ? d_sum_reduction(v, d_sum_user) : v;
   }

   // QuatModel::resetRefPhaseConcentrations (QuatModel.cc:5218-5231); ghost-0 device arrays
   void resetRefPhaseConcentrations(const double* cl_ref, const double* ca_ref)
   {
      if (d_halo) {  // slab rank: the ghost planes of the reference state come from the neighbours (collective)
         check(ampe_rhs_set_ref_concentrations_slab(d_ctx, d_halo, cl_ref, ca_ref, nullptr), "set_ref_slab");
         return;
      }
      check(ampe_rhs_set_ref_concentrations(d_ctx, cl_ref, ca_ref, nullptr), "set_ref");
      if (d_conc_l_ref_id < 0) return;
      const Box& b = d_patch->getBox();
      if (cl_ref && ca_ref) {
         check(ampe_k_fill_periodic(b.ndim, b.lower, b.upper, 1, cl_ref,
                                    d_patch->cell<double>(d_conc_l_ref_id)->getPointer(), d_ng, nullptr),
               "fill ref");
         check(ampe_k_fill_periodic(b.ndim, b.lower, b.upper, 1, ca_ref,
                                    d_patch->cell<double>(d_conc_a_ref_id)->getPointer(), d_ng, nullptr),
               "fill ref");
      } else {
         auto cl = d_patch->cell<double>(d_conc_l_id), ca = d_patch->cell<double>(d_conc_a_id);
         cudaMemcpy(d_patch->cell<double>(d_conc_l_ref_id)->getPointer(), cl->getPointer(),
                    cl->size() * sizeof(double), cudaMemcpyDeviceToDevice);
         cudaMemcpy(d_patch->cell<double>(d_conc_a_ref_id)->getPointer(), ca->getPointer(),
                    ca->size() * sizeof(double), cudaMemcpyDeviceToDevice);
      }
   }
   // quat_symm_rotation: one int per lower face per direction, ghost-0 device arrays
   void setSymmetryRotations(const int* const* iqrot)
   {
      check(ampe_rhs_set_symmetry_rotations(d_ctx, iqrot, nullptr), "set_rotations");
      const Box& b = d_patch->getBox();
      auto rot = d_patch->side<int>(d_quat_symm_rotation_id);
      for (int a = 0; a < b.ndim; a++)
         check(ampe_k_fill_periodic_int(b.ndim, b.lower, b.upper, a, iqrot[a], rot->getPointer(a), 1,
                                        nullptr),
               "fill rotations");
   }

   // QuatIntegrator::evaluateRHSFunction(time, y, y_dot, fd_flag); y / y_dot: ghost-0 device arrays
   int evaluateRHSFunction(double time, const ampe_rhs_fields* y, const ampe_rhs_fields* y_dot,
                           int fd_flag)
   {
      if (d_use_fused) {
         if (d_halo)
            check(ampe_rhs_eval_slab(d_ctx, d_halo, time, y, y_dot, fd_flag, nullptr), "ampe_rhs_eval_slab");
         else
            check(ampe_rhs_eval(d_ctx, time, y, y_dot, fd_flag, nullptr), "ampe_rhs_eval");
         return 0;
      }
      if (d_cfg.nranks > 1) throw std::runtime_error("the unfused Strategy path runs on one rank");
      const ampe_rhs_config& p = d_cfg;
      const bool recompute_quat_sidegrad = (fd_flag == 0) || !p.lag_quat_sidegrad;  // :3189
      setCoefficients(time, y, recompute_quat_sidegrad);
      const Box& b = d_patch->getBox();
      if (p.with_phase) {
         d_phase_rhs_strategy->evaluateRHS(time, d_hierarchy, d_ydot_phase_id, fd_flag == 0);
         copyOut(d_ydot_phase_id, y_dot->phase, 1);
      }
      if (p.evolve_quat) {
         // evaluateQuatRHS (:2691-2780): the literal `true` = use_gradq_for_flux
         d_quat_sys_solver->evaluateRHS(d_phase_scratch_id, d_temperature_scratch_id,
                                        d_quat_grad_side_id, d_quat_grad_side_copy_id, -1,
                                        d_quat_mobility_id, d_quat_scratch_id, d_ydot_quat_id, true);
         if (p.symmetry_aware) correctRhsForSymmetry();
         copyOut(d_ydot_quat_id, y_dot->quat, p.qlen);
      }
      if (p.with_concentration) {
         if (recompute_quat_sidegrad) d_composition_rhs_strategy->setDiffusionCoeff(d_hierarchy, time);
         d_composition_rhs_strategy->computeFluxOnPatch(*d_patch, d_flux_conc_id);
         auto flux = d_patch->side<double>(d_flux_conc_id);
         auto f = flux->pointers(0);
         auto rhs = d_patch->cell<double>(d_ydot_conc_id);
         check(ampe_k_computerhsconcentration(b.ndim, b.lower, b.upper, d_patch->getDx(), f.data(),
                                              flux->getGhostCellWidth(), p.conc_mobility,
                                              rhs->getPointer(), 0, nullptr),
               "COMPUTERHSCONCENTRATION");
         copyOut(d_ydot_conc_id, y_dot->conc, 1);
      }
      if (p.with_unsteady_temperature) {
         d_temperature_rhs_strategy->evaluateRHS(d_hierarchy, d_ydot_temperature_id,
                                                 p.with_phase ? d_ydot_phase_id : -1);
         copyOut(d_ydot_temperature_id, y_dot->temperature, 1);
      }
      cuda_check(cudaDeviceSynchronize(), "evaluateRHSFunction");
      return 0;  // Always successful (QuatIntegrator.cc:3294)
   }

   // ---- what sits around the RHS in the CVODE loop, device resident --------------
   // Sundials_SAMRAIVector operations on the evolved components (samrai/Sundials_SAMRAIVector.cc)
   void linearSum(double a, const ampe_rhs_fields* x, double b, const ampe_rhs_fields* y,
                  const ampe_rhs_fields* z)
   {
      check(ampe_vec_linear_sum(d_ctx, a, x, b, y, z, nullptr), "linearSum");
   }
   void scale(double a, const ampe_rhs_fields* x, const ampe_rhs_fields* z)
   {
      check(ampe_vec_scale(d_ctx, a, x, z, nullptr), "scale");
   }
   double dotWith(const ampe_rhs_fields* x, const ampe_rhs_fields* y)
   {
      double r = 0.0;
      check(ampe_vec_dot(d_ctx, x, y, &r, nullptr), "dotWith");
      return r;
   }
   double weightedRMSNorm(const ampe_rhs_fields* x, const ampe_rhs_fields* w)
   {
      double r = 0.0;
      check(ampe_vec_wrms_norm(d_ctx, x, w, &r, nullptr), "weightedRMSNorm");
      return r;
   }
   double maxNorm(const ampe_rhs_fields* x)
   {
      double r = 0.0;
      check(ampe_vec_max_norm(d_ctx, x, &r, nullptr), "maxNorm");
      return r;
   }
   // QuatModel::normalizeQuat (QuatModel.cc:4222-4262)
   void normalizeQuat(const ampe_rhs_fields* y) { check(ampe_normalize_quat(d_ctx, y, nullptr), "normalizeQuat"); }
   // QuatModel::evaluateEnergy (QuatModel.cc:4888-4976): total, phase, orient, qint, well, free
   void evaluateEnergy(const ampe_rhs_fields* y, double& total_energy, double& total_phase_e,
                       double& total_orient_e, double& total_qint_e, double& total_well_e,
                       double& total_free_e)
   {
      double e[8];
      check(ampe_energy_eval(d_ctx, y, e, nullptr), "evaluateEnergy");
      total_energy = e[0], total_phase_e = e[1], total_orient_e = e[2], total_qint_e = e[3];
      total_well_e = e[4], total_free_e = e[5];
   }
   // QuatIntegrator::applyProjection(time, y, corr, epsProj, err) (QuatIntegrator.cc:3911-3962): CVODE's
   // projection hook; "Always successful" like the reference.
   int applyProjection(double time, const ampe_rhs_fields* y, const ampe_rhs_fields* corr, double epsProj,
                       const ampe_rhs_fields* err)
   {
      (void)time;
      (void)epsProj;
      check(ampe_apply_projection(d_ctx, y, corr, err, nullptr), "applyProjection");
      return 0;
   }
   // QuatModel::computeSymmetryRotations (QuatModel.cc:4978-5055) and makeQuatFundamental (:5059-5104)
   void computeSymmetryRotations(const ampe_rhs_fields* y)
   {
      check(ampe_rhs_compute_symmetry_rotations(d_ctx, y, nullptr), "computeSymmetryRotations");
   }
   // QuatModel::computeGrainDiagnostics (QuatModel.cc:2690-2705): findAndNumberGrains + computeGrainVolumes; the map
   // the reference prints as "Volume of grain N = V" (Grains.cc:693-697)
   std::map<int, double> computeGrainDiagnostics(const ampe_rhs_fields* y, double phase_threshold = 0.85,
                                                 int max_grains = 65536)
   {
      std::vector<int> ids(max_grains);
      std::vector<double> vol(max_grains);
      int n = 0;
      check(ampe_grain_volumes(d_ctx, y, phase_threshold, max_grains, &n, ids.data(), vol.data(), nullptr),
            "computeGrainDiagnostics");
      std::map<int, double> out;
      for (int i = 0; i < n; i++) out[ids[i]] = vol[i];
      return out;
   }
   void makeQuatFundamental(const ampe_rhs_fields* y)
   {
      check(ampe_quat_fundamental(d_ctx, y, nullptr), "makeQuatFundamental");
   }
   // nsteps fixed BDF steps (ImplicitIntegrator.h: Newton + matrix-free GMRES, fd_flag = 1 Jacobian-vector
   // products, projection): the CVODE-shaped stand-in for QuatIntegrator::Advance, y stays on the device
   int integrateImplicit(const ampe_rhs_fields* y, double t0, double dt, int nsteps, const ImplicitOptions& opt,
                         ImplicitStats* stats)
   {
      DeviceVectorOps ops(d_ctx, d_cfg, this);
      ImplicitOptions o = opt;
      o.precondition_left = d_precond_left;
      ImplicitIntegrator<DeviceVectorOps> integ(ops, o);
      ampe_rhs_fields yy = *y;
      const int rc = integ.advance(yy, t0, dt, nsteps);
      if (stats) *stats = integ.stats();
      cuda_check(cudaDeviceSynchronize(), "integrateImplicit");
      return rc;
   }
   // variable steps with the local error test from t0 to tend (ImplicitIntegrator::advanceTo): what
   // QuatIntegrator::Advance asks of CVode(tend, CV_NORMAL), y stays on the device
   int integrateAdaptive(const ampe_rhs_fields* y, double t0, double tend, double h0, const ImplicitOptions& opt,
                         ImplicitStats* stats)
   {
      DeviceVectorOps ops(d_ctx, d_cfg, this);
      ImplicitOptions o = opt;
      o.precondition_left = d_precond_left;
      ImplicitIntegrator<DeviceVectorOps> integ(ops, o);
      ampe_rhs_fields yy = *y;
      const int rc = integ.advanceTo(yy, t0, tend, h0);
      if (stats) *stats = integ.stats();
      cuda_check(cudaDeviceSynchronize(), "integrateAdaptive");
      return rc;
   }
   // fixed-step explicit stand-in for QuatIntegrator::Advance (scheme 0 Euler, 1 Heun)
   void integrateFixed(const ampe_rhs_fields* y, const ampe_rhs_fields* work1, const ampe_rhs_fields* work2,
                       double t0, double dt, int nsteps, int scheme)
   {
      check(ampe_integrate_fixed(d_ctx, y, work1, work2, t0, dt, nsteps, scheme, nullptr), "integrateFixed");
   }

   // ---- block preconditioners ----------------------------------------------
   // QuatIntegrator::setupPreconditioners (QuatIntegrator.cc:428-545) with the Preconditioner{} block's
   // defaults except precond_has_dquatdphi = false (block diagonal).  ncycles = V-cycles per block
   // solve (the reference iterates FAC cycles to CVODE's delta; a fixed count keeps the
   // preconditioner a fixed linear operator); 0 switches the preconditioner off.
   void setupPreconditioners(int ncycles, bool precond_has_dquatdphi = false, bool precondition_left = false)
   {
      const ampe_rhs_config& p = d_cfg;
      d_precond_cycles = ncycles;
      d_precond_left = precondition_left;  // PREC_LEFT (QuatIntegrator.cc:1583) instead of right preconditioning
      d_precond_has_dquatdphi = precond_has_dquatdphi && p.with_phase && p.evolve_quat;  // :485
      d_use_preconditioner = ncycles > 0;
      if (!d_use_preconditioner) return;
      // Slab ranks: block Jacobi over the ranks.  Every rank runs its block solvers on its own slab with no flux
      // through the faces it shares with its neighbours (the zero-slope treatment of a physical boundary), so a
      // solve needs no communication; GMRES sees the coupling between the slabs through the Jacobian-vector
      // products.  The reference's FAC / hypre solvers span the ranks; this changes the Krylov iteration count, not
      // the converged Newton step.
      int zs[3] = {p.zero_slope[0], p.zero_slope[1], p.zero_slope[2]};
      if (p.nranks > 1) zs[p.ndim - 1] = 1;
      if (p.with_phase && !d_phase_sys_solver) {
         d_phase_precond_c_id = cellVar<double>(1, 0);
         d_phase_sys_solver.reset(new PhaseFACSolver(d_hierarchy, d_phase_precond_c_id));
         d_phase_sys_solver->setBoundaries(zs);
      }
      const bool kks = p.conc_rhs_form == AMPE_CONC_KKS || p.conc_rhs_form == AMPE_CONC_EBS;
      if (p.with_concentration && kks && !d_conc_sys_solver) {
         d_conc_sys_solver.reset(new ConcFACSolver(d_hierarchy));
         d_conc_sys_solver->setBoundaries(zs);
         d_conc_l_g0_id = cellVar<double>(1, 0);
         d_conc_a_g0_id = cellVar<double>(1, 0);
      }
      if (p.with_unsteady_temperature && !d_temperature_sys_solver) {
         d_temperature_sys_solver.reset(new TemperatureFACSolver(d_hierarchy));
         d_temperature_sys_solver->setBoundaries(zs);
      }
      if (d_precond_has_dquatdphi && !d_diffusion4quatderiv) {
         // RegisterVariables with d_precond_has_dquatdphi (QuatIntegrator.cc:1170-1190) and the
         // solver-owned scratch of QuatFACOps (d_sqrt_m_id, d_face_coef_scratch_id)
         d_quat_mobility_deriv_id = cellVar<double>(1, 0);
         d_quat_diffusion_deriv_id = sideVar<double>(2, 0);
         d_phase_sol_id = cellVar<double>(1, 1);
         d_quat_rhs_id = cellVar<double>(p.qlen, 0);
         d_sqrt_m_id = cellVar<double>(1, 1);
         d_face_coef_scratch_id = sideVar<double>(1, 0);
         d_diffusion4quatderiv.reset(new DerivDiffusionCoeffForQuat(p, d_quat_diffusion_deriv_id));
      }
   }
   bool usePreconditioner() const { return d_use_preconditioner; }
   // QuatIntegrator::CVSpgmrPrecondSet(t, y, fy, jok, jcurPtr, gamma) (QuatIntegrator.cc:3300-3376)
   int CVSpgmrPrecondSet(double t, const ampe_rhs_fields* y, double gamma)
   {
      if (!d_use_preconditioner) throw std::runtime_error("CVSpgmrPrecondSet: call setupPreconditioners first");
      const ampe_rhs_config& p = d_cfg;
      const bool kks = p.conc_rhs_form == AMPE_CONC_KKS || p.conc_rhs_form == AMPE_CONC_EBS;
      if (d_use_fused && d_phase_conc_strategy) {
         // the fused evaluation at this y (the fd_flag = 0 residual evaluation that precedes every
         // set-up) has already solved the per-cell KKS problem: take c_l, c_a from the context
         // instead of repeating the Newton solve
         setCoefficients(t, y, true, false);
         auto cl0 = d_patch->cell<double>(d_conc_l_g0_id), ca0 = d_patch->cell<double>(d_conc_a_g0_id);
         check(ampe_rhs_copy_phase_concentrations(d_ctx, cl0->getPointer(), ca0->getPointer(), nullptr),
               "copy_phase_concentrations");
         fillScratchField(cl0->getPointer(), d_conc_l_id, 1);
         fillScratchField(ca0->getPointer(), d_conc_a_id, 1);
      } else {
         setCoefficients(t, y, true);
      }
      if (p.with_phase)
         d_phase_sys_solver->setOperatorCoefficients(d_phase_scratch_id, d_phase_mobility_id, p.epsilon_phase, gamma,
                                                     p.phi_well_scale, "double", &p.phi_mobility);
      if (p.with_unsteady_temperature)
         d_temperature_sys_solver->setOperatorCoefficients(1., 1., -gamma * p.thermal_diffusivity);
      if (p.with_concentration && kks) {
         // setCompositionOperatorCoefficients (:3378-3387): D_pfm = D_l + D_a (EBS,
         // setDiffusionCoeffForPreconditioner) or D0 (KKS)
         d_composition_rhs_strategy->setDiffusionCoeff(d_hierarchy, t);
         d_conc_sys_solver->setOperatorCoefficients(gamma, d_diff0_id,
                                                    p.conc_rhs_form == AMPE_CONC_EBS ? d_diff1_id : -1,
                                                    p.conc_mobility);
      }
      if (d_precond_has_dquatdphi) {
         // setCoefficients with d_precond_has_dquatdphi (:2978-2983, :3064-3070)
         d_mobility_strategy->computeQuatMobilityDeriv(d_hierarchy, d_phase_scratch_id, d_quat_mobility_deriv_id);
         d_diffusion4quatderiv->setDerivDiffusion(d_hierarchy, d_phase_scratch_id, d_temperature_scratch_id,
                                                  d_quat_grad_side_copy_id);
      }
      if (p.evolve_quat)
         d_quat_sys_solver->setOperatorCoefficients(gamma, d_quat_mobility_id,
                                                    d_precond_has_dquatdphi ? d_quat_mobility_deriv_id : -1,
                                                    d_phase_scratch_id, d_temperature_scratch_id,
                                                    d_precond_has_dquatdphi ? d_quat_diffusion_deriv_id : -1,
                                                    d_quat_grad_side_copy_id, d_quat_scratch_id);
      d_precond_gamma = gamma;
      d_precond_setups++;
      return 0;
   }
   // QuatIntegrator::CVSpgmrPrecondSolve(t, y, fy, r, z, gamma, delta, lr) (QuatIntegrator.cc:3666-3771):
   // z_block = A_block^-1 r_block, block by block; r, z: ghost-0 device vectors
   int CVSpgmrPrecondSolve(const ampe_rhs_fields* r, const ampe_rhs_fields* z)
   {
      if (!d_use_preconditioner) throw std::runtime_error("CVSpgmrPrecondSolve: call setupPreconditioners first");
      const ampe_rhs_config& p = d_cfg;
      if (p.with_phase) d_phase_sys_solver->solveSystem(z->phase, r->phase, d_precond_cycles);
      if (p.evolve_quat) {
         const double* r_quat = r->quat;
         if (d_precond_has_dquatdphi) {
            // QuatPrecondSolve (:3602-3612): quat_rhs = r_quat + gamma [dF_q/dphi] phase_sol
            fillScratchField(z->phase, d_phase_sol_id, 1);
            d_quat_sys_solver->multiplyDQuatDPhiBlock(d_phase_sol_id, d_quat_rhs_id, d_sqrt_m_id,
                                                      d_face_coef_scratch_id);
            auto qr = d_patch->cell<double>(d_quat_rhs_id);
            const Box& b = d_patch->getBox();
            check(ampe_k_cell_axpy(b.ndim, b.lower, b.upper, p.qlen, d_precond_gamma, qr->getPointer(), 0, r->quat, 0,
                                   qr->getPointer(), 0, nullptr),
                  "axpy");
            r_quat = qr->getPointer();
         }
         d_quat_sys_solver->solveSystem(z->quat, r_quat, d_precond_cycles);
      }
      if (p.with_unsteady_temperature)
         d_temperature_sys_solver->solveSystem(z->temperature, r->temperature, d_precond_cycles);
      if (p.with_concentration) {
         if (d_conc_sys_solver)
            d_conc_sys_solver->solveSystem(z->conc, r->conc, d_precond_cycles);
         else if (z->conc != r->conc)  // Cahn-Hilliard: no block solver, identity
            cuda_check(cudaMemcpy(z->conc, r->conc, d_ncell * sizeof(double), cudaMemcpyDeviceToDevice),
                       "CVSpgmrPrecondSolve");
      }
      d_precond_solves++;
      return 0;
   }
   // the device multigrid of a block (0 phase, 1 quaternion, 2 composition, 3 temperature), or NULL
   ampe_mg* preconditionerLevelSolver(int block) const
   {
      if (block == 0) return d_phase_sys_solver ? d_phase_sys_solver->levelSolver() : nullptr;
      if (block == 1) return d_quat_sys_solver ? d_quat_sys_solver->levelSolver() : nullptr;
      if (block == 2) return d_conc_sys_solver ? d_conc_sys_solver->levelSolver() : nullptr;
      if (block == 3) return d_temperature_sys_solver ? d_temperature_sys_solver->levelSolver() : nullptr;
      return nullptr;
   }
   // QuatSysSolver::multiplyDQuatDPhiBlock on a ghost-0 phase vector; out: ghost-0 device array of depth qlen
   void multiplyDQuatDPhiBlock(const double* phase, double* out)
   {
      if (!d_precond_has_dquatdphi) throw std::runtime_error("multiplyDQuatDPhiBlock: coupling block not set up");
      fillScratchField(phase, d_phase_sol_id, 1);
      d_quat_sys_solver->multiplyDQuatDPhiBlock(d_phase_sol_id, d_quat_rhs_id, d_sqrt_m_id, d_face_coef_scratch_id);
      copyOut(d_quat_rhs_id, out, d_cfg.qlen);
   }
   long precondSetups() const { return d_precond_setups; }
   long precondSolves() const { return d_precond_solves; }

   std::shared_ptr<Patch> patch() const { return d_patch; }
   ampe_rhs_ctx* fusedContext() const { return d_ctx; }
   int concLId() const { return d_conc_l_id; }
   int concAId() const { return d_conc_a_id; }
   int nghosts() const { return d_ng; }

 private:
   template <typename T>
   int cellVar(int depth, int ghosts)
   {
      return d_patch->registerPatchData(std::make_shared<CellData<T>>(d_patch->getBox(), depth, ghosts));
   }
   template <typename T>
   int sideVar(int depth, int ghosts)
   {
      return d_patch->registerPatchData(std::make_shared<SideData<T>>(d_patch->getBox(), depth, ghosts));
   }

   // ghost widths as registered by the reference (SURVEY.md 8b "Array layout")
   void RegisterVariables()
   {
      const ampe_rhs_config& p = d_cfg;
      const int Q = p.qlen, D = p.ndim;
      d_ng = (p.conc_rhs_form == AMPE_CONC_CAHN_HILLIARD) ? 2 : 1;  // nghosts_required()
      d_temperature_scratch_id = cellVar<double>(1, d_ng);
      if (p.with_phase) {
         d_phase_scratch_id = cellVar<double>(1, d_ng);
         d_phase_mobility_id = cellVar<double>(1, 1);
         d_flux_id = sideVar<double>(1, 0);
         d_ydot_phase_id = cellVar<double>(1, 0);
      }
      if (Q > 0) d_quat_scratch_id = cellVar<double>(Q, d_ng);
      if (p.evolve_quat) {
         d_quat_diffs_id = sideVar<double>(p.symmetry_aware ? 2 * Q : Q, 1);
         d_quat_grad_cell_id = cellVar<double>(D * Q, 0);
         d_quat_grad_side_id = sideVar<double>(D * Q, 0);
         d_quat_grad_side_copy_id = sideVar<double>(D * Q, 0);
         d_quat_grad_modulus_id = cellVar<double>(1, 0);
         d_quat_mobility_id = cellVar<double>(1, 1);
         d_face_coef_id = sideVar<double>(1, 0);
         d_quat_flux_id = sideVar<double>(Q, 0);
         d_lambda_id = cellVar<double>(1, 0);
         d_ydot_quat_id = cellVar<double>(Q, 0);
         if (p.symmetry_aware) d_quat_symm_rotation_id = sideVar<int>(1, 1);
      }
      if (p.with_concentration) {
         d_conc_scratch_id = cellVar<double>(1, d_ng);
         d_flux_conc_id = sideVar<double>(1, p.conc_rhs_form == AMPE_CONC_CAHN_HILLIARD ? 1 : 0);
         d_ydot_conc_id = cellVar<double>(1, 0);
         if (p.conc_rhs_form == AMPE_CONC_KKS || p.conc_rhs_form == AMPE_CONC_EBS) {
            d_conc_l_id = cellVar<double>(1, d_ng);
            d_conc_a_id = cellVar<double>(1, d_ng);
            d_conc_l_ref_id = cellVar<double>(1, d_ng);
            d_conc_a_ref_id = cellVar<double>(1, d_ng);
            d_f_l_id = cellVar<double>(1, 0);
            d_f_a_id = cellVar<double>(1, 0);
            d_diff0_id = sideVar<double>(1, 0);  // EBS: D_l  | KKS: D0
            d_diff1_id = sideVar<double>(1, 0);  // EBS: D_a  | KKS: D_phi
         }
      }
      if (p.with_unsteady_temperature) {
         d_cp_id = cellVar<double>(1, 0);
         d_ydot_temperature_id = cellVar<double>(1, 0);
         fillConstant(d_cp_id, p.cp);
      }
      if (p.free_energy == AMPE_FE_BIASWELL) {
         d_eq_temperature_id = cellVar<double>(1, 0);
         fillConstant(d_eq_temperature_id, p.meltingT);  // ConstantMeltingTemperatureStrategy
      }
   }

   void buildStrategies()
   {
      const ampe_rhs_config& p = d_cfg;
      if (p.with_phase) {
         // PhaseFluxStrategyFactory.h:11-36
         if (p.phase_flux_type == AMPE_FLUX_ANISOTROPIC)
            d_phase_flux_strategy.reset(
                new PhaseFluxStrategyAnisotropy(p.epsilon_phase, p.epsilon_anisotropy, p.knumber));
         else if (p.phase_flux_type == AMPE_FLUX_ISOTROPIC)
            d_phase_flux_strategy.reset(new PhaseFluxStrategyIsotropic(p.epsilon_phase));
         else
            d_phase_flux_strategy.reset(new PhaseFluxStrategySimple(p.epsilon_phase));
         // FreeEnergyStrategyFactory.h:37-263
         if (p.free_energy == AMPE_FE_BIASWELL)
            d_free_energy_strategy.reset(new BiasDoubleWellUTRCFreeEnergyStrategy(
                p.bias_well_alpha, p.bias_well_gamma, d_eq_temperature_id));
         else if (p.free_energy == AMPE_FE_CALPHAD || p.free_energy == AMPE_FE_QUADRATIC)
            d_free_energy_strategy.reset(new KKSFreeEnergyStrategy(p, d_conc_l_id, d_conc_a_id));
         d_phase_rhs_strategy.reset(new PhaseRHSStrategyWithQ(
             p, d_phase_scratch_id, d_conc_scratch_id, d_quat_scratch_id, d_temperature_scratch_id,
             d_f_l_id, d_f_a_id, d_phase_mobility_id, d_flux_id, d_quat_grad_modulus_id,
             d_phase_flux_strategy, d_free_energy_strategy));
      }
      d_mobility_strategy.reset(new QuatMobilityStrategy(p));
      if (p.evolve_quat) {
         d_quat_grad_strategy.reset(
             new QuatGradStrategy(p.qlen, p.symmetry_aware != 0, d_quat_symm_rotation_id));
         d_quat_face_coeff.reset(new QuatFaceCoeff(p));
         d_quat_sys_solver.reset(new QuatSysSolver(p, d_hierarchy, d_quat_face_coeff, d_face_coef_id,
                                                   d_quat_flux_id, d_lambda_id));
      }
      if (p.with_concentration) {
         // CompositionRHSStrategyFactory.h:27-102
         if (p.conc_rhs_form == AMPE_CONC_CAHN_HILLIARD)
            d_composition_rhs_strategy.reset(new CahnHilliardDoubleWell(p, d_conc_scratch_id));
         else if (p.conc_rhs_form == AMPE_CONC_EBS)
            d_composition_rhs_strategy.reset(new EBSCompositionRHSStrategy(
                p, d_phase_scratch_id, d_conc_l_id, d_conc_a_id, d_diff0_id, d_diff1_id));
         else
            d_composition_rhs_strategy.reset(new KKSCompositionRHSStrategy(
                p, d_conc_scratch_id, d_phase_scratch_id, d_temperature_scratch_id, d_conc_l_id,
                d_conc_a_id, d_diff0_id, d_diff1_id));
         if (p.conc_rhs_form == AMPE_CONC_KKS || p.conc_rhs_form == AMPE_CONC_EBS)
            d_phase_conc_strategy.reset(new PhaseConcentrationsStrategy(
                p, d_conc_l_id, d_conc_a_id, d_conc_l_ref_id, d_conc_a_ref_id));
      }
      if (p.with_unsteady_temperature)
         d_temperature_rhs_strategy.reset(new TemperatureRHSStrategy(
             d_temperature_scratch_id, d_cp_id, p.thermal_diffusivity, p.latent_heat));
   }

   void fillConstant(int id, double v)
   {
      auto c = d_patch->cell<double>(id);
      std::vector<double> h(c->size(), v);
      cuda_check(cudaMemcpy(c->getPointer(), h.data(), h.size() * sizeof(double),
                            cudaMemcpyHostToDevice),
                 "fillConstant");
   }
   // fillScratch (:2873-2955): y -> scratch, ghosts = periodic images
   void fillScratchField(const double* src, int id, int depth)
   {
      const Box& b = d_patch->getBox();
      auto c = d_patch->cell<double>(id);
      check(ampe_k_fill_periodic(b.ndim, b.lower, b.upper, depth, src, c->getPointer(),
                                 c->getGhostCellWidth(), nullptr),
            "fillScratch");
   }
   void copyOut(int id, double* dst, int depth)
   {
      auto c = d_patch->cell<double>(id);
      cuda_check(cudaMemcpy(dst, c->getPointer(), d_ncell * depth * sizeof(double),
                            cudaMemcpyDeviceToDevice),
                 "copyOut");
   }

   // setCoefficients (:2994-3083)
   void setCoefficients(double time, const ampe_rhs_fields* y, bool recompute_quat_sidegrad,
                        bool solve_phase_concentrations = true)
   {
      (void)time;
      const ampe_rhs_config& p = d_cfg;
      if (!p.with_unsteady_temperature && !d_uniform_temperature_filled) {
         // setTemperatureField with a uniform T: the scratch array never changes, fill it once (the
         // preconditioner set-up comes through here at every Newton iteration)
         fillConstant(d_temperature_scratch_id, p.T_uniform);
         d_uniform_temperature_filled = true;
      }
      if (p.with_phase) fillScratchField(y->phase, d_phase_scratch_id, 1);
      if (p.qlen > 0) fillScratchField(y->quat, d_quat_scratch_id, p.qlen);
      if (p.with_concentration) fillScratchField(y->conc, d_conc_scratch_id, 1);
      if (p.with_unsteady_temperature) fillScratchField(y->temperature, d_temperature_scratch_id, 1);
      if (p.evolve_quat) computeQuatGradients(recompute_quat_sidegrad);
      if (d_phase_conc_strategy && solve_phase_concentrations) {
         int nfail = d_phase_conc_strategy->computePhaseConcentrations(
             d_hierarchy, d_temperature_scratch_id, d_phase_scratch_id, -1, d_conc_scratch_id);
         if (nfail > 0) throw std::runtime_error("computePhaseConcentrations: Newton failed");
      }
      // computeMobilities (:2959-2990)
      if (p.with_phase)
         d_mobility_strategy->computePhaseMobility(d_hierarchy, d_phase_scratch_id, d_phase_mobility_id);
      if (p.evolve_quat)
         d_mobility_strategy->computeQuatMobility(d_hierarchy, d_phase_scratch_id, d_quat_mobility_id);
   }
   // computeQuatGradients (:2782-2827)
   void computeQuatGradients(bool recompute_quat_sidegrad)
   {
      d_quat_grad_strategy->computeDiffs(d_hierarchy, d_quat_scratch_id, d_quat_diffs_id);
      d_quat_grad_strategy->computeGradCell(d_hierarchy, d_quat_diffs_id, d_quat_grad_cell_id);
      d_quat_grad_strategy->computeGradSide(d_hierarchy, d_quat_diffs_id, d_quat_grad_side_id);
      if (recompute_quat_sidegrad)
         d_patch->side<double>(d_quat_grad_side_copy_id)->copy(*d_patch->side<double>(d_quat_grad_side_id));
      if (d_cfg.quat_grad_modulus_from_cells)
         d_quat_grad_strategy->computeGradModulus(d_hierarchy, d_quat_grad_cell_id,
                                                  d_quat_grad_modulus_id);
      else
         d_quat_grad_strategy->computeGradModulusFromSides(d_hierarchy, d_quat_grad_side_id,
                                                           d_quat_grad_modulus_id);
   }
   // correctRhsForSymmetry (:3967-4073)
   void correctRhsForSymmetry()
   {
      const Box& b = d_patch->getBox();
      const int Q = d_cfg.qlen;
      auto diffs = d_patch->side<double>(d_quat_diffs_id);
      auto nonsymm = diffs->pointers(Q), symm = diffs->pointers(0);
      auto rhs = d_patch->cell<double>(d_ydot_quat_id), q = d_patch->cell<double>(d_quat_scratch_id);
      auto fc = d_patch->side<double>(d_quat_sys_solver->getFaceDiffCoeffScratchId());
      auto f = fc->pointers(0);
      auto mob = d_patch->cell<double>(d_quat_mobility_id);
      auto rot = d_patch->side<int>(d_quat_symm_rotation_id);
      auto iq = rot->pointers();
      std::vector<const int*> ciq(iq.begin(), iq.end());
      check(ampe_k_correctrhsquatforsymmetry(b.ndim, b.lower, b.upper, Q, d_patch->getDx(),
                                             nonsymm.data(), symm.data(), diffs->getGhostCellWidth(),
                                             rhs->getPointer(), rhs->getGhostCellWidth(),
                                             q->getPointer(), q->getGhostCellWidth(), f.data(),
                                             fc->getGhostCellWidth(), mob->getPointer(),
                                             mob->getGhostCellWidth(), ciq.data(),
                                             rot->getGhostCellWidth(), nullptr),
            "CORRECTRHSQUATFORSYMMETRY");
   }

   ampe_rhs_config d_cfg;
   bool d_use_fused;
   ampe_halo* d_halo = nullptr;
   SumReduction d_sum_reduction = nullptr;
   void* d_sum_user = nullptr;
   ampe_rhs_ctx* d_ctx = nullptr;
   std::shared_ptr<PatchHierarchy> d_hierarchy;
   std::shared_ptr<Patch> d_patch;
   size_t d_ncell;
   int d_ng = 1;
   // PatchData ids, named like the reference's members
   int d_phase_scratch_id = -1, d_quat_scratch_id = -1, d_conc_scratch_id = -1,
       d_temperature_scratch_id = -1;
   int d_phase_mobility_id = -1, d_quat_mobility_id = -1, d_flux_id = -1, d_flux_conc_id = -1;
   int d_quat_diffs_id = -1, d_quat_grad_cell_id = -1, d_quat_grad_side_id = -1,
       d_quat_grad_side_copy_id = -1, d_quat_grad_modulus_id = -1, d_quat_symm_rotation_id = -1;
   int d_face_coef_id = -1, d_quat_flux_id = -1, d_lambda_id = -1;
   int d_conc_l_id = -1, d_conc_a_id = -1, d_conc_l_ref_id = -1, d_conc_a_ref_id = -1;
   int d_f_l_id = -1, d_f_a_id = -1, d_diff0_id = -1, d_diff1_id = -1;
   int d_cp_id = -1, d_eq_temperature_id = -1;
   int d_ydot_phase_id = -1, d_ydot_quat_id = -1, d_ydot_conc_id = -1, d_ydot_temperature_id = -1;
   std::shared_ptr<PhaseFluxStrategy> d_phase_flux_strategy;
   std::shared_ptr<FreeEnergyStrategy> d_free_energy_strategy;
   std::shared_ptr<PhaseRHSStrategyWithQ> d_phase_rhs_strategy;
   std::shared_ptr<QuatMobilityStrategy> d_mobility_strategy;
   std::shared_ptr<QuatGradStrategy> d_quat_grad_strategy;
   std::shared_ptr<QuatFaceCoeff> d_quat_face_coeff;
   std::shared_ptr<QuatSysSolver> d_quat_sys_solver;
   std::shared_ptr<CompositionRHSStrategy> d_composition_rhs_strategy;
   std::shared_ptr<PhaseConcentrationsStrategy> d_phase_conc_strategy;
   std::shared_ptr<TemperatureRHSStrategy> d_temperature_rhs_strategy;
   // block preconditioners (named like the reference's members, QuatIntegrator.h)
   bool d_uniform_temperature_filled = false;
   bool d_use_preconditioner = false;
   int d_precond_cycles = 0;
   long d_precond_setups = 0, d_precond_solves = 0;
   int d_phase_precond_c_id = -1, d_conc_l_g0_id = -1, d_conc_a_g0_id = -1;
   bool d_precond_has_dquatdphi = false;
   bool d_precond_left = false;
   double d_precond_gamma = 0.0;
   int d_quat_mobility_deriv_id = -1, d_quat_diffusion_deriv_id = -1, d_phase_sol_id = -1, d_quat_rhs_id = -1,
       d_sqrt_m_id = -1, d_face_coef_scratch_id = -1;
   std::shared_ptr<DerivDiffusionCoeffForQuat> d_diffusion4quatderiv;
   std::shared_ptr<PhaseFACSolver> d_phase_sys_solver;
   std::shared_ptr<ConcFACSolver> d_conc_sys_solver;
   std::shared_ptr<TemperatureFACSolver> d_temperature_sys_solver;
};

inline double DeviceVectorOps::wdot(const Vec& x, const Vec& y, const Vec& w)
{
   double r = 0.0;
   check(ampe_vec_wdot(d_ctx, &x, &y, &w, &r, nullptr), "wdot");
   return (d_owner && d_cfg.nranks > 1) ? d_owner->sumReduction(r) : r;
}
inline long long DeviceVectorOps::length() const
{
   if (!(d_owner && d_cfg.nranks > 1)) return d_length;
   if (d_global_length < 0) d_global_length = (long long)d_owner->sumReduction((double)d_length);  // exact up to 2^53
   return d_global_length;
}
inline int DeviceVectorOps::rhs(double t, const Vec& y, Vec& ydot, int fd_flag)
{
   if (d_owner && d_owner->halo())
      check(ampe_rhs_eval_slab(d_ctx, d_owner->halo(), t, &y, &ydot, fd_flag, nullptr), "ampe_rhs_eval_slab");
   else if (d_cfg.nranks > 1)
      throw std::runtime_error("slab ranks: connect the ghost-plane exchange first (haloExport / haloConnect)");
   else
      check(ampe_rhs_eval(d_ctx, t, &y, &ydot, fd_flag, nullptr), "ampe_rhs_eval");
   return 0;
}
// what QuatModel::Advance does after the integrator returns: normalizeQuat (QuatModel.cc:4222-4262)
// and resetRefPhaseConcentrations (QuatModel.cc:5218-5231)
inline void DeviceVectorOps::postStep(Vec& y)
{
   if (d_cfg.evolve_quat) check(ampe_normalize_quat(d_ctx, &y, nullptr), "normalizeQuat");
   const bool kks = d_cfg.conc_rhs_form == AMPE_CONC_KKS || d_cfg.conc_rhs_form == AMPE_CONC_EBS;
   if (kks && d_cfg.free_energy == AMPE_FE_CALPHAD) {
      if (d_owner && d_owner->halo())
         check(ampe_rhs_set_ref_concentrations_slab(d_ctx, d_owner->halo(), nullptr, nullptr, nullptr), "resetRef");
      else
         check(ampe_rhs_set_ref_concentrations(d_ctx, nullptr, nullptr, nullptr), "resetRef");
   }
}
inline bool DeviceVectorOps::preconditioned() const { return d_owner && d_owner->usePreconditioner(); }
inline int DeviceVectorOps::precondSetup(double t, const Vec& y, double gamma)
{
   return d_owner->CVSpgmrPrecondSet(t, &y, gamma);
}
inline void DeviceVectorOps::precondSolve(const Vec& r, Vec& z) { d_owner->CVSpgmrPrecondSolve(&r, &z); }

}  // namespace ampe_host
