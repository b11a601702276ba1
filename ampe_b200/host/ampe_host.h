// Host-side mirror of the reference's Strategy-class API for the RHS path, without
// SAMRAI.  Class and method names, argument meaning and the order of calls follow the
// reference (citations per class); data are referenced by integer PatchData ids resolved
// through patch->getPatchData(id) exactly like SAMRAI; every method launches the
// corresponding ampe_k_* CUDA kernel (include/ampe_b200_kernels.h) on device arrays laid
// out like pdat::CellData / pdat::SideData with ghost widths.
//
// This layer keeps the reference's UNFUSED pass structure so that a Strategy can be
// swapped individually; the production path is QuatIntegrator::evaluateRHSFunction ->
// ampe_rhs_eval (fused).  `use_fused` selects between the two in QuatIntegrator below.
#pragma once
#include <cuda_runtime.h>

#include <cassert>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ampe_b200.h"
#include "../../include/ampe_b200_kernels.h"
#include "../../include/ampe_b200_precond.h"

namespace ampe_host {

inline void check(int rc, const char* what)
{
   if (rc < 0) throw std::runtime_error(std::string(what) + ": " + ampe_last_error());
}
inline void cuda_check(cudaError_t e, const char* what)
{
   if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// ---- SAMRAI stand-ins (hier::Box, pdat::CellData, pdat::SideData, hier::Patch, ...) --------
struct Box {
   int ndim = 2;
   int lower[3] = {0, 0, 0};
   int upper[3] = {0, 0, 0};
   int numberCells(int d) const { return upper[d] - lower[d] + 1; }
};

struct PatchData {
   virtual ~PatchData() {}
};

template <typename T>
class CellData : public PatchData
{
 public:
   CellData(const Box& box, int depth, int ghosts) : d_box(box), d_depth(depth), d_ng(ghosts)
   {
      d_comp = 1;
      for (int d = 0; d < 3; d++) d_comp *= (size_t)(box.numberCells(d) + (d < box.ndim ? 2 * ghosts : 0));
      cuda_check(cudaMalloc(&d_ptr, d_comp * depth * sizeof(T)), "CellData");
      cuda_check(cudaMemset(d_ptr, 0, d_comp * depth * sizeof(T)), "CellData");
   }
   ~CellData() { cudaFree(d_ptr); }
   T* getPointer(int d = 0) const { return d_ptr + d_comp * d; }
   int getDepth() const { return d_depth; }
   int getGhostCellWidth() const { return d_ng; }
   const Box& getBox() const { return d_box; }
   size_t size() const { return d_comp * d_depth; }
   void fillAll(int byte) { cuda_check(cudaMemset(d_ptr, byte, size() * sizeof(T)), "fillAll"); }

 private:
   Box d_box;
   int d_depth, d_ng;
   size_t d_comp;
   T* d_ptr = nullptr;
};

template <typename T>
class SideData : public PatchData
{
 public:
   SideData(const Box& box, int depth, int ghosts) : d_box(box), d_depth(depth), d_ng(ghosts)
   {
      for (int a = 0; a < 3; a++) {
         d_ptr[a] = nullptr;
         d_comp[a] = 0;
      }
      for (int a = 0; a < box.ndim; a++) {
         d_comp[a] = 1;
         for (int d = 0; d < 3; d++)
            d_comp[a] *= (size_t)(box.numberCells(d) + (d < box.ndim ? 2 * ghosts : 0) + (d == a ? 1 : 0));
         cuda_check(cudaMalloc(&d_ptr[a], d_comp[a] * depth * sizeof(T)), "SideData");
         cuda_check(cudaMemset(d_ptr[a], 0, d_comp[a] * depth * sizeof(T)), "SideData");
      }
   }
   ~SideData()
   {
      for (int a = 0; a < 3; a++) cudaFree(d_ptr[a]);
   }
   T* getPointer(int axis, int d = 0) const { return d_ptr[axis] + d_comp[axis] * d; }
   int getDepth() const { return d_depth; }
   int getGhostCellWidth() const { return d_ng; }
   void fillAll(int byte)
   {
      for (int a = 0; a < d_box.ndim; a++)
         cuda_check(cudaMemset(d_ptr[a], byte, d_comp[a] * d_depth * sizeof(T)), "fillAll");
   }
   void copy(const SideData<T>& src)
   {
      for (int a = 0; a < d_box.ndim; a++)
         cuda_check(cudaMemcpy(d_ptr[a], src.d_ptr[a], d_comp[a] * d_depth * sizeof(T),
                               cudaMemcpyDeviceToDevice),
                    "SideData::copy");
   }
   // pointers of depth component d for all axes (what the kernels take)
   std::vector<T*> pointers(int d = 0) const
   {
      std::vector<T*> v(3, nullptr);
      for (int a = 0; a < d_box.ndim; a++) v[a] = getPointer(a, d);
      return v;
   }

 private:
   Box d_box;
   int d_depth, d_ng;
   size_t d_comp[3];
   T* d_ptr[3];
};

class Patch
{
 public:
   Patch(const Box& box, const double* dx) : d_box(box)
   {
      for (int d = 0; d < 3; d++) d_dx[d] = d < box.ndim ? dx[d] : 1.0;
   }
   const Box& getBox() const { return d_box; }
   const double* getDx() const { return d_dx; }
   int registerPatchData(std::shared_ptr<PatchData> pd)
   {
      d_data.push_back(pd);
      return (int)d_data.size() - 1;
   }
   std::shared_ptr<PatchData> getPatchData(int id) const
   {
      assert(id >= 0 && id < (int)d_data.size());
      return d_data[id];
   }
   template <typename T>
   std::shared_ptr<CellData<T>> cell(int id) const
   {
      return std::dynamic_pointer_cast<CellData<T>>(getPatchData(id));
   }
   template <typename T>
   std::shared_ptr<SideData<T>> side(int id) const
   {
      return std::dynamic_pointer_cast<SideData<T>>(getPatchData(id));
   }

 private:
   Box d_box;
   double d_dx[3];
   std::vector<std::shared_ptr<PatchData>> d_data;
};

// single uniform level with one patch per rank (AMR is off on the GPU path: SURVEY.md 8e)
struct PatchLevel {
   std::vector<std::shared_ptr<Patch>> patches;
   std::vector<std::shared_ptr<Patch>>::iterator begin() { return patches.begin(); }
   std::vector<std::shared_ptr<Patch>>::iterator end() { return patches.end(); }
};
struct PatchHierarchy {
   std::shared_ptr<PatchLevel> level;
   int getFinestLevelNumber() const { return 0; }
   std::shared_ptr<PatchLevel> getPatchLevel(int) const { return level; }
};

#define AMPE_BOX_ARGS(patch) (patch)->getBox().ndim, (patch)->getBox().lower, (patch)->getBox().upper

// ---- PhaseFluxStrategy (PhaseFluxStrategy.h:26-28) --------------------------------------------
class PhaseFluxStrategy
{
 public:
   virtual ~PhaseFluxStrategy() {}
   virtual void computeFluxes(std::shared_ptr<PatchLevel> level, int phase_id, int quat_id,
                              int flux_id) = 0;
};
// PhaseFluxStrategySimple.cc:19-73 -> GRADIENT_FLUX
class PhaseFluxStrategySimple : public PhaseFluxStrategy
{
 public:
   explicit PhaseFluxStrategySimple(double epsilon_phase) : d_epsilon_phase(epsilon_phase) {}
   void computeFluxes(std::shared_ptr<PatchLevel> level, int phase_id, int, int flux_id) override
   {
      for (auto& patch : *level) {
         auto phase = patch->cell<double>(phase_id);
         auto flux = patch->side<double>(flux_id);
         auto f = flux->pointers();
         check(ampe_k_gradient_flux(AMPE_BOX_ARGS(patch), patch->getDx(), d_epsilon_phase,
                                    phase->getPointer(), phase->getGhostCellWidth(), f.data(),
                                    flux->getGhostCellWidth(), nullptr),
               "GRADIENT_FLUX");
      }
   }

 private:
   double d_epsilon_phase;
};
// PhaseFluxStrategyIsotropic.cc:18-64 -> COMPUTE_FLUX_ISOTROPIC
class PhaseFluxStrategyIsotropic : public PhaseFluxStrategy
{
 public:
   explicit PhaseFluxStrategyIsotropic(double epsilon_phase) : d_epsilon_phase(epsilon_phase) {}
   void computeFluxes(std::shared_ptr<PatchLevel> level, int phase_id, int, int flux_id) override
   {
      for (auto& patch : *level) {
         auto phase = patch->cell<double>(phase_id);
         auto flux = patch->side<double>(flux_id);
         auto f = flux->pointers();
         check(ampe_k_compute_flux_isotropic(AMPE_BOX_ARGS(patch), patch->getDx(), d_epsilon_phase,
                                             phase->getPointer(), phase->getGhostCellWidth(),
                                             f.data(), flux->getGhostCellWidth(), nullptr),
               "COMPUTE_FLUX_ISOTROPIC");
      }
   }

 private:
   double d_epsilon_phase;
};
// PhaseFluxStrategyAnisotropy.cc:19-78 -> ANISOTROPIC_GRADIENT_FLUX
class PhaseFluxStrategyAnisotropy : public PhaseFluxStrategy
{
 public:
   PhaseFluxStrategyAnisotropy(double epsilon_phase, double nu, int knumber)
       : d_epsilon_phase(epsilon_phase), d_nu(nu), d_knumber(knumber)
   {
   }
   void computeFluxes(std::shared_ptr<PatchLevel> level, int phase_id, int quat_id,
                      int flux_id) override
   {
      assert(quat_id >= 0);
      for (auto& patch : *level) {
         auto phase = patch->cell<double>(phase_id);
         auto quat = patch->cell<double>(quat_id);
         auto flux = patch->side<double>(flux_id);
         auto f = flux->pointers();
         check(ampe_k_anisotropic_gradient_flux(
                   AMPE_BOX_ARGS(patch), patch->getDx(), d_epsilon_phase, d_nu, d_knumber,
                   phase->getPointer(), phase->getGhostCellWidth(), quat->getPointer(),
                   quat->getGhostCellWidth(), quat->getDepth(), f.data(), flux->getGhostCellWidth(),
                   nullptr),
               "ANISOTROPIC_GRADIENT_FLUX");
      }
   }

 private:
   double d_epsilon_phase, d_nu;
   int d_knumber;
};

// ---- FreeEnergyStrategy (FreeEnergyStrategy.h:44-92) --------------------------------------------
class FreeEnergyStrategy
{
 public:
   virtual ~FreeEnergyStrategy() {}
   virtual void computeFreeEnergyLiquid(Patch& patch, int temperature_id, int fl_id, bool gp) = 0;
   virtual void computeFreeEnergySolidA(Patch& patch, int temperature_id, int fa_id, bool gp) = 0;
   virtual void addDrivingForce(double time, Patch& patch, int temperature_id, int phase_id,
                                int eta_id, int conc_id, int f_l_id, int f_a_id, int f_b_id,
                                int rhs_id) = 0;
};
// CALPHADFreeEnergyStrategyBinary.cc:251-327, 521-683 and QuadraticFreeEnergyStrategy.cc
class KKSFreeEnergyStrategy : public FreeEnergyStrategy
{
 public:
   KKSFreeEnergyStrategy(const ampe_rhs_config& cfg, int conc_l_id, int conc_a_id)
       : d_cfg(cfg), d_conc_l_id(conc_l_id), d_conc_a_id(conc_a_id)
   {
   }
   void computeFreeEnergyLiquid(Patch& patch, int, int fl_id, bool) override
   {
      auto c = patch.cell<double>(d_conc_l_id);
      auto f = patch.cell<double>(fl_id);
      check(ampe_k_compute_free_energy(&d_cfg, patch.getBox().lower, patch.getBox().upper,
                                       c->getPointer(), c->getGhostCellWidth(), f->getPointer(), 0,
                                       nullptr),
            "computeFreeEnergyLiquid");
   }
   void computeFreeEnergySolidA(Patch& patch, int, int fa_id, bool) override
   {
      auto c = patch.cell<double>(d_conc_a_id);
      auto f = patch.cell<double>(fa_id);
      check(ampe_k_compute_free_energy(&d_cfg, patch.getBox().lower, patch.getBox().upper,
                                       c->getPointer(), c->getGhostCellWidth(), f->getPointer(), 1,
                                       nullptr),
            "computeFreeEnergySolidA");
   }
   void addDrivingForce(double, Patch& patch, int, int phase_id, int, int, int f_l_id, int f_a_id,
                        int, int rhs_id) override
   {
      auto phi = patch.cell<double>(phase_id);
      auto fl = patch.cell<double>(f_l_id), fa = patch.cell<double>(f_a_id);
      auto cl = patch.cell<double>(d_conc_l_id), ca = patch.cell<double>(d_conc_a_id);
      auto rhs = patch.cell<double>(rhs_id);
      check(ampe_k_add_driving_force(&d_cfg, patch.getBox().lower, patch.getBox().upper,
                                     phi->getPointer(), phi->getGhostCellWidth(), fl->getPointer(),
                                     fa->getPointer(), cl->getPointer(), ca->getPointer(),
                                     cl->getGhostCellWidth(), rhs->getPointer(),
                                     rhs->getGhostCellWidth(), nullptr),
            "addDrivingForce");
   }

 private:
   ampe_rhs_config d_cfg;
   int d_conc_l_id, d_conc_a_id;
};
// BiasDoubleWellUTRCFreeEnergyStrategy.cc:32-77 -> COMPUTERHSBIASWELL with a constant melting T
class BiasDoubleWellUTRCFreeEnergyStrategy : public FreeEnergyStrategy
{
 public:
   BiasDoubleWellUTRCFreeEnergyStrategy(double alpha, double gamma, int eq_temperature_id)
       : d_alpha(alpha), d_gamma(gamma), d_te_id(eq_temperature_id)
   {
   }
   void computeFreeEnergyLiquid(Patch&, int, int, bool) override {}
   void computeFreeEnergySolidA(Patch&, int, int, bool) override {}
   void addDrivingForce(double, Patch& patch, int temperature_id, int phase_id, int, int, int, int,
                        int, int rhs_id) override
   {
      auto phi = patch.cell<double>(phase_id), T = patch.cell<double>(temperature_id);
      auto te = patch.cell<double>(d_te_id), rhs = patch.cell<double>(rhs_id);
      check(ampe_k_computerhsbiaswell(patch.getBox().ndim, patch.getBox().lower, patch.getBox().upper,
                                      phi->getPointer(), phi->getGhostCellWidth(), T->getPointer(),
                                      T->getGhostCellWidth(), d_alpha, d_gamma, te->getPointer(),
                                      te->getGhostCellWidth(), rhs->getPointer(),
                                      rhs->getGhostCellWidth(), nullptr),
            "COMPUTERHSBIASWELL");
   }

 private:
   double d_alpha, d_gamma;
   int d_te_id;
};

// ---- PhaseConcentrationsStrategy (PhaseConcentrationsStrategy.cc:26-125) -------------------------
class PhaseConcentrationsStrategy
{
 public:
   PhaseConcentrationsStrategy(const ampe_rhs_config& cfg, int conc_l_id, int conc_a_id,
                               int conc_l_ref_id, int conc_a_ref_id)
       : d_cfg(cfg), d_cl(conc_l_id), d_ca(conc_a_id), d_clr(conc_l_ref_id), d_car(conc_a_ref_id)
   {
   }
   // returns the number of cells whose Newton failed (the reference aborts: .cc:118)
   int computePhaseConcentrations(std::shared_ptr<PatchHierarchy> hierarchy, int temperature_id,
                                  int phase_id, int eta_id, int conc_id)
   {
      (void)temperature_id;
      (void)eta_id;
      int nfail = 0;
      for (auto& patch : *hierarchy->getPatchLevel(0)) {
         auto phi = patch->cell<double>(phase_id), c = patch->cell<double>(conc_id);
         auto cl = patch->cell<double>(d_cl), ca = patch->cell<double>(d_ca);
         auto clr = patch->cell<double>(d_clr), car = patch->cell<double>(d_car);
         int rc = ampe_k_compute_phase_concentrations(
             &d_cfg, patch->getBox().lower, patch->getBox().upper, phi->getPointer(),
             phi->getGhostCellWidth(), c->getPointer(), c->getGhostCellWidth(), clr->getPointer(),
             car->getPointer(), cl->getPointer(), ca->getPointer(), cl->getGhostCellWidth(), nullptr);
         check(rc, "computePhaseConcentrationsOnPatch");
         nfail += rc;
      }
      return nfail;
   }

 private:
   ampe_rhs_config d_cfg;
   int d_cl, d_ca, d_clr, d_car;
};

// ---- QuatGradStrategy (SimpleQuatGradStrategy.cc:49-157, computeQDiffs.cc) ----------------------
class QuatGradStrategy
{
 public:
   QuatGradStrategy(int qlen, bool symmetry_aware, int rotation_id)
       : d_qlen(qlen), d_symm(symmetry_aware), d_rot_id(rotation_id)
   {
   }
   void computeDiffs(std::shared_ptr<PatchHierarchy> h, int quat_id, int diff_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto q = patch->cell<double>(quat_id);
         auto d = patch->side<double>(diff_id);
         // symmetric diffs first, non-symmetric offset by qlen (computeQDiffs.cc:246-249)
         auto nonsymm = d->pointers(d_symm ? d_qlen : 0);
         check(ampe_k_quatdiffs(AMPE_BOX_ARGS(patch), d_qlen, q->getPointer(), q->getGhostCellWidth(),
                                nonsymm.data(), d->getGhostCellWidth(), nullptr),
               "QUATDIFFS");
         if (d_symm) {
            auto symm = d->pointers(0);
            auto rot = patch->side<int>(d_rot_id);
            auto iq = rot->pointers();
            std::vector<const int*> ciq(iq.begin(), iq.end());
            check(ampe_k_quatdiffs_symm(AMPE_BOX_ARGS(patch), d_qlen, q->getPointer(),
                                        q->getGhostCellWidth(), symm.data(), d->getGhostCellWidth(),
                                        ciq.data(), rot->getGhostCellWidth(), nullptr),
                  "QUATDIFFS_SYMM");
         }
      }
   }
   void computeGradCell(std::shared_ptr<PatchHierarchy> h, int diff_id, int grad_cell_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto d = patch->side<double>(diff_id);
         auto g = patch->cell<double>(grad_cell_id);
         auto dp = d->pointers(0);
         std::vector<double*> gp(3, nullptr);
         for (int a = 0; a < patch->getBox().ndim; a++) gp[a] = g->getPointer(a * d_qlen);
         if (d_symm) {
            auto rot = patch->side<int>(d_rot_id);
            auto iq = rot->pointers();
            std::vector<const int*> ciq(iq.begin(), iq.end());
            check(ampe_k_quatgrad_cell_symm(AMPE_BOX_ARGS(patch), d_qlen, patch->getDx(), dp.data(),
                                            d->getGhostCellWidth(), gp.data(), g->getGhostCellWidth(),
                                            ciq.data(), rot->getGhostCellWidth(), nullptr),
                  "QUATGRAD_CELL_SYMM");
         } else {
            check(ampe_k_quatgrad_cell(AMPE_BOX_ARGS(patch), d_qlen, patch->getDx(), dp.data(),
                                       d->getGhostCellWidth(), gp.data(), g->getGhostCellWidth(),
                                       nullptr),
                  "QUATGRAD_CELL");
         }
      }
   }
   void computeGradSide(std::shared_ptr<PatchHierarchy> h, int diff_id, int grad_side_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto d = patch->side<double>(diff_id);
         auto g = patch->side<double>(grad_side_id);
         auto dp = d->pointers(0);
         auto gp = g->pointers(0);
         if (d_symm) {
            auto rot = patch->side<int>(d_rot_id);
            auto iq = rot->pointers();
            std::vector<const int*> ciq(iq.begin(), iq.end());
            check(ampe_k_quatgrad_side_symm(AMPE_BOX_ARGS(patch), d_qlen, patch->getDx(), dp.data(),
                                            d->getGhostCellWidth(), gp.data(), g->getGhostCellWidth(),
                                            ciq.data(), rot->getGhostCellWidth(), nullptr),
                  "QUATGRAD_SIDE_SYMM");
         } else {
            check(ampe_k_quatgrad_side(AMPE_BOX_ARGS(patch), d_qlen, patch->getDx(), dp.data(),
                                       d->getGhostCellWidth(), gp.data(), g->getGhostCellWidth(),
                                       nullptr),
                  "QUATGRAD_SIDE");
         }
      }
   }
   // QuatGradModulusStrategy.cc:18-107
   void computeGradModulus(std::shared_ptr<PatchHierarchy> h, int grad_cell_id, int modulus_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto g = patch->cell<double>(grad_cell_id);
         auto m = patch->cell<double>(modulus_id);
         std::vector<double*> gp(3, nullptr);
         for (int a = 0; a < patch->getBox().ndim; a++) gp[a] = g->getPointer(a * d_qlen);
         check(ampe_k_quatgrad_modulus(AMPE_BOX_ARGS(patch), d_qlen, gp.data(), g->getGhostCellWidth(),
                                       m->getPointer(), m->getGhostCellWidth(), nullptr),
               "QUATGRAD_MODULUS");
      }
   }
   void computeGradModulusFromSides(std::shared_ptr<PatchHierarchy> h, int grad_side_id,
                                    int modulus_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto g = patch->side<double>(grad_side_id);
         auto m = patch->cell<double>(modulus_id);
         auto gp = g->pointers(0);
         check(ampe_k_quatgrad_modulus_from_sides_compact(AMPE_BOX_ARGS(patch), d_qlen, gp.data(),
                                                          g->getGhostCellWidth(), m->getPointer(),
                                                          m->getGhostCellWidth(), nullptr),
               "QUATGRAD_MODULUS_FROM_SIDES_COMPACT");
      }
   }

 private:
   int d_qlen;
   bool d_symm;
   int d_rot_id;
};

// ---- QuatMobilityStrategy (QuatModel.cc:4277-4291, 4532-4577) ---------------------------------
class QuatMobilityStrategy
{
 public:
   explicit QuatMobilityStrategy(const ampe_rhs_config& cfg) : d_cfg(cfg) {}
   void computePhaseMobility(std::shared_ptr<PatchHierarchy> h, int, int mobility_id)
   {
      // computeUniformPhaseMobility: fill with phi_mobility (a constant: once per array)
      if (d_uniform_filled_id == mobility_id) return;
      d_uniform_filled_id = mobility_id;
      for (auto& patch : *h->getPatchLevel(0)) {
         auto m = patch->cell<double>(mobility_id);
         std::vector<double> v(m->size(), d_cfg.phi_mobility);
         cuda_check(cudaMemcpy(m->getPointer(), v.data(), v.size() * sizeof(double),
                               cudaMemcpyHostToDevice),
                    "computeUniformPhaseMobility");
      }
   }
   void computeQuatMobility(std::shared_ptr<PatchHierarchy> h, int phase_id, int mobility_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto phi = patch->cell<double>(phase_id);
         auto m = patch->cell<double>(mobility_id);
         check(ampe_k_quatmobility(AMPE_BOX_ARGS(patch), phi->getPointer(), phi->getGhostCellWidth(),
                                   m->getPointer(), m->getGhostCellWidth(), d_cfg.quat_mobility,
                                   d_cfg.min_quat_mobility, d_cfg.quat_mobility_func,
                                   d_cfg.quat_mobility_alt_scale, nullptr),
               "QUATMOBILITY");
      }
   }
   // computeQuatMobilityDeriv (QuatIntegrator.cc:2978-2983, with precond_has_dquatdphi)
   void computeQuatMobilityDeriv(std::shared_ptr<PatchHierarchy> h, int phase_id, int mobility_deriv_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto phi = patch->cell<double>(phase_id);
         auto m = patch->cell<double>(mobility_deriv_id);
         check(ampe_k_quatmobilityderiv(AMPE_BOX_ARGS(patch), phi->getPointer(), phi->getGhostCellWidth(),
                                        m->getPointer(), m->getGhostCellWidth(), d_cfg.quat_mobility,
                                        d_cfg.min_quat_mobility, d_cfg.quat_mobility_func,
                                        d_cfg.quat_mobility_alt_scale, nullptr),
               "QUATMOBILITYDERIV");
      }
   }

 private:
   ampe_rhs_config d_cfg;
   int d_uniform_filled_id = -1;
};

// ---- DerivDiffusionCoeffForQuat (DerivDiffusionCoeffForQuat.cc:60-148) ---------------------------
class DerivDiffusionCoeffForQuat
{
 public:
   DerivDiffusionCoeffForQuat(const ampe_rhs_config& cfg, int quat_diffusion_deriv_id)
       : d_cfg(cfg), d_quat_diffusion_deriv_id(quat_diffusion_deriv_id)
   {
   }
   void setDerivDiffusion(std::shared_ptr<PatchHierarchy> h, int phase_scratch_id, int temperature_scratch_id,
                          int quat_grad_side_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto phi = patch->cell<double>(phase_scratch_id), T = patch->cell<double>(temperature_scratch_id);
         auto gq = patch->side<double>(quat_grad_side_id), dd = patch->side<double>(d_quat_diffusion_deriv_id);
         auto g = gq->pointers(0), d = dd->pointers(0);
         check(ampe_k_quatdiffusionderiv(AMPE_BOX_ARGS(patch), 2. * d_cfg.H_parameter, T->getPointer(),
                                         T->getGhostCellWidth(), phi->getPointer(), phi->getGhostCellWidth(),
                                         d_cfg.qlen, g.data(), gq->getGhostCellWidth(), d.data(),
                                         dd->getGhostCellWidth(), d_cfg.quat_grad_floor, d_cfg.grad_floor_type,
                                         d_cfg.orient_interp1, d_cfg.avg_func, nullptr),
               "QUATDIFFUSIONDERIV");
      }
   }

 private:
   ampe_rhs_config d_cfg;
   int d_quat_diffusion_deriv_id;
};

// ---- QuatFaceCoeff (QuatFaceCoeff.cc:39-121) ----------------------------------------------------
class QuatFaceCoeff
{
 public:
   explicit QuatFaceCoeff(const ampe_rhs_config& cfg) : d_cfg(cfg) {}
   void computeFaceCoefs(std::shared_ptr<PatchHierarchy> h, int phase_id, int temp_id, int grad_q_id,
                         int face_coef_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto phi = patch->cell<double>(phase_id), T = patch->cell<double>(temp_id);
         auto gq = patch->side<double>(grad_q_id), fc = patch->side<double>(face_coef_id);
         auto g = gq->pointers(0), f = fc->pointers(0);
         check(ampe_k_compute_face_coef(AMPE_BOX_ARGS(patch), d_cfg.qlen, d_cfg.epsilon_q,
                                        phi->getPointer(), phi->getGhostCellWidth(), T->getPointer(),
                                        T->getGhostCellWidth(), 2. * d_cfg.H_parameter, g.data(),
                                        gq->getGhostCellWidth(), f.data(), fc->getGhostCellWidth(),
                                        d_cfg.quat_grad_floor, d_cfg.grad_floor_type,
                                        d_cfg.orient_interp1, d_cfg.orient_interp2, d_cfg.avg_func,
                                        nullptr),
               "COMPUTE_FACE_COEF");
      }
   }

 private:
   ampe_rhs_config d_cfg;
};

// ---- QuatSysSolver::evaluateRHS -> QuatFACOps::evaluateRHS (QuatFACOps.cc:1861-1889, 1959-2136)
class QuatSysSolver
{
 public:
   QuatSysSolver(const ampe_rhs_config& cfg, std::shared_ptr<PatchHierarchy> h,
                 std::shared_ptr<QuatFaceCoeff> face_coeff, int face_coef_scratch_id,
                 int flux_scratch_id, int lambda_id)
       : d_cfg(cfg), d_h(h), d_face_coeff(face_coeff), d_fc_id(face_coef_scratch_id),
         d_flux_id(flux_scratch_id), d_lambda_id(lambda_id)
   {
   }
   ~QuatSysSolver() { ampe_mg_destroy(d_mg); }
   int getFaceDiffCoeffScratchId() const { return d_fc_id; }
   // ---- preconditioner side ---------------------------------------------
   // QuatSysSolver::setOperatorCoefficients (QuatSysSolver.cc:268-292) -> QuatFACOps::
   // setOperatorCoefficients (QuatFACOps.cc:735-818): face coefficients from (phase, T, grad_q), the
   // square root of the mobility, then QuatLevelSolver::setMatrixCoefficients.  The derivative ids
   // feed the dquat/dphi coupling block only (precond_has_dquatdphi; -1 without it).
   void setOperatorCoefficients(double gamma, int mobility_id, int mobility_deriv_id, int phase_id,
                                int temperature_id, int face_coef_deriv_id, int grad_q_id, int q_id)
   {
      d_q_local_id = q_id;  // the reference copies q (q_local_data->copy); the scratch q is not touched in between
      d_mobility_id = mobility_id;
      d_m_deriv_id = mobility_deriv_id;
      d_face_coef_deriv_id = face_coef_deriv_id;
      d_face_coeff->computeFaceCoefs(d_h, phase_id, temperature_id, grad_q_id, d_fc_id);
      auto patch = d_h->getPatchLevel(0)->patches.front();
      if (!d_mg) {
         int n[3];
         for (int d = 0; d < 3; d++) n[d] = patch->getBox().numberCells(d);
         check(ampe_mg_create_multi(patch->getBox().ndim, n, patch->getDx(), 1, d_cfg.qlen, &d_mg),
               "ampe_mg_create(quat)");
         // QuatFACOps::setPhysicalBcCoefObject: the Quat block of BoundaryConditions
         // (slab ranks: block Jacobi over the ranks -- no flux through the faces between ranks)
         int zs[3] = {d_cfg.zero_slope[0], d_cfg.zero_slope[1], d_cfg.zero_slope[2]};
         if (d_cfg.nranks > 1) zs[d_cfg.ndim - 1] = 1;
         check(ampe_mg_set_zero_slope(d_mg, zs), "ampe_mg_set_zero_slope(quat)");
      }
      auto mob = patch->cell<double>(mobility_id);
      auto fc = patch->side<double>(d_fc_id);
      auto f = fc->pointers(0);
      std::vector<const double*> cf(f.begin(), f.end());
      check(ampe_mg_set_quat(d_mg, gamma, mob->getPointer(), mob->getGhostCellWidth(), cf.data(),
                             fc->getGhostCellWidth(), nullptr),
            "QuatSysSolver::setOperatorCoefficients");
   }
   // QuatSysSolver::solveSystem (QuatSysSolver.cc:296-347), one depth after the other like
   // QuatLevelSolver::solveSystem; q_soln / q_rhs: ghost-0 device arrays of depth qlen
   bool solveSystem(double* q_soln, const double* q_rhs, int ncycles)
   {
      if (!d_mg) throw std::runtime_error("QuatSysSolver::solveSystem before setOperatorCoefficients");
      // all qlen depths in one solve: they share the matrix, every pass reads the coefficients once
      check(ampe_mg_solve(d_mg, q_rhs, q_soln, ncycles, 1, nullptr), "QuatSysSolver::solveSystem");
      return true;
   }
   ampe_mg* levelSolver() const { return d_mg; }
   // QuatSysSolver::multiplyDQuatDPhiBlock -> QuatFACOps::multiplyDQuatDPhiBlock (QuatFACOps.cc:1892-1956):
   // out = [mobility'(phi) div(fc grad q)] phase + sqrt_m div(fc'[phase] grad q); phase_id: CellData with
   // filled ghosts, out_id: CellData depth qlen; scratch ids: sqrt_m (cell, ghosts of the mobility),
   // face_coef_scratch (side, depth 1)
   void multiplyDQuatDPhiBlock(int phase_id, int out_id, int sqrt_m_id, int face_coef_scratch_id)
   {
      if (d_m_deriv_id < 0 || d_face_coef_deriv_id < 0)
         throw std::runtime_error("multiplyDQuatDPhiBlock: setOperatorCoefficients was called without the derivatives");
      const int Q = d_cfg.qlen;
      for (auto& patch : *d_h->getPatchLevel(0)) {
         auto out = patch->cell<double>(out_id), q = patch->cell<double>(d_q_local_id);
         auto phase = patch->cell<double>(phase_id);
         auto mob = patch->cell<double>(d_mobility_id), sq = patch->cell<double>(sqrt_m_id);
         auto mder = patch->cell<double>(d_m_deriv_id);
         auto fc = patch->side<double>(d_fc_id), fcs = patch->side<double>(face_coef_scratch_id);
         auto dpr = patch->side<double>(d_face_coef_deriv_id), flux = patch->side<double>(d_flux_id);
         auto c = fc->pointers(0), cs = fcs->pointers(0), dp = dpr->pointers(0), f = flux->pointers(0);
         out->fillAll(0);
         // takeSquareRootOnPatch of a copy of the mobility (QuatFACOps.cc:779-783)
         cuda_check(cudaMemcpy(sq->getPointer(), mob->getPointer(), mob->size() * sizeof(double),
                               cudaMemcpyDeviceToDevice),
                    "sqrt_m copy");
         check(ampe_k_take_square_root(AMPE_BOX_ARGS(patch), sq->getPointer(), sq->getGhostCellWidth(), nullptr),
               "TAKE_SQUARE_ROOT");
         // accumulateOperatorOnLevel(d_m_deriv_id, d_face_coef_id, d_q_local_id, -1, out_id, ...)
         check(ampe_k_compute_flux(AMPE_BOX_ARGS(patch), Q, c.data(), fc->getGhostCellWidth(), q->getPointer(),
                                   q->getGhostCellWidth(), patch->getDx(), f.data(), flux->getGhostCellWidth(),
                                   nullptr),
               "COMPUTE_FLUX");
         check(ampe_k_add_quat_op(AMPE_BOX_ARGS(patch), Q, mder->getPointer(), mder->getGhostCellWidth(), f.data(),
                                  flux->getGhostCellWidth(), patch->getDx(), out->getPointer(),
                                  out->getGhostCellWidth(), nullptr),
               "ADD_QUAT_OP");
         check(ampe_k_multicomponent_multiply(AMPE_BOX_ARGS(patch), phase->getPointer(), phase->getGhostCellWidth(),
                                              out->getPointer(), out->getGhostCellWidth(), Q, nullptr),
               "MULTICOMPONENT_MULTIPLY");
         // computeDQuatDPhiFaceCoefs + accumulateOperatorOnLevel(d_sqrt_m_id, d_face_coef_scratch_id, ...)
         check(ampe_k_compute_dquatdphi_face_coef(AMPE_BOX_ARGS(patch), Q, dp.data(), dpr->getGhostCellWidth(),
                                                  phase->getPointer(), phase->getGhostCellWidth(), cs.data(),
                                                  fcs->getGhostCellWidth(), nullptr),
               "COMPUTE_DQUATDPHI_FACE_COEF");
         check(ampe_k_compute_flux(AMPE_BOX_ARGS(patch), Q, cs.data(), fcs->getGhostCellWidth(), q->getPointer(),
                                   q->getGhostCellWidth(), patch->getDx(), f.data(), flux->getGhostCellWidth(),
                                   nullptr),
               "COMPUTE_FLUX");
         check(ampe_k_add_quat_op(AMPE_BOX_ARGS(patch), Q, sq->getPointer(), sq->getGhostCellWidth(), f.data(),
                                  flux->getGhostCellWidth(), patch->getDx(), out->getPointer(),
                                  out->getGhostCellWidth(), nullptr),
               "ADD_QUAT_OP");
      }
   }
   void evaluateRHS(int phase_id, int temperature_id, int grad_q_id, int grad_q_copy_id,
                    int rotations_id, int mobility_id, int solution_id, int rhs_id,
                    bool use_gradq_for_flux)
   {
      (void)rotations_id;
      const int Q = d_cfg.qlen;
      for (auto& patch : *d_h->getPatchLevel(0)) patch->cell<double>(rhs_id)->fillAll(0);
      d_face_coeff->computeFaceCoefs(d_h, phase_id, temperature_id, grad_q_copy_id, d_fc_id);
      for (auto& patch : *d_h->getPatchLevel(0)) {
         auto fc = patch->side<double>(d_fc_id), flux = patch->side<double>(d_flux_id);
         auto q = patch->cell<double>(solution_id), mob = patch->cell<double>(mobility_id);
         auto rhs = patch->cell<double>(rhs_id), lam = patch->cell<double>(d_lambda_id);
         auto c = fc->pointers(0), f = flux->pointers(0);
         if (use_gradq_for_flux) {
            auto g = patch->side<double>(grad_q_id)->pointers(0);
            check(ampe_k_compute_flux_from_gradq(AMPE_BOX_ARGS(patch), Q, c.data(),
                                                 fc->getGhostCellWidth(), g.data(), f.data(),
                                                 flux->getGhostCellWidth(), nullptr),
                  "COMPUTE_FLUX_FROM_GRADQ");
         } else {
            check(ampe_k_compute_flux(AMPE_BOX_ARGS(patch), Q, c.data(), fc->getGhostCellWidth(),
                                      q->getPointer(), q->getGhostCellWidth(), patch->getDx(), f.data(),
                                      flux->getGhostCellWidth(), nullptr),
                  "COMPUTE_FLUX");
         }
         if (Q != 1) {
            check(ampe_k_compute_lambda_flux(AMPE_BOX_ARGS(patch), Q, f.data(),
                                             flux->getGhostCellWidth(), q->getPointer(),
                                             q->getGhostCellWidth(), patch->getDx(), lam->getPointer(),
                                             lam->getGhostCellWidth(), nullptr),
                  "COMPUTE_LAMBDA_FLUX");
            check(ampe_k_add_quat_proj_op(AMPE_BOX_ARGS(patch), Q, mob->getPointer(),
                                          mob->getGhostCellWidth(), f.data(), flux->getGhostCellWidth(),
                                          q->getPointer(), q->getGhostCellWidth(), lam->getPointer(),
                                          lam->getGhostCellWidth(), patch->getDx(), rhs->getPointer(),
                                          rhs->getGhostCellWidth(), nullptr),
                  "ADD_QUAT_PROJ_OP");
         } else {
            check(ampe_k_add_quat_op(AMPE_BOX_ARGS(patch), Q, mob->getPointer(), mob->getGhostCellWidth(),
                                     f.data(), flux->getGhostCellWidth(), patch->getDx(),
                                     rhs->getPointer(), rhs->getGhostCellWidth(), nullptr),
                  "ADD_QUAT_OP");
         }
      }
   }

 private:
   ampe_rhs_config d_cfg;
   std::shared_ptr<PatchHierarchy> d_h;
   std::shared_ptr<QuatFaceCoeff> d_face_coeff;
   int d_fc_id, d_flux_id, d_lambda_id;
   ampe_mg* d_mg = nullptr;
   int d_q_local_id = -1, d_mobility_id = -1, d_m_deriv_id = -1, d_face_coef_deriv_id = -1;
};

// ---- PhaseRHSStrategyWithQ (PhaseRHSStrategyWithQ.cc:90-312) ------------------------------------
class PhaseRHSStrategy
{
 public:
   virtual ~PhaseRHSStrategy() {}
   virtual void evaluateRHS(double time, std::shared_ptr<PatchHierarchy> hierarchy, int ydot_phase_id,
                            bool eval_flag) = 0;
};
class PhaseRHSStrategyWithQ : public PhaseRHSStrategy
{
 public:
   PhaseRHSStrategyWithQ(const ampe_rhs_config& cfg, int phase_scratch_id, int conc_scratch_id,
                         int quat_scratch_id, int temperature_scratch_id, int f_l_id, int f_a_id,
                         int phase_mobility_id, int flux_id, int quat_grad_modulus_id,
                         std::shared_ptr<PhaseFluxStrategy> phase_flux_strategy,
                         std::shared_ptr<FreeEnergyStrategy> free_energy_strategy)
       : d_cfg(cfg), d_phase_scratch_id(phase_scratch_id), d_conc_scratch_id(conc_scratch_id),
         d_quat_scratch_id(quat_scratch_id), d_temperature_scratch_id(temperature_scratch_id),
         d_f_l_id(f_l_id), d_f_a_id(f_a_id), d_phase_mobility_id(phase_mobility_id),
         d_flux_id(flux_id), d_quat_grad_modulus_id(quat_grad_modulus_id),
         d_phase_flux_strategy(phase_flux_strategy), d_free_energy_strategy(free_energy_strategy)
   {
   }
   void evaluateRHS(double time, std::shared_ptr<PatchHierarchy> hierarchy, int ydot_phase_id,
                    bool eval_flag) override
   {
      (void)eval_flag;
      auto level = hierarchy->getPatchLevel(0);
      d_phase_flux_strategy->computeFluxes(level, d_phase_scratch_id, d_quat_scratch_id, d_flux_id);
      for (auto& patch : *level) evaluateRHS(time, patch, ydot_phase_id);
   }

 private:
   void evaluateRHS(double time, std::shared_ptr<Patch> patch, int ydot_phase_id)
   {
      if (d_free_energy_strategy) {
         d_free_energy_strategy->computeFreeEnergyLiquid(*patch, d_temperature_scratch_id, d_f_l_id,
                                                         false);
         d_free_energy_strategy->computeFreeEnergySolidA(*patch, d_temperature_scratch_id, d_f_a_id,
                                                         false);
      }
      auto phase = patch->cell<double>(d_phase_scratch_id);
      auto rhs = patch->cell<double>(ydot_phase_id);
      auto flux = patch->side<double>(d_flux_id);
      auto T = patch->cell<double>(d_temperature_scratch_id);
      const int with_orient = d_cfg.evolve_quat ? 1 : 0;
      double* qgm = with_orient ? patch->cell<double>(d_quat_grad_modulus_id)->getPointer() : nullptr;
      auto f = flux->pointers(0);
      const char well_func_type = 'd';
      const char interpf = d_cfg.energy_interp;
      const char oi1 = d_cfg.orient_interp1, oi2 = d_cfg.orient_interp2;
      // PhaseRHSStrategyWithQ.cc:247-263 argument for argument (two-phase models: no eta, three_phase = 0)
      check(ampe_k_computerhspbg(AMPE_BOX_ARGS(patch), patch->getDx(), 2.0 * d_cfg.H_parameter,
                                 d_cfg.epsilon_q, f.data(), flux->getGhostCellWidth(), T->getPointer(),
                                 T->getGhostCellWidth(), d_cfg.phi_well_scale, 0.0, phase->getPointer(),
                                 phase->getGhostCellWidth(), nullptr, 0, qgm, 0, rhs->getPointer(), 0,
                                 &well_func_type, &well_func_type, &interpf, &oi1, &oi2, with_orient, 0, nullptr),
            "COMPUTERHSPBG");
      if (d_free_energy_strategy)
         d_free_energy_strategy->addDrivingForce(time, *patch, d_temperature_scratch_id,
                                                 d_phase_scratch_id, -1, d_conc_scratch_id, d_f_l_id,
                                                 d_f_a_id, -1, ydot_phase_id);
      // mathops.multiply(phase_rhs, phase_mobility, phase_rhs, pbox): uniform mobility
      multiplyByMobility(*patch, ydot_phase_id);
   }
   void multiplyByMobility(Patch& patch, int rhs_id)
   {
      auto rhs = patch.cell<double>(rhs_id);
      auto mob = patch.cell<double>(d_phase_mobility_id);
      check(ampe_k_cell_multiply(patch.getBox().ndim, patch.getBox().lower, patch.getBox().upper,
                                 rhs->getPointer(), rhs->getGhostCellWidth(), mob->getPointer(),
                                 mob->getGhostCellWidth(), rhs->getPointer(), rhs->getGhostCellWidth(),
                                 nullptr),
            "multiply(rhs, mobility)");
   }

   ampe_rhs_config d_cfg;
   int d_phase_scratch_id, d_conc_scratch_id, d_quat_scratch_id, d_temperature_scratch_id;
   int d_f_l_id, d_f_a_id, d_phase_mobility_id, d_flux_id, d_quat_grad_modulus_id;
   std::shared_ptr<PhaseFluxStrategy> d_phase_flux_strategy;
   std::shared_ptr<FreeEnergyStrategy> d_free_energy_strategy;
};

// ---- CompositionRHSStrategy (CompositionRHSStrategy.h:32-46) -------------------------------------
class CompositionRHSStrategy
{
 public:
   virtual ~CompositionRHSStrategy() {}
   virtual void computeFluxOnPatch(Patch& patch, int flux_id) = 0;
   virtual void setDiffusionCoeff(std::shared_ptr<PatchHierarchy>, double) {}
};
// CahnHilliardDoubleWell.cc:67-109
class CahnHilliardDoubleWell : public CompositionRHSStrategy
{
 public:
   CahnHilliardDoubleWell(const ampe_rhs_config& cfg, int conc_scratch_id)
       : d_cfg(cfg), d_conc_scratch_id(conc_scratch_id)
   {
   }
   void computeFluxOnPatch(Patch& patch, int flux_id) override
   {
      auto conc = patch.cell<double>(d_conc_scratch_id);
      auto flux = patch.side<double>(flux_id);
      assert(conc->getGhostCellWidth() > 1);
      flux->fillAll(0);
      auto f = flux->pointers(0);
      check(ampe_k_add_cahnhilliarddoublewell_flux(
                patch.getBox().ndim, patch.getBox().lower, patch.getBox().upper, patch.getDx(),
                conc->getPointer(), conc->getGhostCellWidth(), d_cfg.ch_mobility, d_cfg.ch_ca,
                d_cfg.ch_cb, d_cfg.ch_well_scale, d_cfg.ch_kappa, f.data(), flux->getGhostCellWidth(),
                nullptr),
            "ADD_CAHNHILLIARDDOUBLEWELL_FLUX");
   }

 private:
   ampe_rhs_config d_cfg;
   int d_conc_scratch_id;
};
// EBSCompositionRHSStrategy.cc:174-318 + MobilityCompositionDiffusionStrategy::setDiffusion
class EBSCompositionRHSStrategy : public CompositionRHSStrategy
{
 public:
   EBSCompositionRHSStrategy(const ampe_rhs_config& cfg, int phase_scratch_id, int conc_l_id,
                             int conc_a_id, int diff_l_id, int diff_a_id)
       : d_cfg(cfg), d_phase(phase_scratch_id), d_cl(conc_l_id), d_ca(conc_a_id), d_dl(diff_l_id),
         d_da(diff_a_id)
   {
   }
   // CompositionDiffusionStrategy::setDiffusion(hierarchy, temperature_id, phase_id)
   void setDiffusionCoeff(std::shared_ptr<PatchHierarchy> h, double) override
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto phi = patch->cell<double>(d_phase);
         auto cl = patch->cell<double>(d_cl), ca = patch->cell<double>(d_ca);
         auto dl = patch->side<double>(d_dl)->pointers(0), da = patch->side<double>(d_da)->pointers(0);
         check(ampe_k_set_ebs_diffusion(&d_cfg, patch->getBox().lower, patch->getBox().upper,
                                        phi->getPointer(), phi->getGhostCellWidth(), cl->getPointer(),
                                        ca->getPointer(), cl->getGhostCellWidth(), dl.data(), da.data(),
                                        nullptr),
               "setDiffusion");
      }
   }
   void computeFluxOnPatch(Patch& patch, int flux_id) override
   {
      auto flux = patch.side<double>(flux_id);
      flux->fillAll(0);
      auto f = flux->pointers(0);
      auto cl = patch.cell<double>(d_cl), ca = patch.cell<double>(d_ca);
      auto dl = patch.side<double>(d_dl)->pointers(0), da = patch.side<double>(d_da)->pointers(0);
      check(ampe_k_add_flux(patch.getBox().ndim, patch.getBox().lower, patch.getBox().upper,
                            patch.getDx(), cl->getPointer(), cl->getGhostCellWidth(), 1, dl.data(), 0,
                            f.data(), flux->getGhostCellWidth(), nullptr),
            "ADD_FLUX(liquid)");
      check(ampe_k_add_flux(patch.getBox().ndim, patch.getBox().lower, patch.getBox().upper,
                            patch.getDx(), ca->getPointer(), ca->getGhostCellWidth(), 1, da.data(), 0,
                            f.data(), flux->getGhostCellWidth(), nullptr),
            "ADD_FLUX(solid)");
   }

 private:
   ampe_rhs_config d_cfg;
   int d_phase, d_cl, d_ca, d_dl, d_da;
};
// KKSCompositionRHSStrategy.cc:72-154, 210-373, 377-441
class KKSCompositionRHSStrategy : public CompositionRHSStrategy
{
 public:
   KKSCompositionRHSStrategy(const ampe_rhs_config& cfg, int conc_scratch_id, int phase_scratch_id,
                             int temperature_scratch_id, int conc_l_id, int conc_a_id,
                             int pfm_diffusion_id, int phase_coupling_diffusion_id)
       : d_cfg(cfg), d_conc(conc_scratch_id), d_phase(phase_scratch_id), d_T(temperature_scratch_id),
         d_cl(conc_l_id), d_ca(conc_a_id), d_d0(pfm_diffusion_id), d_dphi(phase_coupling_diffusion_id)
   {
   }
   void setDiffusionCoeff(std::shared_ptr<PatchHierarchy> h, double) override
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto phi = patch->cell<double>(d_phase), T = patch->cell<double>(d_T);
         auto cl = patch->cell<double>(d_cl), ca = patch->cell<double>(d_ca);
         auto d0 = patch->side<double>(d_d0)->pointers(0), dp = patch->side<double>(d_dphi)->pointers(0);
         check(ampe_k_concentration_pfmdiffusion(
                   AMPE_BOX_ARGS(patch), phi->getPointer(), phi->getGhostCellWidth(), d0.data(), 0,
                   T->getPointer(), T->getGhostCellWidth(), d_cfg.D_liquid, d_cfg.Q0_liquid,
                   d_cfg.D_solid, d_cfg.Q0_solid, 8.314472, d_cfg.energy_interp, d_cfg.conc_avg_func,
                   nullptr),
               "CONCENTRATION_PFMDIFFUSION");
         check(ampe_k_set_kks_phase_diffusion(&d_cfg, patch->getBox().lower, patch->getBox().upper,
                                              phi->getPointer(), phi->getGhostCellWidth(),
                                              cl->getPointer(), ca->getPointer(),
                                              cl->getGhostCellWidth(), d0.data(), dp.data(), nullptr),
               "setDiffCoeffForPhaseOnPatch");
      }
   }
   void computeFluxOnPatch(Patch& patch, int flux_id) override
   {
      auto conc = patch.cell<double>(d_conc), phi = patch.cell<double>(d_phase);
      auto flux = patch.side<double>(flux_id);
      auto f = flux->pointers(0);
      auto d0 = patch.side<double>(d_d0)->pointers(0), dp = patch.side<double>(d_dphi)->pointers(0);
      check(ampe_k_concentrationflux(patch.getBox().ndim, patch.getBox().lower, patch.getBox().upper,
                                     patch.getDx(), conc->getPointer(), conc->getGhostCellWidth(),
                                     phi->getPointer(), phi->getGhostCellWidth(), d0.data(), 0,
                                     dp.data(), 0, f.data(), flux->getGhostCellWidth(), nullptr),
            "CONCENTRATIONFLUX");
   }

 private:
   ampe_rhs_config d_cfg;
   int d_conc, d_phase, d_T, d_cl, d_ca, d_d0, d_dphi;
};

// ---- TemperatureRHSStrategy (SimpleTemperatureRHSStrategy.cc:31-95) ----------------------------
class TemperatureRHSStrategy
{
 public:
   TemperatureRHSStrategy(int temperature_scratch_id, int cp_id, double thermal_diffusivity,
                          double latent_heat)
       : d_T(temperature_scratch_id), d_cp(cp_id), d_alpha(thermal_diffusivity), d_L(latent_heat)
   {
   }
   void evaluateRHS(std::shared_ptr<PatchHierarchy> h, int temperature_rhs_id, int dphidt_id)
   {
      for (auto& patch : *h->getPatchLevel(0)) {
         auto T = patch->cell<double>(d_T), cp = patch->cell<double>(d_cp);
         auto rhs = patch->cell<double>(temperature_rhs_id);
         const bool with_phase = dphidt_id > -1;
         const double* pr = with_phase ? patch->cell<double>(dphidt_id)->getPointer() : nullptr;
         const int ngpr = with_phase ? patch->cell<double>(dphidt_id)->getGhostCellWidth() : 0;
         check(ampe_k_computerhstemp(AMPE_BOX_ARGS(patch), patch->getDx(), d_alpha, d_L,
                                     T->getPointer(), T->getGhostCellWidth(), cp->getPointer(),
                                     cp->getGhostCellWidth(), with_phase ? 1 : 0, pr, ngpr,
                                     rhs->getPointer(), rhs->getGhostCellWidth(), nullptr),
               "COMPUTERHSTEMP");
      }
   }

 private:
   int d_T, d_cp;
   double d_alpha, d_L;
};

// ---- block preconditioner solvers ----------------------------------------
// EllipticFACSolver + EllipticFACOps (EllipticFACSolver.cc, EllipticFACOps.h:35): M div(D grad u) + C u
// on the single level, inverted by the device multigrid behind ampe_mg_* (csrc/mg.cu) instead of
// hypre PFMG.  The coefficient setters keep the reference's names; finalizeCoefficients() hands
// them to the device (the reference reads them lazily at solve time).
class EllipticFACSolver
{
 public:
   explicit EllipticFACSolver(std::shared_ptr<PatchHierarchy> h) : d_h(h)
   {
      auto patch = d_h->getPatchLevel(0)->patches.front();
      int n[3];
      for (int d = 0; d < 3; d++) n[d] = patch->getBox().numberCells(d);
      check(ampe_mg_create(patch->getBox().ndim, n, patch->getDx(), 0, &d_mg), "ampe_mg_create");
   }
   virtual ~EllipticFACSolver() { ampe_mg_destroy(d_mg); }
   // EllipticFACSolver::setBoundaries("Mixed", ...) with the deck's Robin coefficients: slope-0 per direction
   void setBoundaries(const int* zero_slope) { check(ampe_mg_set_zero_slope(d_mg, zero_slope), "setBoundaries"); }
   void setM(int m_id) { d_m_id = m_id; }
   void setMConstant(double m) { d_m_id = -1, d_m_const = m; }
   void setCPatchDataId(int c_id) { d_c_id = c_id; }
   void setCConstant(double c) { d_c_id = -1, d_c_const = c; }
   // D = scale * (side data id [+ side data id2])
   void setDPatchDataId(int d_id, int d_id2 = -1, double scale = 1.0) { d_d_id = d_id, d_d_id2 = d_id2, d_d_scale = scale; }
   void setDConstant(double d) { d_d_id = d_d_id2 = -1, d_d_const = d; }
   void finalizeCoefficients()
   {
      auto patch = d_h->getPatchLevel(0)->patches.front();
      const double *m = nullptr, *c = nullptr;
      int ngm = 0, ngc = 0, ngd = 0;
      if (d_m_id >= 0) m = patch->cell<double>(d_m_id)->getPointer(), ngm = patch->cell<double>(d_m_id)->getGhostCellWidth();
      if (d_c_id >= 0) c = patch->cell<double>(d_c_id)->getPointer(), ngc = patch->cell<double>(d_c_id)->getGhostCellWidth();
      std::vector<const double*> d1, d2;
      if (d_d_id >= 0) {
         auto sd = patch->side<double>(d_d_id);
         auto v = sd->pointers(0);
         d1.assign(v.begin(), v.end());
         ngd = sd->getGhostCellWidth();
         if (d_d_id2 >= 0) {
            auto sd2 = patch->side<double>(d_d_id2);
            if (sd2->getGhostCellWidth() != ngd) throw std::runtime_error("EllipticFACSolver: ghost widths of D differ");
            auto v2 = sd2->pointers(0);
            d2.assign(v2.begin(), v2.end());
         }
      }
      check(ampe_mg_set_elliptic(d_mg, m, ngm, d_m_const, c, ngc, d_c_const, d1.empty() ? nullptr : d1.data(),
                                 d2.empty() ? nullptr : d2.data(), ngd, d_d_scale, d_d_const, nullptr),
            "EllipticFACSolver::finalizeCoefficients");
   }
   // solveSystem(u, f): ghost-0 device arrays; zero initial guess, fixed number of V-cycles
   bool solveSystem(double* u, const double* f, int ncycles)
   {
      check(ampe_mg_solve(d_mg, f, u, ncycles, 0, nullptr), "EllipticFACSolver::solveSystem");
      return true;
   }
   ampe_mg* levelSolver() const { return d_mg; }

 protected:
   std::shared_ptr<PatchHierarchy> d_h;
   ampe_mg* d_mg = nullptr;
   int d_m_id = -1, d_c_id = -1, d_d_id = -1, d_d_id2 = -1;
   double d_m_const = 1.0, d_c_const = 1.0, d_d_const = 0.0, d_d_scale = 1.0;
};

// PhaseFACSolver / PhaseFACOps::setOperatorCoefficients (PhaseFACOps.cc:33-51)
class PhaseFACSolver : public EllipticFACSolver
{
 public:
   PhaseFACSolver(std::shared_ptr<PatchHierarchy> h, int c_scratch_id) : EllipticFACSolver(h), d_c_scratch_id(c_scratch_id) {}
   // uniform_mobility: the mobility field is known to be one number (computeUniformPhaseMobility,
   // QuatModel.cc:4277-4291): M is then handed over as a constant and not read per cell by the sweeps
   void setOperatorCoefficients(int phase_id, int phase_mobility_id, double epsilon_phase, double gamma,
                                double phase_well_scale, const std::string& phase_well_func_type,
                                const double* uniform_mobility = nullptr)
   {
      if (uniform_mobility)
         setMConstant(*uniform_mobility);
      else
         setM(phase_mobility_id);
      // C to be set after M since it uses M (setC, PhaseFACOps.cc:58-98)
      auto patch = d_h->getPatchLevel(0)->patches.front();
      auto phi = patch->cell<double>(phase_id), m = patch->cell<double>(phase_mobility_id);
      auto c = patch->cell<double>(d_c_scratch_id);
      check(ampe_k_phasefacops_setc(AMPE_BOX_ARGS(patch), phi->getPointer(), phi->getGhostCellWidth(),
                                    m->getPointer(), m->getGhostCellWidth(), gamma, phase_well_scale,
                                    phase_well_func_type.c_str(), c->getPointer(), c->getGhostCellWidth(), nullptr),
            "PhaseFACOps::setC");
      setCPatchDataId(d_c_scratch_id);
      setDConstant(-gamma * epsilon_phase * epsilon_phase);
      finalizeCoefficients();
   }

 private:
   int d_c_scratch_id;
};

// ConcFACSolver / ConcFACOps::setOperatorCoefficients (ConcFACOps.cc:19-50); diffusion_id2 >= 0: the
// sum EBSCompositionRHSStrategy::setDiffusionCoeffForPreconditioner forms (D_l + D_a)
class ConcFACSolver : public EllipticFACSolver
{
 public:
   explicit ConcFACSolver(std::shared_ptr<PatchHierarchy> h) : EllipticFACSolver(h) {}
   void setOperatorCoefficients(double gamma, int diffusion_id, int diffusion_id2, double mobility)
   {
      assert(gamma >= 0. && mobility > 0.);
      setDPatchDataId(diffusion_id, diffusion_id2, -gamma);
      setCConstant(1.);
      setMConstant(mobility);
      finalizeCoefficients();
   }
};

// TemperatureFACSolver::setOperatorCoefficients(m, c, d) (QuatIntegrator.cc:3340-3346)
class TemperatureFACSolver : public EllipticFACSolver
{
 public:
   explicit TemperatureFACSolver(std::shared_ptr<PatchHierarchy> h) : EllipticFACSolver(h) {}
   void setOperatorCoefficients(double m, double c, double d)
   {
      setMConstant(m);
      setCConstant(c);
      setDConstant(d);
      finalizeCoefficients();
   }
};

}  // namespace ampe_host
