// Host side of the fused path: context (device intermediates, lagged face
// data), parameter derivation, kernel dispatch and the extern "C" entry points
// declared in include/ampe_b200.h.  Mirrors the wiring QuatIntegrator does in
// RegisterVariables / evaluateRHSFunction (source/QuatIntegrator.cc:787-985,
// 3134-3295) without SAMRAI.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "rhs_common.cuh"

using namespace ampe;

static thread_local std::string g_err;
int ampe_set_err(int code, const std::string& msg);
static int set_err(int code, const std::string& msg) { return ampe_set_err(code, msg); }
int ampe_set_err(int code, const std::string& msg)
{
   g_err = msg;
   return code;
}
#define CUDA_OK(call)                                                              \
   do {                                                                            \
      cudaError_t e_ = (call);                                                     \
      if (e_ != cudaSuccess)                                                       \
         return set_err(AMPE_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
   } while (0)

static const double R_GAS = 8.314472;  // GASCONSTANT_R_JPKPMOL

#include "ctx_internal.h"

// ---- CALPHAD T-dependent coefficients on the host (uniform T) --------------------------
// G = a + bT + cT ln T + d2 T^2 + d3 T^3 + d4 T^4 + d7 T^7 + dm1/T + dm9/T^9
// (thermodynamic_data/calphadAuNi.dat:1-6)
static double species_fenergy(const ampe_calphad_species& s, double T)
{
   int iv = s.nintervals - 1;
   for (int i = 0; i < s.nintervals; i++)
      if (T >= s.Tc[i] && T < s.Tc[i + 1]) {
         iv = i;
         break;
      }
   if (T < s.Tc[0]) iv = 0;
   const double T2 = T * T, T4 = T2 * T2;
   return s.a[iv] + s.b[iv] * T + s.c[iv] * T * log(T) + s.d2[iv] * T2 + s.d3[iv] * T2 * T +
          s.d4[iv] * T4 + s.d7[iv] * T4 * T2 * T + s.dm1[iv] / T + s.dm9[iv] / (T4 * T4 * T);
}
static double getQ(const double* a, double T) { return a[0] + R_GAS * T * log(a[1]); }

static void fill_calphadT(const ampe_calphad_binary& db, double T, CalphadT& o)
{
   for (int ph = 0; ph < 2; ph++) {
      o.fA[ph] = species_fenergy(db.g[0][ph], T);
      o.fB[ph] = species_fenergy(db.g[1][ph], T);
      for (int k = 0; k < 4; k++) o.L[ph][k] = db.L[ph][k][0] + db.L[ph][k][1] * T;
   }
   o.RT = R_GAS * T;
   o.RTinv = 1.0 / (R_GAS * T);
   for (int sp = 0; sp < 2; sp++)
      for (int ph = 0; ph < 2; ph++) {
         o.qA[sp][ph] = getQ(db.qA[sp][ph], T);
         o.qB[sp][ph] = getQ(db.qB[sp][ph], T);
         for (int k = 0; k < 4; k++) o.qAB[sp][ph][k] = getQ(db.qAB[sp][ph][k], T);
      }
}

// max |dG/RT| of the four atomic mobilities (getDeltaG, CALPHADMobility.cc:158-168) over c in [-1/2, 3/2]
static double mobility_exponent_bound(const CalphadT& t)
{
   double worst = 0.0;
   for (int i = 0; i <= 400; i++) {
      const double c0 = -0.5 + 2.0 * i / 400.0, c1 = 1.0 - c0, dc = c0 - c1;
      for (int sp = 0; sp < 2; sp++)
         for (int ph = 0; ph < 2; ph++) {
            const double* qq = t.qAB[sp][ph];
            const double dG = c0 * t.qA[sp][ph] + c1 * t.qB[sp][ph] +
                              c0 * c1 * (qq[0] + dc * (qq[1] + dc * (qq[2] + dc * qq[3])));
            worst = fmax(worst, fabs(dG * t.RTinv));
         }
   }
   return worst * 1.05;  // sampling margin
}

// ScalarTemperatureStrategy::getCurrentTemperature (ScalarTemperatureStrategy.cc:57-75)
double ampe_uniform_temperature(const ampe_rhs_config& c, double time)
{
   double t = c.T_uniform + c.dtemperaturedt * time;
   if (c.dtemperaturedt < 0. && t < c.target_temperature)
      t = c.target_temperature;
   else if (c.target_temperature > 0.0 && c.dtemperaturedt > 0. && t > c.target_temperature)
      t = c.target_temperature;
   return t;
}

int ampe_derive_params_at(const ampe_rhs_config& c, double T_now, Params& p);
int ampe_derive_params(const ampe_rhs_config& c, Params& p) { return ampe_derive_params_at(c, c.T_uniform, p); }

// T_now: the uniform temperature of this evaluation (T_uniform unless the deck ramps it in time)
int ampe_derive_params_at(const ampe_rhs_config& c, double T_now, Params& p)
{
   memset(&p, 0, sizeof(p));
   if (c.ndim != 2 && c.ndim != 3) return set_err(AMPE_EINVAL, "ndim must be 2 or 3");
   if (c.qlen != 0 && c.qlen != 2 && c.qlen != 4)
      return set_err(AMPE_EINVAL, "qlen must be 0, 2 or 4");
   for (int d = 0; d < c.ndim; d++)
      if (c.n[d] < 4) return set_err(AMPE_EINVAL, "each direction needs at least 4 cells");
   p.ndim = c.ndim;
   p.qlen = c.qlen;
   for (int d = 0; d < 3; d++) p.n[d] = d < c.ndim ? c.n[d] : 1;
   p.ng = (c.conc_rhs_form == AMPE_CONC_CAHN_HILLIARD) ? 2 : 1;
   for (int d = 0; d < c.ndim; d++) p.clamp[d] = c.zero_slope[d] ? 1 : 0;
   if ((p.clamp[0] || p.clamp[1] || p.clamp[2]) && p.ng != 1)
      return set_err(AMPE_EINVAL, "zero-slope boundaries: ghost width 1 models only (not Cahn-Hilliard)");
   if ((p.clamp[0] || p.clamp[1] || p.clamp[2]) && c.symmetry_aware)
      return set_err(AMPE_EINVAL, "zero-slope boundaries with the symmetry-aware path are not supported");
   p.with_phase = c.with_phase;
   p.with_conc = c.with_concentration;
   p.with_T = c.with_unsteady_temperature;
   p.evolve_quat = c.evolve_quat && c.qlen > 0;
   p.flux_type = c.phase_flux_type;
   p.conc_form = c.with_concentration ? c.conc_rhs_form : 0;
   p.free_energy = c.free_energy;
   p.symm = c.symmetry_aware && p.evolve_quat;
   p.modulus_from_cells = c.quat_grad_modulus_from_cells;
   p.libm_trig = (getenv("AMPE_B200_LIBM_TRIG") != nullptr) ? 1 : 0;
   p.energy_interp = c.energy_interp;
   p.conc_interp = c.conc_interp;
   p.diffusion_interp = c.diffusion_interp;
   p.orient_interp1 = c.orient_interp1;
   p.orient_interp2 = c.orient_interp2;
   p.avg_func = c.avg_func;
   p.conc_avg_func = c.conc_avg_func;
   p.grad_floor_type = c.grad_floor_type;
   p.quat_mobility_func = c.quat_mobility_func;
   if (p.flux_type == AMPE_FLUX_ANISOTROPIC && c.qlen == 0)
      return set_err(AMPE_EINVAL, "Phase anisotropy requires quaternion orientation");  // PhaseFluxStrategyFactory.h:26
   if (p.flux_type == AMPE_FLUX_ANISOTROPIC && c.ndim == 3 && (c.qlen != 4 || p.symm))
      return set_err(AMPE_EINVAL, "3D anisotropic phase flux: qlen 4 without the symmetry-aware path");
   if (p.flux_type == AMPE_FLUX_ISOTROPIC && c.ndim != 2)
      return set_err(AMPE_EINVAL, "isotropic stencil is incomplete in 3D (reference stops)");
   if (p.conc_form == AMPE_CONC_EBS && c.free_energy != AMPE_FE_CALPHAD && c.free_energy != AMPE_FE_QUADRATIC)
      return set_err(AMPE_EINVAL, "EBS composition RHS needs the CALPHAD or the quadratic free energy");
   if (p.conc_form == AMPE_CONC_KKS && c.free_energy != AMPE_FE_QUADRATIC && c.free_energy != AMPE_FE_CALPHAD)
      return set_err(AMPE_EINVAL, "KKS composition RHS needs the quadratic or the CALPHAD free energy");
   if ((p.conc_form == AMPE_CONC_EBS || p.conc_form == AMPE_CONC_KKS) && p.with_T)
      return set_err(AMPE_EINVAL, "KKS/EBS with an evolved temperature field is not supported");
   if (p.conc_form == AMPE_CONC_CAHN_HILLIARD && (c.with_phase || c.evolve_quat))
      return set_err(AMPE_EINVAL, "Cahn-Hilliard is a composition-only model");

   const double eps2 = c.epsilon_phase * c.epsilon_phase;
   for (int d = 0; d < c.ndim; d++) {
      const double h = c.dx[d];
      p.h[d] = h;
      p.dinv[d] = 1.0 / h;
      p.p5inv[d] = 0.5 / h;
      p.p25inv[d] = 0.25 * (1.0 / h);
      p.dinv2[d] = 1.0 / (h * h);
      p.eps2_dinv[d] = eps2 / h;
      if (d < 2) p.iso_dinv[d] = (1.0 / 12.0) * eps2 / h;
      p.ch_dinv2[d] = p.dinv[d] * p.dinv[d];
      p.ch_mdinv[d] = c.ch_mobility * p.dinv[d];
   }
   p.epsilon_phase = c.epsilon_phase;
   p.nu = c.epsilon_anisotropy;
   p.knumber = c.knumber;
   p.phi_well_scale = c.phi_well_scale;
   p.phi_mobility = c.phi_mobility;
   p.misorientation_factor = 2.0 * c.H_parameter;
   p.epsilonq2_half = 0.5 * c.epsilon_q * c.epsilon_q;
   p.epsq2 = c.epsilon_q * c.epsilon_q;
   p.floor2 = c.quat_grad_floor * c.quat_grad_floor;
   p.max_normi = 1.0 / c.quat_grad_floor;
   p.quat_mobility = c.quat_mobility;
   p.min_quat_mobility = c.min_quat_mobility;
   p.quat_mobility_alt = c.quat_mobility_alt_scale;
   p.T_uniform = T_now;
   p.thermal_diffusivity = c.thermal_diffusivity;
   p.latent_heat = c.latent_heat;
   p.cp = c.cp;
   p.latent_over_cp = c.latent_heat / c.cp;
   p.meltingT = c.meltingT;
   // computerhsbiaswell: pi = 4.*atan(1.) is REAL*4 (2d/quatrhs.m4:834)
   p.bias_coeff = c.bias_well_alpha / (double)(4.f * atanf(1.f));
   p.bias_gamma = c.bias_well_gamma;
   // computerhsdeltatemperature: alpha = latentheat/tm (2d/quatrhs.m4:919)
   if (c.free_energy == AMPE_FE_DELTAT) {
      if (!(c.meltingT > 0.)) return set_err(AMPE_EINVAL, "the linear free energy needs meltingT > 0");
      p.deltaT_alpha = c.latent_heat / c.meltingT;
   }
   p.conc_mobility = c.conc_mobility;
   p.ch_ca = c.ch_ca;
   p.ch_cb = c.ch_cb;
   p.ch_well_scale = c.ch_well_scale;
   p.ch_kappa = c.ch_kappa;
   const double T = T_now;
   p.quad_A[0] = c.quad_A_l;
   p.quad_A[1] = c.quad_A_s;
   p.quad_ceq[0] = c.quad_Ceq_l + (T - c.quad_Tref) * c.quad_m_l;
   p.quad_ceq[1] = c.quad_Ceq_s + (T - c.quad_Tref) * c.quad_m_s;
   if (c.free_energy == AMPE_FE_QUADRATIC) {
      p.quad_rla = c.quad_A_l / c.quad_A_s;
      p.quad_ral = c.quad_A_s / c.quad_A_l;
   }
   if (p.conc_form == AMPE_CONC_KKS || (p.conc_form == AMPE_CONC_EBS && c.free_energy == AMPE_FE_QUADRATIC)) {
      // concentration_pfmdiffusion (3d/concentrationdiffusion.m4:43-75) / concentration_pfmdiffusion_of_temperature
      // (2d/concentrationdiffusion.m4:382-396) for uniform T
      const double q0l = c.Q0_liquid / R_GAS, q0s = c.Q0_solid / R_GAS;
      const double invT = 2.0 / (T + T);
      p.D_liquid = c.D_liquid * exp(-q0l * invT);
      p.D_solid = c.D_solid * exp(-q0s * invT);
   }
   p.inv_vm_l = 1.e-6 / c.vm_liquid;
   p.inv_vm_a = 1.e-6 / c.vm_solid;
   p.newton_max_its = c.newton_max_its;
   p.newton_tol = c.newton_tol;
   p.newton_alpha = c.newton_alpha;
   if (c.free_energy == AMPE_FE_CALPHAD) {
      fill_calphadT(c.calphad, T, p.ct);
      // the face mobilities evaluate exp(dG/RT) without a range clamp (fastmath.cuh exp_fast<false>): bound the
      // exponent over every concentration a face average can take, c in [-1/2, 3/2]
      if (p.conc_form == AMPE_CONC_EBS && mobility_exponent_bound(p.ct) > 690.0)
         return set_err(AMPE_EINVAL, "CALPHAD mobility parameters: |dG/RT| exceeds 690 for c in [-1/2, 3/2]");
   }
   return AMPE_OK;
}

// ---- quaternion symmetry table (setqr, quat.f:165-286) ----------------------------------
static int upload_qr_table(ampe_rhs_ctx* c)
{
   static const int raw[48][4] = {
       {1, 0, 0, 0},    {0, 1, 0, 0},    {0, 0, 1, 0},    {0, 0, 0, 1},    {-1, 0, 0, 0},
       {0, -1, 0, 0},   {0, 0, -1, 0},   {0, 0, 0, -1},   {1, 1, 0, 0},    {1, 0, 1, 0},
       {1, 0, 0, 1},    {0, 1, 1, 0},    {0, 1, 0, 1},    {0, 0, 1, 1},    {-1, 1, 0, 0},
       {-1, 0, 1, 0},   {-1, 0, 0, 1},   {0, -1, 1, 0},   {0, -1, 0, 1},   {0, 0, -1, 1},
       {1, -1, 0, 0},   {1, 0, -1, 0},   {1, 0, 0, -1},   {0, 1, -1, 0},   {0, 1, 0, -1},
       {0, 0, 1, -1},   {-1, -1, 0, 0},  {-1, 0, -1, 0},  {-1, 0, 0, -1},  {0, -1, -1, 0},
       {0, -1, 0, -1},  {0, 0, -1, -1},  {1, 1, 1, 1},    {-1, 1, 1, 1},   {1, -1, 1, 1},
       {1, 1, -1, 1},   {1, 1, 1, -1},   {-1, -1, 1, 1},  {-1, 1, -1, 1},  {-1, 1, 1, -1},
       {1, -1, -1, 1},  {1, -1, 1, -1},  {1, 1, -1, -1},  {1, -1, -1, -1}, {-1, 1, -1, -1},
       {-1, -1, 1, -1}, {-1, -1, -1, 1}, {-1, -1, -1, -1}};
   static const int conj[48] = {1,  6,  7,  8,  5,  2,  3,  4,  21, 22, 23, 30, 31, 32, 27, 28,
                                29, 24, 25, 26, 9,  10, 11, 18, 19, 20, 15, 16, 17, 12, 13, 14,
                                44, 48, 43, 42, 41, 45, 46, 47, 37, 36, 35, 33, 38, 39, 40, 34};
   double qr[48][4];
   for (int n = 0; n < 48; n++) {
      double q[4] = {(double)raw[n][0], (double)raw[n][1], (double)raw[n][2], (double)raw[n][3]};
      const double m = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      const double minv = (m < 1.e-15) ? 0.0 : 1.0 / m;
      for (int k = 0; k < 4; k++) qr[n][k] = q[k] * minv;
   }
   CUDA_OK(cudaMalloc(&c->qr_dev, sizeof(qr)));
   CUDA_OK(cudaMalloc(&c->conj_dev, sizeof(conj)));
   CUDA_OK(cudaMemcpy(c->qr_dev, qr, sizeof(qr), cudaMemcpyHostToDevice));
   CUDA_OK(cudaMemcpy(c->conj_dev, conj, sizeof(conj), cudaMemcpyHostToDevice));
   return AMPE_OK;
}

// ---- kernel dispatch (instantiations live in fused3_*.cu) ----------------------------------
namespace ampe {
template <int ND, int Q>
int dispatch3_runtime(const FusedArgs& A, cudaStream_t st, const char** err);
extern template int dispatch3_runtime<2, 0>(const FusedArgs&, cudaStream_t, const char**);
extern template int dispatch3_runtime<2, 2>(const FusedArgs&, cudaStream_t, const char**);
extern template int dispatch3_runtime<2, 4>(const FusedArgs&, cudaStream_t, const char**);
extern template int dispatch3_runtime<3, 0>(const FusedArgs&, cudaStream_t, const char**);
extern template int dispatch3_runtime<3, 2>(const FusedArgs&, cudaStream_t, const char**);
extern template int dispatch3_runtime<3, 4>(const FusedArgs&, cudaStream_t, const char**);
// compile-time selector sets of the shipped decks; return 1 when they handled the launch
int dispatch3_fixed_dendrite(const FusedArgs& A, cudaStream_t st, const char** err, int* rc);
int dispatch3_fixed_auni2d(const FusedArgs& A, cudaStream_t st, const char** err, int* rc);
int dispatch3_fixed_3d(const FusedArgs& A, cudaStream_t st, const char** err, int* rc);
}  // namespace ampe

template <int ND>
static int dispatch_q(const FusedArgs& A, cudaStream_t st, int* nlaunch)
{
   const char* err = nullptr;
   int rc = AMPE_EINVAL;
   *nlaunch = 1;
   // AMPE_B200_GENERIC=1 forces the runtime-selector instantiation (tests compare both)
   const bool generic_only = A.force_generic != 0;
   bool done = false;
   if (!generic_only) {
      if (ND == 2)
         done = dispatch3_fixed_dendrite(A, st, &err, &rc) || dispatch3_fixed_auni2d(A, st, &err, &rc);
      else {
         const int n = dispatch3_fixed_3d(A, st, &err, &rc);
         done = n != 0;
         if (n > 1) *nlaunch = n;
      }
   }
   if (!done) {
      switch (A.p.qlen) {
         case 0: rc = dispatch3_runtime<ND, 0>(A, st, &err); break;
         case 2: rc = dispatch3_runtime<ND, 2>(A, st, &err); break;
         case 4: rc = dispatch3_runtime<ND, 4>(A, st, &err); break;
         default: err = "unsupported qlen";
      }
   }
   if (rc) return set_err(rc, err ? err : "kernel launch failed");
   return AMPE_OK;
}

// ---- context ------------------------------------------------------------------------------
static long long slab_ghosted_cells(const ampe_rhs_ctx* c) { return c->plane * (c->ns + 2LL * c->ng); }

static int alloc_device_arrays(ampe_rhs_ctx* c);
extern "C" int ampe_rhs_destroy(ampe_rhs_ctx* c);

extern "C" int ampe_rhs_create(const ampe_rhs_config* cfg, ampe_rhs_ctx** out)
{
   if (!cfg || !out) return set_err(AMPE_EINVAL, "null argument");
   int ndev = 0;
   if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      return set_err(AMPE_ENOGPU, "no CUDA device: libampe_b200 has no CPU fallback");
   ampe_rhs_ctx* c = new ampe_rhs_ctx;
   c->cfg = *cfg;
   c->generic_only = getenv("AMPE_B200_GENERIC") != nullptr;
   c->split3d = getenv("AMPE_B200_SPLIT3D") != nullptr;
   int rc = ampe_derive_params(*cfg, c->p);
   if (rc) {
      delete c;
      return rc;
   }
   const Params& p = c->p;
   c->ng = p.ng;
   c->ns = (p.ndim == 3) ? p.n[2] : p.n[1];
   c->plane = (p.ndim == 3) ? (long long)p.n[0] * p.n[1] : p.n[0];
   c->ncell = (long long)p.n[0] * p.n[1] * p.n[2];
   memset(&c->halo_lo, 0, sizeof(c->halo_lo));
   memset(&c->halo_hi, 0, sizeof(c->halo_hi));
   memset(&c->dev_y, 0, sizeof(c->dev_y));
   memset(&c->dev_ydot, 0, sizeof(c->dev_ydot));
   // any failure below releases what was allocated so far and hands no context back
   rc = alloc_device_arrays(c);
   if (rc) {
      ampe_rhs_destroy(c);
      *out = nullptr;
      return rc;
   }
   *out = c;
   return AMPE_OK;
}

static int alloc_device_arrays(ampe_rhs_ctx* c)
{
   const Params& p = c->p;
   const ampe_rhs_config* cfg = &c->cfg;
   int rc = AMPE_OK;
   const size_t gb = (size_t)slab_ghosted_cells(c) * sizeof(double);
   if (p.conc_form == AMPE_CONC_KKS || p.conc_form == AMPE_CONC_EBS) {
      CUDA_OK(cudaMalloc(&c->cl, gb));
      CUDA_OK(cudaMalloc(&c->ca, gb));
      CUDA_OK(cudaMalloc(&c->cl_ref, gb));
      CUDA_OK(cudaMalloc(&c->ca_ref, gb));
      CUDA_OK(cudaMalloc(&c->nfail, sizeof(int)));
      CUDA_OK(cudaMemset(c->nfail, 0, sizeof(int)));
      if (p.free_energy == AMPE_FE_CALPHAD) CUDA_OK(cudaMalloc(&c->df, (size_t)c->ncell * sizeof(double)));
   }
   const size_t lagb = (size_t)c->plane * (c->ns + 1) * sizeof(double);
   if (cfg->lag_quat_sidegrad) {
      for (int d = 0; d < p.ndim; d++) {
         if (p.evolve_quat) CUDA_OK(cudaMalloc(&c->lagN[d], lagb));
         if (p.conc_form == AMPE_CONC_KKS || p.conc_form == AMPE_CONC_EBS) {
            CUDA_OK(cudaMalloc(&c->lagD0[d], lagb));
            CUDA_OK(cudaMalloc(&c->lagD1[d], lagb));
         }
      }
   }
   if (p.symm) {
      rc = upload_qr_table(c);
      if (rc) return rc;
      for (int d = 0; d < p.ndim; d++) {
         const size_t ib = (size_t)slab_ghosted_cells(c) * sizeof(int);
         CUDA_OK(cudaMalloc(&c->iq[d], ib));
         CUDA_OK(cudaMemset(c->iq[d], 0, ib));  // 0 = "no rotation known yet" (quatfindsymm starts from 1)
      }
   }
   return AMPE_OK;
}

extern "C" int ampe_rhs_destroy(ampe_rhs_ctx* c)
{
   if (!c) return AMPE_OK;
   cudaFree(c->cl);
   cudaFree(c->ca);
   cudaFree(c->cl_ref);
   cudaFree(c->ca_ref);
   cudaFree(c->nfail);
   cudaFree(c->df);
   cudaFree(c->partials);
   cudaFree(c->red_out);
   cudaFree(c->qr_dev);
   cudaFree(c->conj_dev);
   for (int d = 0; d < 3; d++) {
      cudaFree(c->iq[d]);
      cudaFree(c->lagN[d]);
      cudaFree(c->lagD0[d]);
      cudaFree(c->lagD1[d]);
   }
   if (c->have_dev) {
      cudaFree(c->dev_y.phase);
      cudaFree(c->dev_y.quat);
      cudaFree(c->dev_y.conc);
      cudaFree(c->dev_y.temperature);
      cudaFree(c->dev_ydot.phase);
      cudaFree(c->dev_ydot.quat);
      cudaFree(c->dev_ydot.conc);
      cudaFree(c->dev_ydot.temperature);
   }
   for (int i = 0; i < 3; i++)
      if (c->ev_t[i]) cudaEventDestroy(c->ev_t[i]);
   if (c->grain_label) cudaFree(c->grain_label);
   if (c->own_stream) {
      cudaStreamDestroy(c->own_stream);
      cudaStreamDestroy(c->k_stream);
      cudaStreamDestroy(c->out_stream);
      for (int j = 0; j < AMPE_MAX_HOST_CHUNKS; j++) {
         cudaEventDestroy(c->ev_in[j]);
         cudaEventDestroy(c->ev_k[j]);
      }
   }
   delete c;
   return AMPE_OK;
}

extern "C" int ampe_rhs_nghosts(const ampe_rhs_ctx* c) { return c ? c->ng : AMPE_EINVAL; }

// copy a ghost-0 array into a slab-ghosted one; ghost planes = periodic wrap (single rank)
template <typename T>
static int fill_slab_ghosted(ampe_rhs_ctx* c, T* dst, const T* src, cudaStream_t st)
{
   const long long pl = c->plane;
   const int ng = c->ng, ns = c->ns;
   CUDA_OK(cudaMemcpyAsync(dst + ng * pl, src, sizeof(T) * pl * ns, cudaMemcpyDeviceToDevice, st));
   // ghost planes: the opposite interior planes (periodic) or the adjacent ones (zero slope, ghost width 1)
   const bool clamp = c->p.clamp[c->p.ndim - 1] != 0;
   CUDA_OK(cudaMemcpyAsync(dst, src + (clamp ? 0 : (long long)(ns - ng) * pl), sizeof(T) * pl * ng,
                           cudaMemcpyDeviceToDevice, st));
   CUDA_OK(cudaMemcpyAsync(dst + (long long)(ng + ns) * pl, src + (clamp ? (long long)(ns - ng) * pl : 0),
                           sizeof(T) * pl * ng, cudaMemcpyDeviceToDevice, st));
   return AMPE_OK;
}

extern "C" int ampe_rhs_set_ref_concentrations(ampe_rhs_ctx* c, const double* cl_ref,
                                               const double* ca_ref, void* stream)
{
   if (!c || !c->cl) return set_err(AMPE_EINVAL, "context has no phase concentrations");
   cudaStream_t st = (cudaStream_t)stream;
   if (cl_ref && ca_ref) {
      if (c->cfg.nranks > 1)
         return set_err(AMPE_EINVAL,
                        "several ranks: use ampe_rhs_set_ref_concentrations_slab (ghost planes come from the "
                        "neighbours), or pass NULL (copy the last c_l, c_a incl. ghost planes)");
      int rc = fill_slab_ghosted(c, c->cl_ref, cl_ref, st);
      if (rc) return rc;
      rc = fill_slab_ghosted(c, c->ca_ref, ca_ref, st);
      if (rc) return rc;
   } else {
      // resetRefPhaseConcentrations: whole-array copy incl. ghosts (QuatModel.cc:5218-5231)
      const size_t gb = (size_t)slab_ghosted_cells(c) * sizeof(double);
      CUDA_OK(cudaMemcpyAsync(c->cl_ref, c->cl, gb, cudaMemcpyDeviceToDevice, st));
      CUDA_OK(cudaMemcpyAsync(c->ca_ref, c->ca, gb, cudaMemcpyDeviceToDevice, st));
   }
   c->have_ref = true;
   return AMPE_OK;
}

// multi-rank warm start: the reference value on ghost planes too (slab-ghosted arrays)
extern "C" int ampe_rhs_set_ref_concentrations_ghosted(ampe_rhs_ctx* c, const double* cl_ref_g,
                                                       const double* ca_ref_g, void* stream)
{
   if (!c || !c->cl) return set_err(AMPE_EINVAL, "context has no phase concentrations");
   const size_t gb = (size_t)slab_ghosted_cells(c) * sizeof(double);
   cudaStream_t st = (cudaStream_t)stream;
   CUDA_OK(cudaMemcpyAsync(c->cl_ref, cl_ref_g, gb, cudaMemcpyDeviceToDevice, st));
   CUDA_OK(cudaMemcpyAsync(c->ca_ref, ca_ref_g, gb, cudaMemcpyDeviceToDevice, st));
   c->have_ref = true;
   return AMPE_OK;
}

extern "C" int ampe_rhs_set_symmetry_rotations(ampe_rhs_ctx* c, const int* const* iqrot,
                                               void* stream)
{
   if (!c || !c->p.symm) return set_err(AMPE_EINVAL, "context is not symmetry aware");
   if (c->cfg.nranks > 1)
      return set_err(AMPE_EINVAL, "several ranks: use ampe_rhs_set_symmetry_rotations_slab (the ghost planes of "
                                  "the indices come from the neighbours)");
   for (int d = 0; d < c->p.ndim; d++) {
      int rc = fill_slab_ghosted(c, c->iq[d], iqrot[d], (cudaStream_t)stream);
      if (rc) return rc;
   }
   return AMPE_OK;
}

extern "C" int ampe_rhs_set_halo(ampe_rhs_ctx* c, const ampe_rhs_fields* lo,
                                 const ampe_rhs_fields* hi)
{
   if (!c) return set_err(AMPE_EINVAL, "null context");
   if (lo && hi) {
      c->halo_lo = *lo;
      c->halo_hi = *hi;
      c->have_halo = true;
   } else {
      c->have_halo = false;
   }
   return AMPE_OK;
}

static Field make_field(const ampe_rhs_ctx* c, const double* base, const double* lo,
                        const double* hi, int depth)
{
   Field f;
   f.base = base;
   f.comp = c->ncell;
   if (c->have_halo && lo && hi) {
      f.lo = lo;
      f.hi = hi;
      f.hcomp = (long long)c->ng * c->plane;
   } else {
      if (c->p.clamp[c->p.ndim - 1]) {
         // zero-slope boundary along the slab axis (ghost width 1): the ghost plane is the adjacent interior plane
         f.lo = base;
         f.hi = base ? base + (long long)(c->ns - c->ng) * c->plane : nullptr;
      } else {
         // periodic wrap inside this rank: the ghost planes ARE the opposite interior planes
         f.lo = base ? base + (long long)(c->ns - c->ng) * c->plane : nullptr;
         f.hi = base;
      }
      f.hcomp = c->ncell;
   }
   (void)depth;
   return f;
}

// One evaluation restricted to slab-axis ranges: the KKS pre-pass runs on kks[r] (slab indices
// incl. ghost planes, [-ng, ns+ng)), the fused / Cahn-Hilliard kernel on cells[r] ([0, ns)).
// `first` resets the launch counter, `last` marks the lagged data valid.
struct Ranges {
   int n = 0;
   int r[3][2];
   void add(int b, int e)
   {
      if (e > b) r[n][0] = b, r[n++][1] = e;
   }
};

static int eval_ranges(ampe_rhs_ctx* c, const ampe_rhs_fields* y, const ampe_rhs_fields* ydot,
                       int fd_flag, cudaStream_t st, const Ranges& kks, const Ranges& cells,
                       bool first, bool last, double* energy_partials = nullptr)
{
   if (!c || !y || !ydot) return set_err(AMPE_EINVAL, "null argument");
   const Params& p = c->p;
   const bool en = energy_partials != nullptr;  // energy diagnostics: ydot is not written
   if (p.with_phase && (!y->phase || (!en && !ydot->phase))) return set_err(AMPE_EINVAL, "phase missing");
   if (p.qlen > 0 && !y->quat) return set_err(AMPE_EINVAL, "quat missing");
   if (p.evolve_quat && !en && !ydot->quat) return set_err(AMPE_EINVAL, "ydot quat missing");
   if (p.with_conc && (!y->conc || (!en && !ydot->conc))) return set_err(AMPE_EINVAL, "conc missing");
   if (p.with_T && (!y->temperature || (!en && !ydot->temperature)))
      return set_err(AMPE_EINVAL, "temperature missing");
   if (first) c->launches = 0;

   // QuatIntegrator.cc:3189
   const bool recompute = (fd_flag == 0) || !c->cfg.lag_quat_sidegrad;
   const bool has_lag_data = c->lagN[0] || c->lagD0[0];
   const bool use_lag = !recompute && has_lag_data && !en;
   if (use_lag && !c->lag_valid)
      return set_err(AMPE_EINVAL, "fd_flag=1 before any fd_flag=0 evaluation (no lagged data)");

   const bool timed = c->time_kernels && first && last && !en;
   // ---- Cahn-Hilliard: composition-only model -------------------------------------------
   if (p.conc_form == AMPE_CONC_CAHN_HILLIARD) {
      if (timed) {
         CUDA_OK(cudaEventRecord(c->ev_t[0], st));
         CUDA_OK(cudaEventRecord(c->ev_t[1], st));
      }
      ChArgs A;
      A.p = p;
      A.conc = make_field(c, y->conc, c->halo_lo.conc, c->halo_hi.conc, 1);
      A.out_c = ydot->conc;
      for (int r = 0; r < cells.n; r++) {
         A.s_begin = cells.r[r][0];
         A.s_end = cells.r[r][1];
         const long long total = c->plane * (A.s_end - A.s_begin);
         const int blocks = (int)((total + 255) / 256);
         if (p.ndim == 2)
            ch_kernel<2><<<blocks, 256, 0, st>>>(A);
         else
            ch_kernel<3><<<blocks, 256, 0, st>>>(A);
         CUDA_OK(cudaGetLastError());
         c->launches++;
      }
      if (timed) CUDA_OK(cudaEventRecord(c->ev_t[2], st));
      return AMPE_OK;
   }

   Field fphi = make_field(c, y->phase, c->halo_lo.phase, c->halo_hi.phase, 1);
   Field fT = make_field(c, y->temperature, c->halo_lo.temperature, c->halo_hi.temperature, 1);
   Field fq = make_field(c, y->quat, c->halo_lo.quat, c->halo_hi.quat, p.qlen);
   Field fc = make_field(c, y->conc, c->halo_lo.conc, c->halo_hi.conc, 1);

   if (timed) CUDA_OK(cudaEventRecord(c->ev_t[0], st));
   // ---- per-cell KKS solve on the slab and its ghost planes ------------------------------
   if (p.conc_form == AMPE_CONC_KKS || p.conc_form == AMPE_CONC_EBS) {
      if (p.free_energy == AMPE_FE_CALPHAD && !c->have_ref)
         return set_err(AMPE_EINVAL, "ampe_rhs_set_ref_concentrations must be called first");
      KksArgs K;
      K.p = p;
      K.phi = fphi;
      K.conc = fc;
      K.cl_ref = c->cl_ref;
      K.ca_ref = c->ca_ref;
      K.cl = c->cl;
      K.ca = c->ca;
      K.df = c->df;
      K.nfail = c->nfail;
      for (int r = 0; r < kks.n; r++) {
         K.s_begin = kks.r[r][0];
         K.s_end = kks.r[r][1];
         const long long total = c->plane * (K.s_end - K.s_begin);
         const int blocks = (int)((total + 255) / 256);
         if (p.ndim == 2)
            kks_kernel<2><<<blocks, 256, 0, st>>>(K);
         else
            kks_kernel<3><<<blocks, 256, 0, st>>>(K);
         CUDA_OK(cudaGetLastError());
         c->launches++;
      }
   }

   if (timed) CUDA_OK(cudaEventRecord(c->ev_t[1], st));
   FusedArgs A;
   A.p = p;
   A.phi = fphi;
   A.T = fT;
   A.q = fq;
   A.conc = fc;
   A.cl = c->cl;
   A.ca = c->ca;
   A.qr = c->qr_dev;
   A.conj = c->conj_dev;
   for (int d = 0; d < 3; d++) {
      A.iq[d] = c->iq[d];
      A.lagN[d] = c->lagN[d];
      A.lagD0[d] = c->lagD0[d];
      A.lagD1[d] = c->lagD1[d];
   }
   A.out_phi = ydot->phase;
   A.out_q = ydot->quat;
   A.out_c = ydot->conc;
   A.out_T = ydot->temperature;
   A.force_generic = (c->generic_only || en) ? 1 : 0;
   A.energy_partials = energy_partials;
   A.wrap_slab = c->have_halo ? 0 : 1;
   A.df = c->df;
   A.split3d = c->split3d ? 1 : 0;
   A.wait_flag[0] = c->wait_flag[0];
   A.wait_flag[1] = c->wait_flag[1];
   A.wait_epoch = c->wait_epoch;
   A.use_lag = use_lag ? 1 : 0;
   A.write_lag = (recompute && c->cfg.lag_quat_sidegrad && !en) ? 1 : 0;
   for (int r = 0; r < cells.n; r++) {
      A.s_begin = cells.r[r][0];
      A.s_end = cells.r[r][1];
      int nlaunch = 1;
      int rc = (p.ndim == 2) ? dispatch_q<2>(A, st, &nlaunch) : dispatch_q<3>(A, st, &nlaunch);
      if (rc) return rc;
      c->launches += nlaunch;
   }
   if (timed) CUDA_OK(cudaEventRecord(c->ev_t[2], st));
   if (last && A.write_lag) c->lag_valid = true;
   return AMPE_OK;
}

// energy diagnostics through the fused kernel family: KKS pre-pass + energy_tile_kernel
int ampe_launch_energy(ampe_rhs_ctx* c, const ampe_rhs_fields* y, cudaStream_t st, long long* nblocks)
{
   const Params& p = c->p;
   const int ns = c->ns, ng = c->ng;
   const int TY = (p.ndim == 2) ? AMPE_T2Y : AMPE_T3Y, TZ = (p.ndim == 2) ? 1 : AMPE_T3Z;
   long long nb = (p.n[0] + 31) / 32;
   if (p.ndim == 2)
      nb *= (ns + TY - 1) / TY;
   else
      nb *= (long long)((p.n[1] + TY - 1) / TY) * ((ns + TZ - 1) / TZ);
   // one allocator for the per-block partial sums (vecops.cu): the capacity is counted in blocks of
   // 8 doubles everywhere, so an energy evaluation can never leave a buffer the reductions overrun
   {
      int rc = ampe_ensure_scratch(c, nb);
      if (rc) return rc;
   }
   *nblocks = nb;
   Ranges kks, cells;
   kks.add(-ng, ns + ng);
   cells.add(0, ns);
   ampe_rhs_fields none;
   memset(&none, 0, sizeof(none));
   const int saved = c->launches;
   int rc = eval_ranges(c, y, &none, 0, st, kks, cells, false, false, c->partials);
   c->launches = saved;
   return rc;
}

// part: 0 = everything, 1 = interior (no ghost plane needed), 2 = boundary planes
// the uniform temperature of an evaluation at `time` (decks that ramp it: setTemperatureField at
// QuatIntegrator.cc:3191 runs inside every evaluateRHSFunction); T-dependent parameters follow
static int set_time(ampe_rhs_ctx* c, double time)
{
   if (c->cfg.dtemperaturedt == 0.0) return AMPE_OK;
   const double T = ampe_uniform_temperature(c->cfg, time);
   if (T == c->p.T_uniform) return AMPE_OK;
   return ampe_derive_params_at(c->cfg, T, c->p);
}

static int eval_part(ampe_rhs_ctx* c, double time, const ampe_rhs_fields* y,
                     const ampe_rhs_fields* ydot, int fd_flag, cudaStream_t st, int part)
{
   if (!c) return set_err(AMPE_EINVAL, "null argument");
   {
      int rc = set_time(c, time);
      if (rc) return rc;
   }
   const int ns = c->ns, ng = c->ng;
   if (part != 0 && ns < 4 * ng) return set_err(AMPE_EINVAL, "slab too thin to split");
   Ranges kks, cells;
   if (part == 0) {
      kks.add(-ng, ns + ng);
      cells.add(0, ns);
   } else if (part == 1) {
      kks.add(0, ns);
      cells.add(ng, ns - ng);
   } else {
      kks.add(-ng, 0);
      kks.add(ns, ns + ng);
      cells.add(0, ng);
      cells.add(ns - ng, ns);
   }
   return eval_ranges(c, y, ydot, fd_flag, st, kks, cells, part != 2, part != 1);
}

extern "C" int ampe_rhs_eval(ampe_rhs_ctx* c, double time, const ampe_rhs_fields* y,
                             const ampe_rhs_fields* ydot, int fd_flag, void* stream)
{
   return eval_part(c, time, y, ydot, fd_flag, (cudaStream_t)stream, 0);
}
extern "C" int ampe_rhs_eval_interior(ampe_rhs_ctx* c, double time, const ampe_rhs_fields* y,
                                      const ampe_rhs_fields* ydot, int fd_flag, void* stream)
{
   return eval_part(c, time, y, ydot, fd_flag, (cudaStream_t)stream, 1);
}
extern "C" int ampe_rhs_eval_boundary(ampe_rhs_ctx* c, double time, const ampe_rhs_fields* y,
                                      const ampe_rhs_fields* ydot, int fd_flag, void* stream)
{
   return eval_part(c, time, y, ydot, fd_flag, (cudaStream_t)stream, 2);
}

extern "C" int ampe_rhs_get_phase_concentrations(ampe_rhs_ctx* c, double** cl, double** ca)
{
   if (!c || !c->cl) return set_err(AMPE_EINVAL, "context has no phase concentrations");
   // interior planes of the slab-ghosted arrays are contiguous ghost-0 arrays
   *cl = c->cl + (long long)c->ng * c->plane;
   *ca = c->ca + (long long)c->ng * c->plane;
   return AMPE_OK;
}

extern "C" int ampe_rhs_copy_phase_concentrations(ampe_rhs_ctx* c, double* cl, double* ca,
                                                  void* stream)
{
   if (!c || !c->cl) return set_err(AMPE_EINVAL, "context has no phase concentrations");
   const size_t nb = (size_t)c->ncell * sizeof(double);
   cudaStream_t st = (cudaStream_t)stream;
   CUDA_OK(cudaMemcpyAsync(cl, c->cl + (long long)c->ng * c->plane, nb, cudaMemcpyDeviceToDevice, st));
   CUDA_OK(cudaMemcpyAsync(ca, c->ca + (long long)c->ng * c->plane, nb, cudaMemcpyDeviceToDevice, st));
   return AMPE_OK;
}

extern "C" int ampe_rhs_newton_failures(ampe_rhs_ctx* c, void* stream)
{
   if (!c) return AMPE_EINVAL;
   if (!c->nfail) return 0;
   int h = 0;
   cudaStream_t st = (cudaStream_t)stream;
   if (cudaMemcpyAsync(&h, c->nfail, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess)
      return AMPE_ECUDA;
   if (cudaStreamSynchronize(st) != cudaSuccess) return AMPE_ECUDA;
   cudaMemsetAsync(c->nfail, 0, sizeof(int), st);
   return h;
}

extern "C" int ampe_rhs_last_launch_count(const ampe_rhs_ctx* c) { return c ? c->launches : 0; }

// Per-kernel device times of whole-slab evaluations (bench.py's roofline of the dominant kernel): CUDA events on
// the launching stream around the KKS pre-pass and around the fused kernel.
extern "C" int ampe_rhs_set_kernel_timing(ampe_rhs_ctx* c, int on)
{
   if (!c) return set_err(AMPE_EINVAL, "null context");
   if (on && !c->ev_t[0])
      for (int i = 0; i < 3; i++) CUDA_OK(cudaEventCreate(&c->ev_t[i]));
   c->time_kernels = on != 0;
   return AMPE_OK;
}
extern "C" int ampe_rhs_last_kernel_ms(ampe_rhs_ctx* c, double* kks_ms, double* fused_ms)
{
   if (!c || !c->ev_t[0] || !kks_ms || !fused_ms) return set_err(AMPE_EINVAL, "kernel timing is not enabled");
   float a = 0.f, b = 0.f;
   CUDA_OK(cudaEventSynchronize(c->ev_t[2]));
   CUDA_OK(cudaEventElapsedTime(&a, c->ev_t[0], c->ev_t[1]));
   CUDA_OK(cudaEventElapsedTime(&b, c->ev_t[1], c->ev_t[2]));
   *kks_ms = a;
   *fused_ms = b;
   return AMPE_OK;
}

// h != NULL: slab rank with neighbours (halo.cu) -- the ghost planes travel device to device, pushed as soon as
// the chunks that hold this rank's boundary planes have landed
static int eval_host_impl(ampe_rhs_ctx* c, ampe_halo* h, double time, const ampe_rhs_fields* yh,
                          const ampe_rhs_fields* ydh, int fd_flag)
{
   if (!c || !yh || !ydh) return set_err(AMPE_EINVAL, "null argument");
   {
      int rc = set_time(c, time);
      if (rc) return rc;
   }
   const Params& p = c->p;
   const size_t nb = (size_t)c->ncell * sizeof(double);
   if (!c->have_dev) {
      if (p.with_phase) {
         CUDA_OK(cudaMalloc(&c->dev_y.phase, nb));
         CUDA_OK(cudaMalloc(&c->dev_ydot.phase, nb));
      }
      if (p.qlen > 0) {
         CUDA_OK(cudaMalloc(&c->dev_y.quat, nb * p.qlen));
         CUDA_OK(cudaMalloc(&c->dev_ydot.quat, nb * p.qlen));
      }
      if (p.with_conc) {
         CUDA_OK(cudaMalloc(&c->dev_y.conc, nb));
         CUDA_OK(cudaMalloc(&c->dev_ydot.conc, nb));
      }
      if (p.with_T) {
         CUDA_OK(cudaMalloc(&c->dev_y.temperature, nb));
         CUDA_OK(cudaMalloc(&c->dev_ydot.temperature, nb));
      }
      CUDA_OK(cudaStreamCreate(&c->own_stream));
      CUDA_OK(cudaStreamCreate(&c->k_stream));
      CUDA_OK(cudaStreamCreate(&c->out_stream));
      for (int j = 0; j < AMPE_MAX_HOST_CHUNKS; j++) {
         CUDA_OK(cudaEventCreateWithFlags(&c->ev_in[j], cudaEventDisableTiming));
         CUDA_OK(cudaEventCreateWithFlags(&c->ev_k[j], cudaEventDisableTiming));
      }
      c->have_dev = true;
   }
   // Pipeline over chunks of slab planes: the H2D copy of chunk j+1, the kernels of chunk j and
   // the D2H copy of chunk j-1 run concurrently on three streams (PCIe is full duplex, and the
   // copies, not the kernels, bound this entry point).  A plane can be evaluated once its two
   // neighbour planes are on the device; plane 0 and the last plane close the periodic wrap
   // at the end.  Small problems run as one chunk.
   const int ns = c->ns, ng = c->ng;
   const long long pl = c->plane;
   int nfields = (p.with_phase ? 1 : 0) + p.qlen + (p.with_conc ? 1 : 0) + (p.with_T ? 1 : 0);
   const double chunk_target = 16.0e6;  // bytes per chunk (swept on B200: profiles/README.md)
   int nchunk = (int)((double)nb * nfields / chunk_target);
   if (nchunk > AMPE_MAX_HOST_CHUNKS) nchunk = AMPE_MAX_HOST_CHUNKS;
   if (const char* e = getenv("AMPE_B200_HOST_CHUNKS")) nchunk = atoi(e);
   if (nchunk > ns / (4 * ng)) nchunk = ns / (4 * ng);
   if (nchunk > AMPE_MAX_HOST_CHUNKS) nchunk = AMPE_MAX_HOST_CHUNKS;
   if (nchunk < 1) nchunk = 1;
   if (c->have_halo && !h) nchunk = 1;  // caller-owned halo buffers: the caller sequences the exchange itself
   cudaStream_t s_in = c->own_stream, s_k = c->k_stream, s_out = c->out_stream;
   auto h2d_planes = [&](int b, int e) -> int {
      const size_t off = (size_t)b * pl, bytes = (size_t)(e - b) * pl * sizeof(double);
      if (p.with_phase)
         CUDA_OK(cudaMemcpyAsync(c->dev_y.phase + off, yh->phase + off, bytes, cudaMemcpyHostToDevice, s_in));
      for (int m = 0; m < p.qlen; m++)
         CUDA_OK(cudaMemcpyAsync(c->dev_y.quat + m * c->ncell + off, yh->quat + m * c->ncell + off, bytes,
                                 cudaMemcpyHostToDevice, s_in));
      if (p.with_conc)
         CUDA_OK(cudaMemcpyAsync(c->dev_y.conc + off, yh->conc + off, bytes, cudaMemcpyHostToDevice, s_in));
      if (p.with_T)
         CUDA_OK(cudaMemcpyAsync(c->dev_y.temperature + off, yh->temperature + off, bytes,
                                 cudaMemcpyHostToDevice, s_in));
      return AMPE_OK;
   };
   auto d2h_planes = [&](int b, int e) -> int {
      if (e <= b) return AMPE_OK;
      const size_t off = (size_t)b * pl, bytes = (size_t)(e - b) * pl * sizeof(double);
      if (p.with_phase)
         CUDA_OK(cudaMemcpyAsync(ydh->phase + off, c->dev_ydot.phase + off, bytes, cudaMemcpyDeviceToHost, s_out));
      if (p.evolve_quat)
         for (int m = 0; m < p.qlen; m++)
            CUDA_OK(cudaMemcpyAsync(ydh->quat + m * c->ncell + off, c->dev_ydot.quat + m * c->ncell + off,
                                    bytes, cudaMemcpyDeviceToHost, s_out));
      if (p.with_conc)
         CUDA_OK(cudaMemcpyAsync(ydh->conc + off, c->dev_ydot.conc + off, bytes, cudaMemcpyDeviceToHost, s_out));
      if (p.with_T)
         CUDA_OK(cudaMemcpyAsync(ydh->temperature + off, c->dev_ydot.temperature + off, bytes,
                                 cudaMemcpyDeviceToHost, s_out));
      return AMPE_OK;
   };
   int done_to = ng;  // cells [ng, done_to) have been evaluated
   for (int j = 0; j < nchunk; j++) {
      const int b = (int)((long long)ns * j / nchunk), e = (int)((long long)ns * (j + 1) / nchunk);
      int rc = h2d_planes(b, e);
      if (rc) return rc;
      CUDA_OK(cudaEventRecord(c->ev_in[j], s_in));
      CUDA_OK(cudaStreamWaitEvent(s_k, c->ev_in[j], 0));
      Ranges kks, cells;
      const bool fin = (j == nchunk - 1);
      if (h) {
         // my lowest planes are in the first chunk, my highest in the last: push each side as soon as it is on
         // the device; before the closing evaluation wait for the neighbours' planes
         if (j == 0) {
            rc = ampe_halo_push(h, &c->dev_y, 1, s_k);
            if (rc) return rc;
         }
         if (fin) {
            rc = ampe_halo_push(h, &c->dev_y, 2, s_k);
            if (rc) return rc;
            rc = ampe_halo_wait(h, s_k);
            if (rc) return rc;
         }
      }
      if (nchunk == 1) {
         kks.add(-ng, ns + ng);
         cells.add(0, ns);
      } else if (!fin) {
         kks.add(b, e);
         cells.add(done_to, e - ng);
      } else {
         // everything is on the device: ghost planes of the KKS arrays, the rest of the slab,
         // then the first ng planes (their lower neighbours are the last planes)
         kks.add(b, ns + ng);
         kks.add(-ng, 0);
         cells.add(done_to, ns);
         cells.add(0, ng);
      }
      rc = eval_ranges(c, &c->dev_y, &c->dev_ydot, fd_flag, s_k, kks, cells, j == 0, fin);
      if (rc) return rc;
      CUDA_OK(cudaEventRecord(c->ev_k[j], s_k));
      CUDA_OK(cudaStreamWaitEvent(s_out, c->ev_k[j], 0));
      for (int r = 0; r < cells.n; r++) {
         rc = d2h_planes(cells.r[r][0], cells.r[r][1]);
         if (rc) return rc;
      }
      if (!fin) done_to = (e - ng > done_to) ? e - ng : done_to;
   }
   CUDA_OK(cudaStreamSynchronize(s_out));
   // the next call overwrites dev_y on s_in: it must not overtake this call's kernels
   CUDA_OK(cudaStreamSynchronize(s_k));
   return AMPE_OK;
}

extern "C" int ampe_rhs_eval_host(ampe_rhs_ctx* c, double time, const ampe_rhs_fields* yh,
                                  const ampe_rhs_fields* ydh, int fd_flag)
{
   return eval_host_impl(c, nullptr, time, yh, ydh, fd_flag);
}
extern "C" int ampe_rhs_eval_slab_host(ampe_rhs_ctx* c, ampe_halo* h, double time, const ampe_rhs_fields* yh,
                                       const ampe_rhs_fields* ydh, int fd_flag)
{
   if (!h) return set_err(AMPE_EINVAL, "ampe_rhs_eval_slab_host: null halo");
   return eval_host_impl(c, h, time, yh, ydh, fd_flag);
}

extern "C" const char* ampe_last_error(void) { return g_err.c_str(); }
extern "C" const char* ampe_version(void) { return "ampe_b200 0.1 (sm_100a)"; }
extern "C" int ampe_abi_sizeof_config(void) { return (int)sizeof(ampe_rhs_config); }
