// TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of AMPE's phase-field
// right-hand-side evaluation.  Nothing in the product path (ampe_b200/) may
// include, link or call this.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it.
//
// Parity status: the in-tree Fortran/C++ part is pinned by the reference's own
// known-answer tests (tests/testGradQ.cc, tests/testFlux.cc,
// tests/testInterpolationFunctions.cc); the CALPHAD part is pinned by the
// golden vector of tests/CALPHADbinaryEquilibrium/test.input and the property
// tests of tests/testCALPHADbinaryKKS.cc / tests/testCALPHADFunctions.cc.
// Per-cell Newton ITERATES of Thermo4PFM (not in the reference tree, version
// unpinned) and the Quadratic closed forms are "parity unpinned" beyond those.
//
// Every routine follows one reference routine, same loop order, same
// unfused passes, same operation order (compile with -ffp-contract=off).
#pragma once
#include <cstddef>
#include <vector>
#include "../include/ampe_b200.h"

namespace oracle {

// Inclusive cell box.  2D: lo[2]=hi[2]=0 and no ghosts in direction 2.
struct Box {
   int ndim;
   int lo[3], hi[3];
};

// View of a SAMRAI CellData (axis=-1) or one axis of a SideData (axis=0..2):
// Fortran order, i fastest, depth slowest (source/fortran/pdat_m4arrdim*.i).
struct View {
   double* p = nullptr;
   int lo[3] = {0, 0, 0};
   int n[3] = {1, 1, 1};
   size_t comp = 0;  // stride between depth components
   inline double& operator()(int i, int j, int k = 0, int m = 0) const
   {
      return p[(size_t)(i - lo[0]) +
               (size_t)n[0] * ((size_t)(j - lo[1]) + (size_t)n[1] * (size_t)(k - lo[2])) +
               comp * (size_t)m];
   }
   // view starting at depth component m0 (getPointer(m0))
   View at(int m0) const
   {
      View v = *this;
      v.p = p + comp * (size_t)m0;
      return v;
   }
};

struct IView {
   int* p = nullptr;
   int lo[3] = {0, 0, 0};
   int n[3] = {1, 1, 1};
   inline int& operator()(int i, int j, int k = 0) const
   {
      return p[(size_t)(i - lo[0]) +
               (size_t)n[0] * ((size_t)(j - lo[1]) + (size_t)n[1] * (size_t)(k - lo[2]))];
   }
};

View make_view(double* p, const Box& b, int axis, int ng, int depth);
IView make_iview(int* p, const Box& b, int axis, int ng);
size_t view_size(const Box& b, int axis, int ng, int depth);

struct Field {
   std::vector<double> data;
   View v;
   void alloc(const Box& b, int axis, int ng, int depth);
};
struct SideField {
   Field a[3];
   void alloc(const Box& b, int ng, int depth);
};

// ---- functions.f -----------------------------------------------------------
double interp_func(double phi, char type);
double deriv_interp_func(double phi, char type);
double second_deriv_interp_func(double phi, char type);
double well_func(double phi, char type);
double deriv_well_func(double phi, char type);
double second_deriv_well_func(double phi, char type);
double average_func(double a, double b, char type);
double deriv_average_func(double avg_phi, double next_phi, char type);
double interp_ratio_func(double phi, char t1, char t2);
double compl_interp_ratio_func(double phi, char t1, char t2);

// ---- quat.f ----------------------------------------------------------------
void quatmult4(const double* q1, const double* q2, double* q);
void quatmult2(const double* q1, const double* q2, double* q);
void quatconj(const double* q1, double* q2);
void quatsymmrotate(const double* q, int iq, double* q_prime, int qlen);
double eval_grad_normi(double grad_norm2, char floor_type, double floor_grad_norm2,
                       double max_grad_normi);
const double* qr_table4();  // 48x4, normalised (setqr)

// ---- symmetry pre-pass and CVODE projection hook (symmetry.cc) --------------
void quatfindsymm(const double* q1, const double* q2, int* iq_io, double* q2_prime, int qlen);
void quat_symm_rotation(const Box& b, View q, int depth, IView* rot);
void quat_fundamental(const Box& b, View quat, int depth);
void project(const Box& b, int depth, View q, View corr, View err);

// ---- quatrhs.m4 ------------------------------------------------------------
void gradient_flux(const Box& b, const double* h, double epsilon, View phase, View* flux);
void compute_flux_isotropic(const Box& b, const double* h, double epsilon, View phase,
                            View* flux);
void anisotropic_gradient_flux(const Box& b, const double* h, double epsilon, double nu,
                               int knumber, View phase, View quat, int qlen, View* flux);
void computerhspbg(const Box& b, const double* dx, double misorientation_factor,
                   double epsilonq, View* flux, View temp, double phi_well_scale,
                   double eta_well_scale, View phi, View eta, View orient_grad_mod, View rhs,
                   char phi_well_type, char eta_well_type, char energy_interp_type,
                   char orient_interp1, char orient_interp2, int with_orient, int three_phase);
void computerhsdeltatemperature(const Box& b, View phi, View temp, double tm, double latentheat, View rhs,
                                char energy_interp_type);
void phaserhs_fenergy(const Box& b, View fl, View fa, View phi, View rhs, char interp);
void computerhstemp(const Box& b, const double* dx, double thermal_diffusivity,
                    double latent_heat, View temp, View cp, int with_phase, View phi_rhs,
                    View rhs);
void computerhsbiaswell(const Box& b, View phi, View temp, double alpha, double gamma,
                        View te, View rhs);
void laplacian(const Box& b, const double* dx, double coeff, View field, View rhs);

// ---- quatdiffs.m4 / quatgrad.m4 --------------------------------------------
void quatdiffs(const Box& b, int depth, View q, View* diff);
void quatdiffs_symm(const Box& b, int depth, View q, View* diff, IView* iqrot);
void quatgrad_cell(const Box& b, int depth, const double* h, View* diff, View* grad);
void quatgrad_cell_symm(const Box& b, int depth, const double* h, View* diff, View* grad,
                        IView* iqrot);
// grad[a] : side array of axis a, depth ndim*qlen, component index dir*qlen+m
void quatgrad_side(const Box& b, int depth, const double* h, View* diff, View* grad);
void quatgrad_side_symm(const Box& b, int depth, const double* h, View* diff, View* grad,
                        IView* iqrot);
void quatgrad_modulus(const Box& b, int depth, View* grad_cell, View grad_mod);
void quatgrad_modulus_from_sides_compact(const Box& b, int depth, View* grad_side,
                                         View grad_mod);

// ---- quatfacops.m4 ---------------------------------------------------------
void compute_face_coef(const Box& b, int depth, double eps_q, View phi, View temp,
                       double misorientation_factor, View* gq, View* fc,
                       double gradient_floor, char floor_type, char interp1, char interp2,
                       char avg_type);
void compute_flux(const Box& b, int depth, View* fc, View q, const double* h, View* f);
void compute_flux_from_gradq(const Box& b, int depth, View* fc, View* grad_side, View* f);
void compute_lambda_flux(const Box& b, int depth, View* f, View q, const double* h,
                         View lambda);
void add_quat_proj_op(const Box& b, int depth, View mobility, View* f, View q, View lambda,
                      const double* h, View rhs);
void add_quat_op(const Box& b, int depth, View mobility, View* f, const double* h, View rhs);
void correctrhsquatforsymmetry(const Box& b, int depth, const double* dx, View* nonsymm_diff,
                               View* symm_diff, View rhs, View quat, View* facecoeff,
                               View mobility, IView* iqrot);

// ---- mobility.m4 -----------------------------------------------------------
void quatmobility(const Box& b, View phase, View mobility, int ngmobility,
                  double scale_mobility, double min_mobility, char func_type,
                  double alt_scale_factor);

// ---- concentrationrhs.m4 / flux.m4 / concentrationdiffusion.m4 -------------
void add_cahnhilliarddoublewell_flux(const Box& b, const double* dx, View conc,
                                     double mobility, double ca, double cb,
                                     double well_scale, double kappa, View* flux);
void computerhsconcentration(const Box& b, const double* dx, View* flux, double mobility,
                             View rhs);
void concentrationflux(const Box& b, const double* dx, View conc, View phase, View* diffconc,
                       View* dphi, View* flux);
void add_flux(const Box& b, const double* dx, View conc, int ncomp, View* diffconc,
              View* flux);
void concentration_pfmdiffusion(const Box& b, View phi, View* diff, View temp, double d_liquid,
                                double q0_liquid, double d_solid_A, double q0_solid_A,
                                double gas_constant_R, char interp_type, char avg_type);

// ---- Thermo4PFM stand-in (thermo.cc) ---------------------------------------
double xlogx(double x);
double xlogx_deriv(double x);
double xlogx_deriv2(double x);
double calphad_fmix(double l0, double l1, double l2, double l3, double c);
double calphad_fmix_deriv(double l0, double l1, double l2, double l3, double c);
double calphad_fmix_deriv2(double l0, double l1, double l2, double l3, double c);
double calphad_species_fenergy(const ampe_calphad_species& s, double T);
struct CalphadT {  // T-dependent parameters (computeTdependentParameters)
   double fA[2], fB[2], L[2][4];
};
void calphad_Tdep(const ampe_calphad_binary& db, double T, CalphadT& out);
// phase index 0 = liquid, 1 = solid A
double calphad_free_energy(const ampe_calphad_binary& db, double T, double c, int pi);
double calphad_deriv_free_energy(const ampe_calphad_binary& db, double T, double c, int pi);
double calphad_second_deriv_free_energy(const ampe_calphad_binary& db, double T, double c,
                                        int pi);
// KKS: returns Newton iteration count or -1
int calphad_phase_concentrations(const ampe_calphad_binary& db, double T, double c0,
                                 double hphi, double* x, double tol, int max_its,
                                 double alpha);
// two-phase equilibrium (computeCeqT): returns iteration count or -1
int calphad_ceq(const ampe_calphad_binary& db, double T, double* ceq, double tol,
                int max_its, double alpha);
extern const double GASCONSTANT_R_JPKPMOL;

// quadratic (QuadraticFreeEnergyFunctionsBinary)
struct Quadratic {
   double Tref, A[2], Ceq[2], m[2];
};
double quadratic_free_energy(const Quadratic& p, double T, double c, int pi);
double quadratic_deriv_free_energy(const Quadratic& p, double T, double c, int pi);
void quadratic_phase_concentrations(const Quadratic& p, double T, double c0, double hphi,
                                    double* x);

// CALPHADMobility / CompositionStrategyMobilities (in-tree formulas)
double calphad_diffusion_mobility_binary(const ampe_calphad_binary& db, int phase, double c0,
                                         double T);

// ---- driver (restatement of QuatIntegrator::evaluateRHSFunction) -----------
struct Ctx;
Ctx* create(const ampe_rhs_config& cfg);
void destroy(Ctx*);
void set_ref(Ctx*, const double* cl_ref, const double* ca_ref);
void set_rotations(Ctx*, const int* const* iqrot);
// y / ydot: ghost-0 host arrays.  Returns 0, or -3 if a Newton failed.
int eval(Ctx*, double time, const ampe_rhs_fields* y, const ampe_rhs_fields* ydot,
         int fd_flag);
// scalar energy diagnostics (energy.cc): out[8] = total, phi, orient, qint, well, free, 0, 0
int energy(Ctx* c, const ampe_rhs_fields* y, double* out);
void get_phase_concentrations(Ctx*, double* cl, double* ca);

}  // namespace oracle
