/*
 * ampe_b200_kernels.h -- one C symbol per Fortran kernel that AMPE's Strategy classes
 * call on the RHS path (prototypes in source/fortran/QuatFort.h, ConcFort.h), plus the
 * per-patch C++ loops of the CALPHAD / Quadratic / EBS / KKS strategies.
 *
 * Conventions ("lower boundary"):
 *  - same argument ORDER as the reference prototype; scalars by value (the Fortran ABI
 *    passes them by reference); the 2D and 3D variants of the reference (compile-time
 *    NDIM) are one symbol with a leading `ndim` and `ifirst[ndim]`, `ilast[ndim]` arrays;
 *  - every array is a DEVICE pointer in SAMRAI layout: CellData(box, depth, ghosts) is
 *    one block, i fastest, depth slowest, extents box + 2*ng per direction; SideData has
 *    one array per normal axis a with extent +1 in direction a; side index i is the LOWER
 *    face of cell i.  Ghost widths are passed per array exactly like the reference does;
 *  - selector strings become one char (only the first character is significant,
 *    functions.f:28-83);
 *  - trailing `void* stream` (cudaStream_t); return 0 or a negative AMPE_E* code
 *    (unknown selector -> AMPE_EINVAL where the Fortran would `stop`).
 */
#ifndef AMPE_B200_KERNELS_H
#define AMPE_B200_KERNELS_H
#include "ampe_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- quatrhs.m4 --------------------------------------------------------------------- */
/* GRADIENT_FLUX (QuatFort.h:50) */
int ampe_k_gradient_flux(int ndim, const int* ifirst, const int* ilast, const double* dx,
                         double epsilon, const double* phase, int ngphase, double* const* flux,
                         int ngflux, void* stream);
/* COMPUTE_FLUX_ISOTROPIC (QuatFort.h:62), 2D only like the reference */
int ampe_k_compute_flux_isotropic(int ndim, const int* ifirst, const int* ilast, const double* dx,
                                  double epsilon, const double* phase, int ngphase,
                                  double* const* flux, int ngflux, void* stream);
/* ANISOTROPIC_GRADIENT_FLUX (QuatFort.h:75); 2D: nu, 3D: eps4 */
int ampe_k_anisotropic_gradient_flux(int ndim, const int* ifirst, const int* ilast,
                                     const double* dx, double epsilon, double nu, int knumber,
                                     const double* phase, int ngphase, const double* quat, int ngq,
                                     int qlen, double* const* flux, int ngflux, void* stream);
/* COMPUTERHSPBG (QuatFort.h:90-110): the reference's full argument list in its order (character arguments by
 * pointer, as Fortran takes them); eta / eta_well_type / phi_interp_type are read only when three_phase != 0 */
int ampe_k_computerhspbg(int ndim, const int* ifirst, const int* ilast, const double* dx,
                         double misorientation_factor, double epsilonq, double* const* flux,
                         int ngflux, const double* temp, int ngtemp, double phi_well_scale,
                         double eta_well_scale, const double* phi, int ngphi, const double* eta, int ngeta,
                         const double* orient_grad_mod, int ngogm, double* rhs, int ngrhs,
                         const char* phi_well_type, const char* eta_well_type, const char* phi_interp_type,
                         const char* orient_interp_type1, const char* orient_interp_type2, int with_orient,
                         int three_phase, void* stream);
/* PHASERHS_FENERGY (2d/quatrhs.m4:587) */
int ampe_k_phaserhs_fenergy(int ndim, const int* ifirst, const int* ilast, const double* fl,
                            const double* fa, const double* phi, int ngphi, double* rhs, int ngrhs,
                            char energy_interp_type, void* stream);
/* COMPUTERHSTEMP (QuatFort.h:174) */
int ampe_k_computerhstemp(int ndim, const int* ifirst, const int* ilast, const double* dx,
                          double thermal_diffusivity, double latent_heat, const double* temp,
                          int ngtemp, const double* cp, int ngcp, int with_phase,
                          const double* phi_rhs, int ngphi_rhs, double* rhs, int ngrhs,
                          void* stream);
/* COMPUTERHSDELTATEMPERATURE (QuatFort.h:203): DeltaTemperatureFreeEnergyStrategy::addDrivingForce */
int ampe_k_computerhsdeltatemperature(int ndim, const int* ifirst, const int* ilast, const double* phi, int ngphi,
                                      const double* temp, int ngtemp, double tm, double latentheat, double* rhs,
                                      int ngrhs, const char* energy_interp_type, void* stream);
/* COMPUTERHSBIASWELL (QuatFort.h:185) */
int ampe_k_computerhsbiaswell(int ndim, const int* ifirst, const int* ilast, const double* phi,
                              int ngphi, const double* temp, int ngtemp, double alpha, double gamma,
                              const double* te, int ngte, double* rhs, int ngrhs, void* stream);

/* ---- quatdiffs.m4 / quatgrad.m4 ----------------------------------------------------- */
/* QUATDIFFS (QuatFort.h:342) */
int ampe_k_quatdiffs(int ndim, const int* lo, const int* hi, int depth, const double* q, int ngq,
                     double* const* diff, int ngdiff, void* stream);
/* QUATDIFFS_SYMM (QuatFort.h:353) */
int ampe_k_quatdiffs_symm(int ndim, const int* lo, const int* hi, int depth, const double* q,
                          int ngq, double* const* diff, int ngdiff, const int* const* iqrot,
                          int ngiq, void* stream);
/* QUATGRAD_CELL (QuatFort.h:369): grad[a] = CellData depth `depth` */
int ampe_k_quatgrad_cell(int ndim, const int* lo, const int* hi, int depth, const double* h,
                         double* const* diff, int ngdiff, double* const* grad, int nggrad,
                         void* stream);
/* QUATGRAD_CELL_SYMM (QuatFort.h:386) */
int ampe_k_quatgrad_cell_symm(int ndim, const int* lo, const int* hi, int depth, const double* h,
                              double* const* diff, int ngdiff, double* const* grad, int nggrad,
                              const int* const* iqrot, int ngiq, void* stream);
/* QUATGRAD_SIDE (QuatFort.h:408): grad[a] = side array of axis a, depth ndim*depth,
 * component index dir*depth + m (computeQDiffs.cc:46-60) */
int ampe_k_quatgrad_side(int ndim, const int* lo, const int* hi, int depth, const double* h,
                         double* const* diff, int ngdiff, double* const* grad, int nggrad,
                         void* stream);
/* QUATGRAD_SIDE_SYMM (QuatFort.h:452) */
int ampe_k_quatgrad_side_symm(int ndim, const int* lo, const int* hi, int depth, const double* h,
                              double* const* diff, int ngdiff, double* const* grad, int nggrad,
                              const int* const* iqrot, int ngiq, void* stream);
/* QUATGRAD_MODULUS (QuatFort.h:503) */
int ampe_k_quatgrad_modulus(int ndim, const int* lo, const int* hi, int depth,
                            double* const* grad_cell, int nggq, double* grad_mod, int ngm,
                            void* stream);
/* QUATGRAD_MODULUS_FROM_SIDES_COMPACT (QuatFort.h:523) */
int ampe_k_quatgrad_modulus_from_sides_compact(int ndim, const int* lo, const int* hi, int depth,
                                               double* const* grad_side, int nggq,
                                               double* grad_mod, int ngm, void* stream);

/* ---- quatfacops.m4 ------------------------------------------------------------------ */
/* COMPUTE_FACE_COEF2D/3D (QuatFort.h:783, 905) */
int ampe_k_compute_face_coef(int ndim, const int* lo, const int* hi, int depth, double eps_q,
                             const double* phi, int ngp, const double* temp, int ngt,
                             double misorientation_factor, double* const* gq, int nggq,
                             double* const* fc, int ngf, double gradient_floor, char floor_type,
                             char interp_type1, char interp_type2, char avg_type, void* stream);
/* COMPUTE_FLUX2D/3D (QuatFort.h:796, 927); ghost widths instead of explicit lo/hi boxes */
int ampe_k_compute_flux(int ndim, const int* lo, const int* hi, int depth, double* const* fc,
                        int ngfc, const double* q, int ngq, const double* h, double* const* flux,
                        int ngflux, void* stream);
/* COMPUTE_FLUX2D/3D_FROM_GRADQ (QuatFort.h:804, 938): grad_side as in quatgrad_side */
int ampe_k_compute_flux_from_gradq(int ndim, const int* lo, const int* hi, int depth,
                                   double* const* fc, int ngfc, double* const* grad_side,
                                   double* const* flux, int ngflux, void* stream);
/* COMPUTE_LAMBDA_FLUX2D/3D (QuatFort.h:865, 1004) */
int ampe_k_compute_lambda_flux(int ndim, const int* lo, const int* hi, int depth,
                               double* const* flux, int ngflux, const double* q, int ngq,
                               const double* h, double* lambda, int nglambda, void* stream);
/* ADD_QUAT_PROJ_OP2D/3D (QuatFort.h:846, 993) */
int ampe_k_add_quat_proj_op(int ndim, const int* lo, const int* hi, int depth,
                            const double* mobility, int ngmob, double* const* flux, int ngflux,
                            const double* q, int ngq, const double* lambda, int nglambda,
                            const double* h, double* rhs, int ngrhs, void* stream);
/* ADD_QUAT_OP2D/3D (QuatFort.h:839, 983) */
int ampe_k_add_quat_op(int ndim, const int* lo, const int* hi, int depth, const double* mobility,
                       int ngmob, double* const* flux, int ngflux, const double* h, double* rhs,
                       int ngrhs, void* stream);
/* CORRECTRHSQUATFORSYMMETRY (QuatFort.h:562) */
int ampe_k_correctrhsquatforsymmetry(int ndim, const int* lo, const int* hi, int depth,
                                     const double* dx, double* const* nonsymm_diff,
                                     double* const* symm_diff, int ngdiff, double* rhs, int ngrhs,
                                     const double* quat, int ngq, double* const* facecoeff,
                                     int ngfacecoeff, const double* mobility, int ngmob,
                                     const int* const* iqrot, int ngiq, void* stream);
/* QUATMOBILITY (QuatFort.h:649) */
int ampe_k_quatmobility(int ndim, const int* ifirst, const int* ilast, const double* phase,
                        int ngphase, double* mobility, int ngmobility, double scale_mobility,
                        double min_mobility, char func_type, double alt_scale_factor, void* stream);

/* ---- concentrationrhs.m4 / flux.m4 / concentrationdiffusion.m4 ---------------------- */
/* CONCENTRATIONFLUX (ConcFort.h:16) */
int ampe_k_concentrationflux(int ndim, const int* ifirst, const int* ilast, const double* dx,
                             const double* conc, int ngconc, const double* phi, int ngphi,
                             double* const* diffconc, int ngdiffconc, double* const* dphicoupl,
                             int ngdphicoupl, double* const* flux, int ngflux, void* stream);
/* ADD_CAHNHILLIARDDOUBLEWELL_FLUX (ConcFort.h:40): flux must be zeroed by the caller
 * (CahnHilliardDoubleWell.cc:96) -- this entry computes the complete flux in one pass */
int ampe_k_add_cahnhilliarddoublewell_flux(int ndim, const int* ifirst, const int* ilast,
                                           const double* dx, const double* conc, int ngconc,
                                           double mobility, double ca, double cb,
                                           double well_scale, double kappa, double* const* flux,
                                           int ngflux, void* stream);
/* ADD_FLUX (ConcFort.h:139) */
int ampe_k_add_flux(int ndim, const int* ifirst, const int* ilast, const double* dx,
                    const double* conc, int ngconc, int ncomp, double* const* diffconc, int ngdiff,
                    double* const* flux, int ngflux, void* stream);
/* CONCENTRATION_PFMDIFFUSION (ConcFort.h:205) */
int ampe_k_concentration_pfmdiffusion(int ndim, const int* ifirst, const int* ilast,
                                      const double* phi, int ngphi, double* const* diff, int ngdiff,
                                      const double* temp, int ngtemp, double d_liquid,
                                      double q0_liquid, double d_solid_A, double q0_solid_A,
                                      double gas_constant_R, char interp_type, char avg_type,
                                      void* stream);
/* COMPUTERHSCONCENTRATION (ConcFort.h:419) */
int ampe_k_computerhsconcentration(int ndim, const int* ifirst, const int* ilast, const double* dx,
                                   double* const* flux, int ngflux, double mobility, double* rhs,
                                   int ngrhs, void* stream);

/* ---- per-patch loops of the C++ strategies (uniform temperature T) ------------------ */
/* CALPHADequilibriumPhaseConcentrationsStrategy::computePhaseConcentrationsOnPatch
 * (.cc:162-454) / QuadraticEquilibrium... (.cc:42-144): loops over the ghost box of c_l.
 * free_energy = AMPE_FE_CALPHAD | AMPE_FE_QUADRATIC; returns the number of failed cells */
int ampe_k_compute_phase_concentrations(const ampe_rhs_config* cfg, const int* ifirst,
                                        const int* ilast, const double* phi, int ngphi,
                                        const double* conc, int ngconc, const double* cl_ref,
                                        const double* ca_ref, double* cl, double* ca, int ngc,
                                        void* stream);
/* computeFreeEnergyLiquid / SolidA (CALPHADFreeEnergyStrategyBinary.cc:251-327,
 * QuadraticFreeEnergyStrategy.cc:171-247): f (ghost 0) = f(T, c_i) * 1e-6/V_m; phase 0|1 */
int ampe_k_compute_free_energy(const ampe_rhs_config* cfg, const int* ifirst, const int* ilast,
                               const double* c_i, int ngc, double* f, int phase, void* stream);
/* addDrivingForce (CALPHADFreeEnergyStrategyBinary.cc:521-667, Quadratic...:393-530) */
int ampe_k_add_driving_force(const ampe_rhs_config* cfg, const int* ifirst, const int* ilast,
                             const double* phi, int ngphi, const double* fl, const double* fa,
                             const double* cl, const double* ca, int ngc, double* rhs, int ngrhs,
                             void* stream);
/* MobilityCompositionDiffusionStrategy::setDiffusion (.cc:98-146, 197-611): side arrays
 * D_l*(1-h), D_a*h (ghost 0) from c_l, c_a, phi */
int ampe_k_set_ebs_diffusion(const ampe_rhs_config* cfg, const int* ifirst, const int* ilast,
                             const double* phi, int ngphi, const double* cl, const double* ca,
                             int ngc, double* const* diff_l, double* const* diff_a, void* stream);
/* KKSCompositionRHSStrategy::setDiffCoeffForPhaseOnPatch (.cc:210-373): D_phi = D0 h'(c_l-c_a) */
int ampe_k_set_kks_phase_diffusion(const ampe_rhs_config* cfg, const int* ifirst, const int* ilast,
                                   const double* phi, int ngphi, const double* cl, const double* ca,
                                   int ngc, double* const* d0, double* const* dphi, void* stream);
/* xfer::RefineSchedule::fillData on a fully periodic single patch: copy a ghost-0 array into the
 * interior of a ghosted one and fill the ghosts with the periodic images */
int ampe_k_fill_periodic(int ndim, const int* ifirst, const int* ilast, int depth,
                         const double* src, double* dst, int ng, void* stream);
/* math::PatchCellDataOpsReal::multiply(dst, a, b, box): dst = a * b on the interior box */
int ampe_k_cell_multiply(int ndim, const int* ifirst, const int* ilast, const double* a, int nga,
                         const double* b, int ngb, double* dst, int ngdst, void* stream);
/* ---- quatrotation.m4 / quatfacops.m4: symmetry pre-pass and projection (symmetry.cu) -------- */
/* QUAT_SYMM_ROTATION (QuatFort.h:319): rot[a] = SideData<int> of axis a, ghost ngrot, IN/OUT   */
int ampe_k_quat_symm_rotation(int ndim, const int* ifirst, const int* ilast, const double* q, int ngq,
                              int depth, int* const* rot, int ngrot, void* stream);
/* QUAT_FUNDAMENTAL (QuatFort.h:331): qlo/qhi = ghost box of the array (passed like the reference) */
int ampe_k_quat_fundamental(int ndim, const int* ifirst, const int* ilast, double* quat, const int* qlo,
                            const int* qhi, int depth, void* stream);
/* PROJECT2D / PROJECT3D (QuatFort.h:883, 1022): lo/hi boxes of q, corr, err as in the reference  */
int ampe_k_project(int ndim, const int* lo, const int* hi, int depth, const double* q, const int* qlo,
                   const int* qhi, double* corr, const int* clo, const int* chi, double* err, const int* elo,
                   const int* ehi, void* stream);
int ampe_k_fill_periodic_int(int ndim, const int* ifirst, const int* ilast, int axis,
                             const int* src, int* dst, int ng, void* stream);

/* ---- the dquat/dphi coupling block of the preconditioner (precond_coupling.cu) ----------------- */
/* QUATDIFFUSIONDERIV (QuatFort.h:546): diff[a] = SideData depth 2 (d/dphi of the lower, upper cell) */
int ampe_k_quatdiffusionderiv(int ndim, const int* ifirst, const int* ilast, double misorientation_factor,
                              const double* temperature, int tghosts, const double* var, int ngvar, int depth,
                              double* const* gradq, int nggradq, double* const* diff, int ngdiff,
                              double gradient_floor, char smooth_floor_type, char interp_type, char avg_type,
                              void* stream);
/* QUATMOBILITYDERIV (QuatFort.h:659) */
int ampe_k_quatmobilityderiv(int ndim, const int* ifirst, const int* ilast, const double* phase, int ngphase,
                             double* dmobility, int ngmobility, double scale_mobility, double min_mobility,
                             char func_type, double alt_scale_factor, void* stream);
/* COMPUTE_DQUATDPHI_FACE_COEF2D/3D (QuatFort.h:790) */
int ampe_k_compute_dquatdphi_face_coef(int ndim, const int* lo, const int* hi, int depth, double* const* dprime,
                                       int ngdprime, const double* phi, int ngphi, double* const* face_coef,
                                       int ngfc, void* stream);
/* MULTICOMPONENT_MULTIPLY2D/3D (QuatFort.h:892): var(:,n) *= factor for n < vnc */
int ampe_k_multicomponent_multiply(int ndim, const int* lo, const int* hi, const double* factor, int ngfactor,
                                   double* var, int ngvar, int vnc, void* stream);
/* TAKE_SQUARE_ROOT2D/3D (QuatFort.h:890): in place over the ghost box */
int ampe_k_take_square_root(int ndim, const int* lo, const int* hi, double* data, int ng, void* stream);
/* HierarchyCellDataOpsReal::axpy as used by QuatPrecondSolve (QuatIntegrator.cc:3612): dst = alpha x + y */
int ampe_k_cell_axpy(int ndim, const int* lo, const int* hi, int depth, double alpha, const double* x, int ngx,
                     const double* y, int ngy, double* dst, int ngdst, void* stream);

#ifdef __cplusplus
}
#endif
#endif
