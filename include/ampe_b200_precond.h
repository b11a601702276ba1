/*
 * ampe_b200_precond.h -- C ABI of the block preconditioners of AMPE's CVODE Newton-Krylov loop
 * (reference: QuatIntegrator::CVSpgmrPrecondSet / CVSpgmrPrecondSolve,
 * source/QuatIntegrator.cc:3300-3376 and 3666-3771), part of libampe_b200.so.
 *
 * One ampe_mg object replaces one of the reference's block solvers on a single periodic uniform
 * level: PhaseFACSolver / ConcFACSolver / TemperatureFACSolver (EllipticFACOps + hypre PFMG) or
 * QuatSysSolver (QuatFACOps + QuatLevelSolver).  The operator is
 *
 *      (A u)_i = c_i u_i + m_i sum_faces d_f (s_nb u_nb - s_i u_i),      d_f = D_f / h_f^2
 *
 *   scalar blocks  A u = M div(D grad u) + C u   (EllipticFACOps.h:35, 2d/ellipticfacops.m4:16-56,
 *                  346-393): c = C, m = M, s = 1
 *   quaternion     A w = w + gamma sqrt(mob) div(fc grad(sqrt(mob) w))  (2d/quatlevelsolver.m4:9-118):
 *                  c = 1, m = gamma sqrt(mob), s = sqrt(mob), the same matrix for every component
 *
 * and is inverted approximately by a FIXED number of geometric multigrid V-cycles from a zero
 * initial guess (a fixed linear operator, no host synchronisation) with the reference's red-black
 * Gauss-Seidel update as the smoother (efo_rbgswithfluxmax*, 2d/ellipticfacops.m4:60-130).
 *
 * Implementation notes (csrc/mg.cu): coefficients that are constants of a block are not stored;
 * above 4096 cells a sweep is one fused red-black pass over shared-memory tiles, residual and restriction are
 * one pass, and every level from 4096 cells down runs inside one block -- all bit-identical to the plain
 * colour half-sweeps.  Environment switches read by ampe_mg_create: AMPE_B200_MG_FUSED=0, AMPE_B200_MG_TAIL=0
 * (the plain variants, for A/B runs), AMPE_B200_MG_GRAPH=1 (a solve captured once and replayed; needs a
 * non-default stream).
 *
 * Arrays named "SAMRAI layout" are device pointers laid out like CellData / SideData over the box
 * [0, n-1] with the stated ghost width (i fastest); rhs / soln / u / out are ghost-0 cell arrays.
 * Array-of-pointer arguments are HOST arrays of device pointers.  Return codes: ampe_b200.h.
 */
#ifndef AMPE_B200_PRECOND_H
#define AMPE_B200_PRECOND_H

#include "ampe_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ampe_mg ampe_mg;

/* EllipticFACSolver::initializeSolverState / QuatSysSolver::initializeSolverState: allocates every
 * level on the device.  with_column_scale != 0: the quaternion block (ampe_mg_set_quat).          */
int ampe_mg_create(int ndim, const int* n, const double* dx, int with_column_scale, ampe_mg** out);
/* ncomp (<= 8) components solved together with the one matrix -- the qlen components of the quaternion block,
 * which QuatLevelSolver::solveSystem solves depth by depth with the same matrix (QuatLevelSolver.cc:1386-1635):
 * rhs / soln of ampe_mg_solve are then depth-ncomp arrays; every pass updates all components and reads the
 * coefficients once.                                                                                       */
int ampe_mg_create_multi(int ndim, const int* n, const double* dx, int with_column_scale, int ncomp, ampe_mg** out);
int ampe_mg_num_components(const ampe_mg* mg);
int ampe_mg_destroy(ampe_mg* mg);

/* EllipticFACOps::setM / setMConstant, setCPatchDataId / setCConstant, setDPatchDataId /
 * setDConstant (EllipticFACOps.h:145-244).  m, c: CellData (ghost ngm / ngc) or NULL = the constant.
 * d: SideData (ghost ngd), D = d_scale * (d [+ d2]), or NULL = d_const.  d2 is the second array of
 * EBSCompositionRHSStrategy::setDiffusionCoeffForPreconditioner (D_l + D_a,
 * EBSCompositionRHSStrategy.cc:332-430), NULL otherwise.  Rebuilds the coarse levels.            */
int ampe_mg_set_elliptic(ampe_mg* mg, const double* m, int ngm, double m_const, const double* c, int ngc,
                         double c_const, const double* const* d, const double* const* d2, int ngd,
                         double d_scale, double d_const, void* stream);
/* QuatFACOps::setOperatorCoefficients (QuatFACOps.cc:735-818) -> QuatLevelSolver::
 * setMatrixCoefficients (QuatLevelSolver.cc:753-1080): mobility CellData (ghost ngm; its square root
 * is taken here, takeSquareRootOnPatch), face_coef SideData depth 1 (ghost ngfc).                */
int ampe_mg_set_quat(ampe_mg* mg, double gamma, const double* mobility, int ngm,
                     const double* const* face_coef, int ngfc, void* stream);
/* solveSystem: soln ~ A^-1 rhs by ncycles V-cycles from zero.  symmetrized != 0 (quaternion block):
 * the right-hand side is first divided and the solution finally multiplied by sqrt(mob)
 * (QuatSysSolver::solveSystem, QuatSysSolver.cc:308 and 328).  rhs and soln may alias.           */
int ampe_mg_solve(ampe_mg* mg, const double* rhs, double* soln, int ncycles, int symmetrized, void* stream);
/* out = A u on the finest level (residual check; u != out)                                       */
int ampe_mg_apply(ampe_mg* mg, const double* u, double* out, void* stream);
/* pre / post smoothing sweeps per level (default 1, 1) and sweeps on the coarsest level (8)      */
int ampe_mg_set_sweeps(ampe_mg* mg, int pre, int post, int coarse);
/* Physical boundaries of the level: zero_slope[d] != 0 = homogeneous Neumann on both faces of direction d (the
 * blocks of a deck whose BoundaryConditions are "slope", "0": the FAC solvers get their Robin coefficients from the same
 * database, EllipticFACSolver::setBoundaries / QuatFACOps::setPhysicalBcCoefObject); 0 = periodic.  Call before the
 * set_* functions: the coefficient of every boundary face is zero on every level, the prolongation does not interpolate
 * across the boundary. */
int ampe_mg_set_zero_slope(ampe_mg* mg, const int* zero_slope);
int ampe_mg_num_levels(const ampe_mg* mg);
int ampe_mg_level_extents(const ampe_mg* mg, int level, int* n_out /* [3] */);
/* copy one coefficient array of a level into a caller-owned device array: which = 0 c, 1 m, 2 s,
 * 3..5 d of direction 0..2                                                                        */
int ampe_mg_copy_level(ampe_mg* mg, int level, int which, double* out, void* stream);
/* kernels launched by the last set_* / solve / apply call                                        */
int ampe_mg_last_launch_count(const ampe_mg* mg);

/* PhaseFACOps::setCOnPatchPrivate (PhaseFACOps.cc:100-186): C = 1 + gamma m w g''(phi) with
 * g'' = second_deriv_well_func (functions.f); SAMRAI layout, the box starts at 0.                 */
int ampe_k_phasefacops_setc(int ndim, const int* ifirst, const int* ilast, const double* phi, int ngphi,
                            const double* m, int ngm, double gamma, double phi_well_scale,
                            const char* phi_well_func_type, double* c, int ngc, void* stream);

#ifdef __cplusplus
}
#endif
#endif
