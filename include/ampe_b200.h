/*
 * ampe_b200.h -- C ABI of libampe_b200.so: the B200-native phase-field
 * right-hand-side evaluation that AMPE drives from
 * QuatIntegrator::evaluateRHSFunction (reference: source/QuatIntegrator.cc:3134-3295).
 *
 * Two families of entry points ("lower boundary"):
 *
 *  (i)  ampe_rhs_*       the fused evaluateRHSFunction-shaped path: one context
 *                        per GPU, y / ydot are caller-owned DEVICE pointers laid
 *                        out like SAMRAI CellData with ghost width 0 (i fastest,
 *                        component slowest).
 *  (ii) ampe_k_*         one symbol per Fortran kernel AMPE binds through
 *                        source/fortran/{QuatFort,ConcFort}.h, same argument
 *                        order, scalars by value, arrays are DEVICE pointers in
 *                        SAMRAI CellData/SideData layout with ghost widths
 *                        (declared in ampe_b200_kernels.h).
 *
 * Every function returns 0 on success and a negative AMPE_E* code on error;
 * nothing here ever calls exit() (the Fortran kernels `stop`).
 * No torch types, no C++ types: plain pointers, ints and doubles only.
 */
#ifndef AMPE_B200_H
#define AMPE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define AMPE_OK 0
#define AMPE_EINVAL (-1)   /* bad argument / unsupported combination          */
#define AMPE_ECUDA (-2)    /* CUDA runtime error (see ampe_last_error)        */
#define AMPE_ENEWTON (-3)  /* per-cell KKS Newton did not converge            */
#define AMPE_ENOGPU (-4)   /* no CUDA device: there is no CPU fallback        */

/* phase flux stencils: PhaseFluxStrategyFactory.h:11-36 */
#define AMPE_FLUX_SIMPLE 0      /* gradient_flux            quatrhs.m4        */
#define AMPE_FLUX_ISOTROPIC 1   /* compute_flux_isotropic   2d/quatrhs.m4:106 */
#define AMPE_FLUX_ANISOTROPIC 2 /* anisotropic_gradient_flux                  */

/* concentration RHS forms: CompositionRHSStrategyFactory.h:27-102 */
#define AMPE_CONC_NONE 0
#define AMPE_CONC_CAHN_HILLIARD 1 /* CahnHilliardDoubleWell                   */
#define AMPE_CONC_KKS 2           /* KKSCompositionRHSStrategy (quadratic)    */
#define AMPE_CONC_EBS 3           /* EBSCompositionRHSStrategy (CALPHAD)      */

/* free energy / driving force: FreeEnergyStrategyFactory.h:37-263 */
#define AMPE_FE_NONE 0
#define AMPE_FE_BIASWELL 1  /* BiasDoubleWellUTRCFreeEnergyStrategy           */
#define AMPE_FE_CALPHAD 2   /* CALPHADFreeEnergyStrategyBinary                */
#define AMPE_FE_QUADRATIC 3 /* QuadraticFreeEnergyStrategy                    */
#define AMPE_FE_DELTAT 4    /* DeltaTemperatureFreeEnergyStrategy (FreeEnergyModel type "linear") */

#define AMPE_MAX_TC 6 /* max number of temperature intervals per species G(T) */

/* One species/phase Gibbs energy, thermodynamic_data/calphadAuNi.dat:1-46
 * G = a + b T + c T ln T + d2 T^2 + d3 T^3 + d4 T^4 + d7 T^7 + dm1/T + dm9/T^9
 * on interval [Tc[i], Tc[i+1]).                                              */
typedef struct ampe_calphad_species {
   int nintervals;
   double Tc[AMPE_MAX_TC + 1];
   double a[AMPE_MAX_TC], b[AMPE_MAX_TC], c[AMPE_MAX_TC];
   double d2[AMPE_MAX_TC], d3[AMPE_MAX_TC], d4[AMPE_MAX_TC], d7[AMPE_MAX_TC];
   double dm1[AMPE_MAX_TC], dm9[AMPE_MAX_TC];
} ampe_calphad_species;

/* CALPHAD binary two-phase data base + mobility block (calphadAuNi.dat).     */
typedef struct ampe_calphad_binary {
   ampe_calphad_species g[2][2]; /* [species A,B][phase L,A]                  */
   double L[2][4][2];            /* [phase][k][0:const,1:*T] Redlich-Kister   */
   /* MobilityParameters: [species][phase] qA,qB,q0AB..q3AB, each (a0,a1):
    * Q(T) = a0 + R T ln(a1)   (CALPHADMobility.h getQ)                       */
   double qA[2][2][2], qB[2][2][2], qAB[2][2][4][2];
   int nqAB[2][2];
} ampe_calphad_binary;

/* All parameters of one RHS evaluation (QuatModelParameters subset that the
 * hot path reads; names follow QuatModelParameters.h accessors).             */
typedef struct ampe_rhs_config {
   int ndim;      /* 2 or 3 (reference: compile-time -DNDIM)                  */
   int n[3];      /* LOCAL interior cells per direction (n[2]=1 in 2D)        */
   double dx[3];  /* mesh spacing                                             */
   int qlen;      /* 0: no orientation field; 1, 2 (KWCcomplex) or 4 (Quat)   */

   int with_phase;
   int with_concentration;
   int with_unsteady_temperature; /* Temperature{equation_type="unsteady"}    */
   int evolve_quat;               /* H_parameter > 0                          */

   int phase_flux_type; /* AMPE_FLUX_*                                        */
   int conc_rhs_form;   /* AMPE_CONC_*                                        */
   int free_energy;     /* AMPE_FE_*                                          */
   int symmetry_aware;  /* Symmetry{enabled=TRUE}                             */
   int lag_quat_sidegrad;            /* Integrator default TRUE               */
   int quat_grad_modulus_from_cells; /* quat_grad_modulus_type=="cells"       */

   /* selector characters: only the first character of the reference's
    * strings is significant (functions.f:28-83)                              */
   char energy_interp;    /* phi_interp_func_type  'p','h','l'                */
   char conc_interp;      /* conc_interp_func_type                            */
   char diffusion_interp; /* diffusion_interp_func_type (default 'l')         */
   char orient_interp1;   /* default 'q'                                      */
   char orient_interp2;   /* default 'c'                                      */
   char avg_func;         /* 'a' or 'h'                                       */
   char grad_floor_type;  /* 'm','t','s'                                      */
   char quat_mobility_func; /* 'p','e','i'                                    */
   char conc_avg_func;    /* ConcentrationModel.avg_func_type 'a' or 'h'      */

   /* phase */
   double epsilon_phase, epsilon_anisotropy;
   int knumber;
   double phi_well_scale, phi_mobility;
   /* orientation */
   double H_parameter, epsilon_q, quat_mobility, min_quat_mobility;
   double quat_grad_floor, quat_mobility_alt_scale;
   /* temperature */
   double T_uniform; /* value of the T field when it is not evolved           */
   double thermal_diffusivity, latent_heat, cp, meltingT;
   double bias_well_alpha, bias_well_gamma;
   /* composition */
   double conc_mobility;                          /* ConcentrationModel.mobility */
   double ch_ca, ch_cb, ch_well_scale, ch_kappa, ch_mobility;
   /* quadratic KKS (QuadraticFreeEnergyStrategy.cc:56-68)                    */
   double quad_Tref, quad_A_l, quad_Ceq_l, quad_m_l, quad_A_s, quad_Ceq_s, quad_m_s;
   double D_liquid, D_solid, Q0_liquid, Q0_solid; /* concentration_pfmdiffusion */
   double vm_liquid, vm_solid;                    /* molar volumes [m^3/mol]  */
   /* per-cell Newton (Thermo4PFM NewtonSolver block)                         */
   int newton_max_its;
   double newton_tol, newton_alpha;

   ampe_calphad_binary calphad;

   /* slab decomposition along the slowest axis (z in 3D, y in 2D): this
    * rank owns n[ndim-1] planes; ghost planes come from the neighbours.      */
   int nranks, rank;

   /* physical boundaries (Geometry.periodic_dimension, ModelParameters.BoundaryConditions): 0 = periodic in
    * this direction, 1 = every field has boundary = "slope", "0" on both faces of the direction
    * (QuatRefinePatchStrategy -> CartesianRobinBcHelper with a = 0, b = 1, g = 0: the ghost cells take the value of
    * the adjacent interior cell, edge / corner ghosts included).  Ghost width 1 models only.              */
   int zero_slope[3];
   /* ScalarTemperatureStrategy (ScalarTemperatureStrategy.cc:57-75): the uniform temperature of an evaluation
    * at `time` is T_uniform + dtemperaturedt * time, limited by target_temperature                          */
   double dtemperaturedt, target_temperature;
} ampe_rhs_config;

/* Device pointers of one state / RHS vector, SAMRAI CellData ghost 0:
 * index (i,j,k,m) -> i + n0*(j + n1*(k + n2*m)).  Unused components NULL.
 * Order mirrors createSolutionvector (QuatIntegrator.cc:1623-1674).          */
typedef struct ampe_rhs_fields {
   double* phase;
   double* quat;        /* depth qlen */
   double* conc;
   double* temperature;
} ampe_rhs_fields;

typedef struct ampe_rhs_ctx ampe_rhs_ctx;

/* Replaces the constructor-time wiring of QuatIntegrator (RegisterVariables,
 * QuatIntegrator.cc:787-985): allocates every intermediate on the device.    */
int ampe_rhs_create(const ampe_rhs_config* cfg, ampe_rhs_ctx** out);
int ampe_rhs_destroy(ampe_rhs_ctx* ctx);

/* QuatModel::resetRefPhaseConcentrations (QuatModel.cc:5218-5231): set the
 * Newton initial guess (c_l_ref, c_a_ref), ghost-0 device arrays; NULL means
 * "copy the last computed c_l, c_a".                                         */
int ampe_rhs_set_ref_concentrations(ampe_rhs_ctx* ctx, const double* cl_ref,
                                    const double* ca_ref, void* stream);
/* Multi-rank variant: arrays that already carry the ghost planes along the
 * slab axis (n_slab + 2*nghosts planes, first plane = lower ghost).          */
int ampe_rhs_set_ref_concentrations_ghosted(ampe_rhs_ctx* ctx, const double* cl_ref_g,
                                            const double* ca_ref_g, void* stream);
/* quat_symm_rotation SideData<int> (QuatModel.cc:1714-1722): one int per
 * LOWER face per direction, ghost 0 (device).                                */
int ampe_rhs_set_symmetry_rotations(ampe_rhs_ctx* ctx, const int* const* iqrot,
                                    void* stream);
/* Ghost planes from the slab neighbours (fillScratch, QuatIntegrator.cc:2873-2955).
 * lo/hi: fields with `nghosts` planes each, for the lower / upper neighbour.
 * NULL = periodic wrap inside this rank (single GPU).                        */
int ampe_rhs_set_halo(ampe_rhs_ctx* ctx, const ampe_rhs_fields* lo,
                      const ampe_rhs_fields* hi);
int ampe_rhs_nghosts(const ampe_rhs_ctx* ctx);

/* QuatIntegrator::evaluateRHSFunction(time, y, y_dot, fd_flag)
 * (QuatIntegrator.h:204-222).  y is not modified.  stream: cudaStream_t.     */
int ampe_rhs_eval(ampe_rhs_ctx* ctx, double time, const ampe_rhs_fields* y,
                  const ampe_rhs_fields* ydot, int fd_flag, void* stream);
/* Split evaluation for halo/compute overlap: interior planes first (no ghost
 * needed), boundary planes after the halo arrived.                           */
int ampe_rhs_eval_interior(ampe_rhs_ctx* ctx, double time,
                           const ampe_rhs_fields* y,
                           const ampe_rhs_fields* ydot, int fd_flag,
                           void* stream);
int ampe_rhs_eval_boundary(ampe_rhs_ctx* ctx, double time,
                           const ampe_rhs_fields* y,
                           const ampe_rhs_fields* ydot, int fd_flag,
                           void* stream);

/* ---- slab ghost-plane exchange between GPUs (fillScratch's RefineSchedule::fillData over MPI,
 * QuatIntegrator.cc:2873-2955, 2948-2954) -------------------------------------------------------
 * One rank per GPU, 1-D slabs along the slowest axis, periodic in the rank index.  Every rank owns receive
 * buffers in its own HBM (two parities x {lower, upper} x every state component x nghosts planes) and arrival
 * flags, in ONE allocation that the two neighbours map: CUDA IPC between processes, peer access inside one
 * process.  Per evaluation a rank PUSHES its boundary planes of y straight into the neighbours' buffers with one
 * kernel (NVLink stores, then a system-scope fence, then the epoch flag) and a one-thread kernel waits for the
 * neighbours' flags: no packing, no library collective, no host synchronisation; the interior planes are
 * evaluated while the planes travel.  Double buffering by epoch parity orders the reuse of a buffer behind the
 * neighbour's previous-but-one evaluation without a second handshake.  nranks == 1 needs none of this (the
 * kernels wrap periodically inside the rank).  With zero_slope set along the slab axis the ring is cut: the first
 * rank's lower and the last rank's upper ghost planes are their own adjacent planes (same calls, same handles).
 *
 * Set-up, once: create on every rank, export the opaque handle, ship it to both neighbours with the host
 * application's own transport (MPI_Sendrecv in AMPE; torch.distributed in the tests; a socket in tests/cpp),
 * connect with the neighbours' handles. */
typedef struct ampe_halo ampe_halo;
#define AMPE_HALO_HANDLE_BYTES 128
int ampe_halo_create(ampe_rhs_ctx* ctx, int rank, int nranks, ampe_halo** out);
int ampe_halo_export(ampe_halo* h, void* handle /* AMPE_HALO_HANDLE_BYTES */);
int ampe_halo_connect(ampe_halo* h, const void* handle_prev, const void* handle_next);
/* destroy BEFORE the context, and only after every rank has finished its last evaluation (the neighbours push into
 * this rank's buffers): the caller synchronises the ranks first (MPI_Barrier) */
int ampe_halo_destroy(ampe_halo* h);
/* evaluateRHSFunction on a slab with neighbours: push, interior planes, wait, boundary planes.  Collective over
 * the ranks in the sense that every rank must call it once per evaluation, in the same order. */
int ampe_rhs_eval_slab(ampe_rhs_ctx* ctx, ampe_halo* h, double time, const ampe_rhs_fields* y,
                       const ampe_rhs_fields* ydot, int fd_flag, void* stream);
/* the same through HOST buffers (the slab moves through the chunk pipeline of ampe_rhs_eval_host, the ghost
 * planes device to device) */
int ampe_rhs_eval_slab_host(ampe_rhs_ctx* ctx, ampe_halo* h, double time, const ampe_rhs_fields* y_host,
                            const ampe_rhs_fields* ydot_host, int fd_flag);
/* the two halves, for callers that sequence the overlap themselves: sides bit 0 = my lowest planes to the lower
 * neighbour, bit 1 = my highest planes to the upper neighbour; an epoch is complete when both were pushed */
int ampe_halo_push(ampe_halo* h, const ampe_rhs_fields* y, int sides, void* stream);
int ampe_halo_wait(ampe_halo* h, void* stream);
/* Newton reference concentrations / symmetry rotations of a slab rank: interior arrays (ghost 0) in, the ghost
 * planes along the slab axis come from the neighbours (collective) */
int ampe_rhs_set_ref_concentrations_slab(ampe_rhs_ctx* ctx, ampe_halo* h, const double* cl_ref,
                                         const double* ca_ref, void* stream);
int ampe_rhs_set_symmetry_rotations_slab(ampe_rhs_ctx* ctx, ampe_halo* h, const int* const* iqrot, void* stream);
/* QuatModel::computeSymmetryRotations (QuatModel.cc:4978-5055) on a slab rank: ghost planes of y exchanged, faces
 * searched, ghost planes of the indices fetched from the neighbours (collective) */
int ampe_rhs_compute_symmetry_rotations_slab(ampe_rhs_ctx* ctx, ampe_halo* h, const ampe_rhs_fields* y, void* stream);
/* launches of the last ampe_rhs_eval_slab (exchange kernels included) */
int ampe_halo_last_launch_count(const ampe_halo* h);

/* device pointers to ctx-owned c_l, c_a (ghost 0) after an evaluation        */
int ampe_rhs_get_phase_concentrations(ampe_rhs_ctx* ctx, double** cl,
                                      double** ca);
/* copy c_l, c_a (ghost 0) into caller-owned device arrays                   */
int ampe_rhs_copy_phase_concentrations(ampe_rhs_ctx* ctx, double* cl_out, double* ca_out,
                                       void* stream);
/* number of cells whose Newton failed in the last evaluation (synchronises)  */
int ampe_rhs_newton_failures(ampe_rhs_ctx* ctx, void* stream);
/* kernels launched by the last ampe_rhs_eval                                 */
int ampe_rhs_last_launch_count(const ampe_rhs_ctx* ctx);
/* Measurement aid (no reference counterpart; AMPE times these phases with its own tbox::Timer objects,
 * QuatIntegrator.cc:3140-3160 t_rhs_timer / t_phase_conc_timer): when switched on, every whole-slab evaluation
 * records CUDA events on its stream around the KKS pre-pass and around the fused kernel;
 * ampe_rhs_last_kernel_ms waits for the last evaluation and returns both durations in milliseconds. */
int ampe_rhs_set_kernel_timing(ampe_rhs_ctx* ctx, int on);
int ampe_rhs_last_kernel_ms(ampe_rhs_ctx* ctx, double* kks_ms, double* fused_ms);

/* host-buffer convenience used by the reference-facing plugin path: copies
 * y host->device, evaluates, copies ydot device->host (pinned or pageable).  */
int ampe_rhs_eval_host(ampe_rhs_ctx* ctx, double time, const ampe_rhs_fields* y_host,
                       const ampe_rhs_fields* ydot_host, int fd_flag);

/* ---- device-resident vector operations on the evolved components of the
 * solution vector -- the N_Vector operations CVODE calls (samrai/Sundials_SAMRAIVector.cc:
 * linearSum, scale, dotWith, weightedRMSNorm, maxNorm).  Reductions return to the host and
 * synchronise the stream; this rank's cells only.                                          */
int ampe_vec_linear_sum(ampe_rhs_ctx* ctx, double a, const ampe_rhs_fields* x, double b,
                        const ampe_rhs_fields* y, const ampe_rhs_fields* z, void* stream);
int ampe_vec_scale(ampe_rhs_ctx* ctx, double a, const ampe_rhs_fields* x,
                   const ampe_rhs_fields* z, void* stream);
int ampe_vec_dot(ampe_rhs_ctx* ctx, const ampe_rhs_fields* x, const ampe_rhs_fields* y,
                 double* result, void* stream);
int ampe_vec_wrms_norm(ampe_rhs_ctx* ctx, const ampe_rhs_fields* x, const ampe_rhs_fields* w,
                       double* result, void* stream);
int ampe_vec_max_norm(ampe_rhs_ctx* ctx, const ampe_rhs_fields* x, double* result, void* stream);
/* sum (w x)(w y): inner product of CVODE's scaled SPGMR (CVODESolver.cc:183-187, s1 = s2 = ewt)  */
int ampe_vec_wdot(ampe_rhs_ctx* ctx, const ampe_rhs_fields* x, const ampe_rhs_fields* y,
                  const ampe_rhs_fields* w, double* result, void* stream);
/* CVODE error weights from CVodeSStolerances(rtol, atol) (CVODESolver.cc:179): w = 1/(rtol|y|+atol) */
int ampe_vec_error_weights(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, double rtol, double atol,
                           const ampe_rhs_fields* w, void* stream);
/* QuatModel::normalizeQuat (QuatModel.cc:4222-4262): q <- q/|q| per cell, in place.        */
int ampe_normalize_quat(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, void* stream);
/* Fixed-step explicit integrator keeping y on the device (scheme 0 forward Euler, 1 Heun):
 * stand-in for the CVODE loop of QuatIntegrator::Advance in trajectory tests; after each step
 * normalizeQuat and, for CALPHAD, resetRefPhaseConcentrations.  work1 (and work2 for Heun) are
 * caller-owned vectors shaped like y.                                                      */
int ampe_integrate_fixed(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, const ampe_rhs_fields* work1,
                         const ampe_rhs_fields* work2, double t0, double dt, int nsteps, int scheme,
                         void* stream);
/* the same loop on a slab rank: every evaluation exchanges the ghost planes through `h` (collective over the ranks) */
int ampe_integrate_fixed_slab(ampe_rhs_ctx* ctx, ampe_halo* h, const ampe_rhs_fields* y, const ampe_rhs_fields* work1,
                              const ampe_rhs_fields* work2, double t0, double dt, int nsteps, int scheme, void* stream);
/* ---- grain diagnostics: QuatModel::computeGrainDiagnostics (QuatModel.cc:2690-2705) ->
 * Grains::findAndNumberGrains (Grains.cc:263-520) + Grains::computeGrainVolumes (Grains.cc:647-697), the
 * "Volume of grain N = V" lines the regression decks check.  Cells with phase >= phase_threshold
 * (GrainDiagnostics{phase_threshold}, default 0.85) that touch through faces form a grain (periodic directions
 * wrap, zero-slope boundaries do not connect); a grain's number is the lowest global cell index it contains
 * (i + n0 (j + n1 k)), its volume the sum of its cells' control volumes.  Host outputs, ascending grain number
 * (the reference's std::map order); more than max_grains grains: AMPE_EINVAL with the count in *ngrains.
 * ampe_grain_numbers copies the per-cell grain numbers (-1 outside grains; d_grain_number_id) of the last call
 * into a device array of ncell ints.  Single rank.                                             */
int ampe_grain_volumes(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, double phase_threshold, int max_grains,
                       int* ngrains, int* grain_ids, double* volumes, void* stream);
int ampe_grain_numbers(ampe_rhs_ctx* ctx, int* grain_number);
/* ---- QuatModel::evaluateEnergy (QuatModel.cc:4888-4976) ->
 * quatenergy / bulkenergy ({2d,3d}/quatenergy.m4).  out[8] = total, phase interface,
 * orientational, q interface, double well, bulk free energy, 0, 0 (host array).  For the
 * Cahn-Hilliard model (no evaluator in the reference) the PFHub-1a functional
 * sum [w (c-ca)^2 (cb-c)^2 + kappa/2 |grad c|^2] dV: out[0] total, [1] gradient, [4] well.   */
int ampe_energy_eval(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, double* out, void* stream);

/* QuatModel::printScalarDiagnostics (QuatModel.cc:2543-2690), this rank's cells.  out[12] (host): domain
 * volume, volume of solid (evaluateVolumeSolid, QuatModel.cc:5170: L1 norm of phi), its fraction, integral
 * concentration (:5106), max concentration, integral phase concentration (:5130: L1 norm of c phi),
 * Cex = (cphi - c0 vphi) / c0V0 (:2655), min / max / average temperature, thermal energy
 * (computeThermalEnergy, :5373: -L int phi + int cp T), 0.  A multi-rank caller combines the ranks.   */
int ampe_scalar_diagnostics(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, double* out, void* stream);

/* ---- the CVODE projection hook.
 * QuatIntegrator::applyProjection(time, y, corr, epsProj, err) (QuatIntegrator.cc:3911-3962,
 * CVODEAbstractFunctions.h applyProjection): every evolved component of corr is zeroed; when the
 * orientation is evolved with qlen > 1, QuatSysSolver::applyProjection -> PROJECT{2,3}D
 * (QuatFACOps.cc:2395-2434, 3d/quatfacops.m4:1022-1081) makes y + corr a unit quaternion per cell
 * and removes from err its component along q.  y is not modified; err is updated in place.  */
int ampe_apply_projection(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, const ampe_rhs_fields* corr,
                          const ampe_rhs_fields* err, void* stream);
/* ---- the symmetry pre-pass that produces the inputs of rows a7 / a10.
 * QuatModel::computeSymmetryRotations (QuatModel.cc:4978-5055 -> QUAT_SYMM_ROTATION,
 * {2d,3d}/quatrotation.m4:12-89): rotation index of every lower face from y->quat on the periodic
 * level; kept in the context (same storage as ampe_rhs_set_symmetry_rotations).  The previous
 * indices seed the search like the reference's in/out SideData<int>; 0 after creation.        */
int ampe_rhs_compute_symmetry_rotations(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, void* stream);
/* the context's rotation indices, ghost 0, one caller-owned device array per direction       */
int ampe_rhs_get_symmetry_rotations(ampe_rhs_ctx* ctx, int* const* iqrot_out, void* stream);
/* QuatModel::makeQuatFundamental (QuatModel.cc:5059-5104 -> QUAT_FUNDAMENTAL,
 * quatrotation.m4:93-147): y->quat <- symmetric equivalent closest to the identity, in place.  */
int ampe_quat_fundamental(ampe_rhs_ctx* ctx, const ampe_rhs_fields* y, void* stream);

const char* ampe_last_error(void);
const char* ampe_version(void);
/* sizeof(ampe_rhs_config) as compiled, for binding sanity checks */
int ampe_abi_sizeof_config(void);

#ifdef __cplusplus
}
#endif
#endif
