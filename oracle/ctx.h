// TEST INFRASTRUCTURE ONLY (see oracle.h).  Private context of the oracle driver.
#pragma once
#include "oracle.h"
#include <vector>

#ifdef ORACLE_OMP
#define ORACLE_PAR _Pragma("omp parallel for collapse(2) schedule(static)")
#define ORACLE_PAR_RED(v) _Pragma("omp parallel for collapse(2) schedule(static) reduction(+:nfail)")
#else
#define ORACLE_PAR
#define ORACLE_PAR_RED(v)
#endif

namespace oracle {

struct Ctx {
   ampe_rhs_config cfg;
   double T0 = 0.0;  // T_uniform of the deck; cfg.T_uniform holds the temperature of the current evaluation
   Box box;
   int ng;  // nghosts_required(): 1, or 2 for Cahn-Hilliard (QuatModelParameters.h:305-311)
   // scratch state with ghosts (fillScratch targets)
   Field phase, quat, conc, temp;
   // intermediates (QuatModel.cc:1374-1745, QuatIntegrator.cc:898-977)
   SideField quat_diffs;                  // g=1, depth Q or 2Q
   SideField quat_grad_side, quat_grad_side_copy;  // g=0, depth D*Q
   Field quat_grad_cell[3];               // g=0, depth Q each
   Field quat_grad_modulus;               // g=0
   Field phase_mobility, quat_mobility;   // g=1
   SideField phase_flux;                  // g=0
   SideField face_coef, quat_flux;        // g=0
   Field lambda;                          // g=0
   SideField conc_flux;                   // g=0 (1 for CH)
   Field f_l, f_a;                        // g=0
   Field cl, ca, cl_ref, ca_ref;          // g=ng
   SideField diff_l, diff_a;              // EBS D_l, D_a   g=0
   SideField diff0, dphi;                 // KKS D0, D_phi  g=0
   Field cp, te;                          // heat capacity, melting T  g=0
   std::vector<int> iqrot_data[3];
   IView iqrot[3];
   // rhs work arrays g=0
   Field rhs_phase, rhs_quat, rhs_conc, rhs_temp;
   bool have_ref = false;
   void* precond = nullptr;     // block preconditioners (precond.cc)
   int precond_cycles = 0;      // > 0: the implicit integrator runs right-preconditioned (stepper.cc)
   bool precond_dquatdphi = false;  // with the dquat/dphi coupling block
   bool precond_left = false;       // PREC_LEFT like the reference (default here: right)
   double precond_stats[2] = {0, 0};  // set-ups, solves of the last implicit integration
};


// CALPHAD / quadratic passes (thermo_driver.cc)
int compute_phase_concentrations(Ctx* c);
void compute_free_energies(Ctx* c);
void add_driving_force(Ctx* c);
void set_diffusion_coeff_for_concentration(Ctx* c);
void compute_conc_flux_kks_ebs(Ctx* c);

// scalar energy diagnostics (energy.cc)
int energy(Ctx* c, const ampe_rhs_fields* y, double* out);
int scalar_diagnostics(Ctx* c, const ampe_rhs_fields* y, double* out);

}  // namespace oracle
