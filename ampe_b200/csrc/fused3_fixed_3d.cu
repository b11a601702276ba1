// compile-time model selectors of examples/AuNi_3D (CALPHAD KKS, EBS) and of
// tests/TwoGrainsQuadratic/3d.input = GG3D_HBSM with orientation on (quadratic KKS)
#include "fused_launch3.cuh"
namespace ampe {
int dispatch3_fixed_3d(const FusedArgs& A, cudaStream_t st, const char** err, int* rc)
{
   const Params& p = A.p;
   if (p.ndim != 3 || p.qlen != 4 || p.with_T || p.symm) return 0;
   if (p.conc_form == AMPE_CONC_EBS && sel_matches<SelAuNi>(p)) {
      if (A.split3d && !A.energy_partials) {
         // AMPE_B200_SPLIT3D=1 (experiment, rhs_march.cuh): phase + quaternion RHS and composition RHS
         // as two lighter launches; same expressions, bit-identical outputs.  Returns 2 = two launches.
         *rc = launch_march<4, AMPE_CONC_EBS, false, SelAuNi, 1>(A, st, err);
         if (*rc == AMPE_OK) *rc = launch_march<0, AMPE_CONC_EBS, false, SelAuNi, 2>(A, st, err);
         return 2;
      }
      *rc = launch_any<3, 4, AMPE_CONC_EBS, false, false, SelAuNi>(A, st, err);
      return 1;
   }
   if (p.conc_form == AMPE_CONC_KKS && sel_matches<SelHBSM>(p)) {
      *rc = launch_any<3, 4, AMPE_CONC_KKS, false, false, SelHBSM>(A, st, err);
      return 1;
   }
   return 0;
}
}  // namespace ampe
