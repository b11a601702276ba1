// compile-time model selectors of examples/AuNi_2D (CALPHAD KKS, EBS, quaternion symmetry on/off)
#include "fused_launch3.cuh"
namespace ampe {
int dispatch3_fixed_auni2d(const FusedArgs& A, cudaStream_t st, const char** err, int* rc)
{
   const Params& p = A.p;
   if (p.ndim == 2 && p.qlen == 4 && p.conc_form == AMPE_CONC_EBS && !p.with_T && sel_matches<SelAuNi>(p)) {
      *rc = p.symm ? launch3<2, 4, AMPE_CONC_EBS, true, false, SelAuNi>(A, st, err)
                   : launch_any<2, 4, AMPE_CONC_EBS, false, false, SelAuNi>(A, st, err);
      return 1;
   }
   return 0;
}
}  // namespace ampe
