// compile-time model selectors of examples/Dendrite2D (KWCcomplex, anisotropic flux, bias well,
// unsteady temperature); returns 1 if it handled the launch, 0 if the parameters do not match
#include "fused_launch3.cuh"
namespace ampe {
int dispatch3_fixed_dendrite(const FusedArgs& A, cudaStream_t st, const char** err, int* rc)
{
   const Params& p = A.p;
   if (p.ndim == 2 && p.qlen == 2 && p.conc_form == 0 && p.with_T && !p.symm && sel_matches<SelDendrite>(p)) {
      *rc = launch_any<2, 2, 0, false, true, SelDendrite>(A, st, err);
      return 1;
   }
   return 0;
}
}  // namespace ampe
