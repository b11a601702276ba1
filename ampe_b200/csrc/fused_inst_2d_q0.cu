// explicit instantiation: NDIM=2, qlen=0 (all composition forms / symmetry variants)
#include "fused_launch.cuh"
namespace ampe {
template int dispatch_conc<2, 0>(const FusedArgs&, cudaStream_t, const char**);
}
