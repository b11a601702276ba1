// explicit instantiation: NDIM=3, qlen=0 (all composition forms / symmetry variants)
#include "fused_launch.cuh"
namespace ampe {
template int dispatch_conc<3, 0>(const FusedArgs&, cudaStream_t, const char**);
}
