// explicit instantiation: NDIM=3, qlen=2 (all composition forms / symmetry variants)
#include "fused_launch.cuh"
namespace ampe {
template int dispatch_conc<3, 2>(const FusedArgs&, cudaStream_t, const char**);
}
