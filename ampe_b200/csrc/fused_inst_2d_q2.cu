// explicit instantiation: NDIM=2, qlen=2 (all composition forms / symmetry variants)
#include "fused_launch.cuh"
namespace ampe {
template int dispatch_conc<2, 2>(const FusedArgs&, cudaStream_t, const char**);
}
