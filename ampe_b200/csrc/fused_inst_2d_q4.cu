// explicit instantiation: NDIM=2, qlen=4 (all composition forms / symmetry variants)
#include "fused_launch.cuh"
namespace ampe {
template int dispatch_conc<2, 4>(const FusedArgs&, cudaStream_t, const char**);
}
