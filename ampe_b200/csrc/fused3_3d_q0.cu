// explicit instantiation: NDIM=3, qlen=0, runtime model selectors (all composition forms)
#include "fused_launch3.cuh"
namespace ampe {
template int dispatch3_runtime<3, 0>(const FusedArgs&, cudaStream_t, const char**);
}
