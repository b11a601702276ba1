// Host side of the TMA staging path: CUtensorMap construction through the driver entry point
// (no link-time dependency on libcuda: the library still loads on a machine without a driver,
// which the CPU-side ABI tests rely on).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

namespace ampe {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
   static EncodeTiledFn fn = nullptr;
   static bool tried = false;
   if (!tried) {
      tried = true;
      void* p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
          q == cudaDriverEntryPointSuccess)
         fn = reinterpret_cast<EncodeTiledFn>(p);
      (void)cudaGetLastError();
   }
   return fn;
}

// The persistent TMA kernel is opt-in (AMPE_B200_TMA=1): measured on B200 it is 15 % slower than
// the cp.async tile kernel on Dendrite2D 2048^2 (0.155 vs 0.134 ms, profiles/README.md) -- the
// staging instructions it removes were hidden behind other resident blocks' FP64 phases anyway.
// Tests run both and compare bits.
bool tma_enabled()
{
   const char* e = getenv("AMPE_B200_TMA");
   return e != nullptr && e[0] == '1' && encode_fn() != nullptr;
}

// fp64 tensor (x fastest, rows, components) with a (box0 x box1 x 1) box; 0 on success
int tma_encode_3d(CUtensorMap* out, const double* base, unsigned long long n0, unsigned long long rows,
                  unsigned long long depth, unsigned long long comp_stride, unsigned box0, unsigned box1)
{
   EncodeTiledFn fn = encode_fn();
   if (!fn || !base) return 1;
   if ((reinterpret_cast<uintptr_t>(base) & 15u) || (n0 & 1ull) || (comp_stride & 1ull)) return 2;
   if (box0 > 256 || box1 > 256 || ((box0 * 8u) & 15u)) return 3;
   const cuuint64_t dims[3] = {n0, rows, depth ? depth : 1};
   const cuuint64_t strides[2] = {n0 * sizeof(double), (depth > 1 ? comp_stride : n0 * rows) * sizeof(double)};
   const cuuint32_t box[3] = {box0, box1, 1};
   const cuuint32_t estr[3] = {1, 1, 1};
   const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
   return r == CUDA_SUCCESS ? 0 : 4;
}

}  // namespace ampe
