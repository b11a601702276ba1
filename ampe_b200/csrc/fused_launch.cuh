// Launch + template dispatch of the fused kernel.  Included only by the
// fused_inst_*.cu translation units (one per (NDIM, qlen) pair so that the
// instantiations compile in parallel); ctx.cu sees extern template declarations.
#pragma once
#include <cuda_runtime.h>

#include "rhs_fused.cuh"

namespace ampe {

template <int ND, int Q, int CONC, bool SYMM>
static int launch_fused(const FusedArgs& A, cudaStream_t st, const char** err)
{
   // tile shape and block size (tuned on B200, see profiles/): 2D 32x16 cells, 256 threads;
   // 3D 32x4x4 cells, 512 threads (shared memory, not registers, bounds the 3D occupancy)
#ifndef AMPE_T2Y
#define AMPE_T2Y 16
#define AMPE_NT2 256
#endif
#ifndef AMPE_T3Y
#define AMPE_T3Y 4
#define AMPE_T3Z 4
#define AMPE_NT3 512
#endif
   constexpr int TX = 32;
   constexpr int TY = (ND == 2) ? AMPE_T2Y : AMPE_T3Y;
   constexpr int TZ = (ND == 2) ? 1 : AMPE_T3Z;
   constexpr int NT = (ND == 2) ? AMPE_NT2 : AMPE_NT3;
   using G = TileGeom<ND, TX, TY, TZ>;
   const Params& p = A.p;
   size_t doubles = (size_t)G::S * (1 + (p.with_T ? 1 : 0) + Q + (CONC == AMPE_CONC_KKS ? 1 : 0) +
                                    (CONC != 0 ? 2 : 0));
   doubles += (size_t)ND * G::NFB * (((Q > 0 && p.evolve_quat) ? 1 : 0) + (p.flux_type != AMPE_FLUX_SIMPLE ? 1 : 0) +
                               (CONC != 0 ? 1 : 0));
   size_t bytes = doubles * sizeof(double) + (SYMM ? (size_t)ND * G::S * sizeof(int) : 0);
   auto kern = rhs_fused_kernel<ND, Q, CONC, SYMM, TX, TY, TZ, NT>;
   static size_t configured = 0;
   if (bytes > configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e != cudaSuccess) {
         *err = cudaGetErrorString(e);
         return AMPE_ECUDA;
      }
      configured = bytes;
   }
   const int nslab = A.s_end - A.s_begin;
   if (nslab <= 0) return AMPE_OK;
   dim3 grid;
   grid.x = (p.n[0] + TX - 1) / TX;
   if (ND == 2) {
      grid.y = (nslab + TY - 1) / TY;
      grid.z = 1;
   } else {
      grid.y = (p.n[1] + TY - 1) / TY;
      grid.z = (nslab + TZ - 1) / TZ;
   }
   kern<<<grid, NT, bytes, st>>>(A);
   cudaError_t e2 = cudaGetLastError();
   if (e2 != cudaSuccess) {
      *err = cudaGetErrorString(e2);
      return AMPE_ECUDA;
   }
   return AMPE_OK;
}

template <int ND, int Q>
int dispatch_conc(const FusedArgs& A, cudaStream_t st, const char** err)
{
   const bool symm = A.p.symm;
   switch (A.p.conc_form) {
      case 0:
      case AMPE_CONC_CAHN_HILLIARD:
         if (symm) {
            if constexpr (Q == 4) return launch_fused<ND, Q, 0, true>(A, st, err);
            { *err = "symmetry needs qlen=4 in this build"; return AMPE_EINVAL; }
         }
         return launch_fused<ND, Q, 0, false>(A, st, err);
      case AMPE_CONC_KKS:
         if (symm) {
            if constexpr (Q == 4) return launch_fused<ND, Q, AMPE_CONC_KKS, true>(A, st, err);
            { *err = "symmetry needs qlen=4 in this build"; return AMPE_EINVAL; }
         }
         return launch_fused<ND, Q, AMPE_CONC_KKS, false>(A, st, err);
      case AMPE_CONC_EBS:
         if (symm) {
            if constexpr (Q == 4) return launch_fused<ND, Q, AMPE_CONC_EBS, true>(A, st, err);
            { *err = "symmetry needs qlen=4 in this build"; return AMPE_EINVAL; }
         }
         return launch_fused<ND, Q, AMPE_CONC_EBS, false>(A, st, err);
   }
   *err = "unknown conc_rhs_form";
   return AMPE_EINVAL;
}


}  // namespace ampe
