// Launch + template dispatch of the generation-3 fused kernel.  Included only by the
// fused3_*.cu translation units (one per (NDIM, qlen, selector policy) so that the
// instantiations compile in parallel); ctx.cu sees the dispatch3_* prototypes.
#pragma once
#include <cuda_runtime.h>

#include "energy_tile.cuh"
#include "rhs_march.cuh"
#include "rhs_tile.cuh"
#include "rhs_tile_tma.cuh"
#include "tile_shape.h"

namespace ampe {

#define AMPE_MAX_DEVICES 64
static inline int current_device()
{
   int dev = 0;
   if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= AMPE_MAX_DEVICES) dev = 0;
   return dev;
}

// energy diagnostics on the same tile geometry (energy_tile.cuh)
template <class TT>
static int launch_energy(const FusedArgs& A, cudaStream_t st, const char** err)
{
   const Params& p = A.p;
   auto kern = energy_tile_kernel<TT>;
   // the dynamic shared-memory attribute belongs to the device's context: one flag per device
   static bool configured_dev[AMPE_MAX_DEVICES] = {};
   bool& configured = configured_dev[current_device()];
   if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TT::SMEM_BYTES);
      if (e != cudaSuccess) {
         *err = cudaGetErrorString(e);
         return AMPE_ECUDA;
      }
      configured = true;
   }
   const int nslab = A.s_end - A.s_begin;
   dim3 grid((p.n[0] + TT::TX - 1) / TT::TX, 1, 1);
   if (TT::ND == 2) {
      grid.y = (nslab + TT::TY - 1) / TT::TY;
   } else {
      grid.y = (p.n[1] + TT::TY - 1) / TT::TY;
      grid.z = (nslab + TT::TZ - 1) / TT::TZ;
   }
   kern<<<grid, TT::NT, TT::SMEM_BYTES, st>>>(A);
   cudaError_t e2 = cudaGetLastError();
   if (e2 != cudaSuccess) {
      *err = cudaGetErrorString(e2);
      return AMPE_ECUDA;
   }
   return AMPE_OK;
}

bool tma_enabled();
int tma_encode_3d(CUtensorMap* out, const double* base, unsigned long long n0, unsigned long long rows,
                  unsigned long long depth, unsigned long long comp_stride, unsigned box0, unsigned box1);

// 2D persistent kernel with TMA staging (rhs_tile_tma.cuh).  Returns -1 when this evaluation
// cannot use it (odd row length, misaligned arrays, no driver entry point): the caller then
// launches the cp.async tile kernel.
template <int Q, int CONC, bool WT, class SEL>
static int launch_tma(const FusedArgs& A, cudaStream_t st, const char** err)
{
   using TT = Tile3<2, Q, CONC, false, WT, SEL, 32, AMPE_TMA_TY, 1, AMPE_TMA_NT, 2>;
   using TM = TmaTile<TT>;
   const Params& p = A.p;
   const int nslab = A.s_end - A.s_begin;
   if (nslab <= 0) return AMPE_OK;
   if (!tma_enabled()) return -1;
   if (A.wait_epoch) return -1;              // slab ranks waiting inside the kernel: cp.async tile kernel
   if (p.clamp[0] || p.clamp[1]) return -1;  // physical boundaries: the cp.async tile kernel clamps while staging
   const unsigned long long n0 = p.n[0], ns = p.n[1], ncell = n0 * ns;
   TmaMaps M;
   int bad = tma_encode_3d(&M.phi, A.phi.base, n0, ns, 1, ncell, TT::SX, TT::SY);
   if (WT) bad |= tma_encode_3d(&M.T, A.T.base, n0, ns, 1, ncell, TT::SX, TT::SY);
   if (Q > 0) bad |= tma_encode_3d(&M.q, A.q.base, n0, ns, Q, (unsigned long long)A.q.comp, TT::SX, TT::SY);
   if (CONC == AMPE_CONC_KKS) bad |= tma_encode_3d(&M.conc, A.conc.base, n0, ns, 1, ncell, TT::SX, TT::SY);
   if (CONC != 0) {
      bad |= tma_encode_3d(&M.cl, A.cl, n0, ns + 2, 1, ncell, TT::SX, TT::SY);
      bad |= tma_encode_3d(&M.ca, A.ca, n0, ns + 2, 1, ncell, TT::SX, TT::SY);
   }
   if (bad) return -1;
   auto kern = rhs_tile_tma_kernel<TT>;
   static int resident_dev[AMPE_MAX_DEVICES] = {};  // persistent grid: resident blocks per SM x SMs
   int& resident = resident_dev[current_device()];
   if (!resident) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM::SMEM_BYTES);
      int per_sm = 0, dev = 0, sms = 0;
      if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TT::NT, TM::SMEM_BYTES);
      if (e == cudaSuccess) e = cudaGetDevice(&dev);
      if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (e != cudaSuccess || per_sm < 1) {
         (void)cudaGetLastError();
         return -1;
      }
      resident = per_sm * sms;
   }
   const int tiles_x = (p.n[0] + TT::TX - 1) / TT::TX;
   const int ntiles = tiles_x * ((nslab + TT::TY - 1) / TT::TY);
   const int grid = ntiles < resident ? ntiles : resident;
   kern<<<grid, TT::NT, TM::SMEM_BYTES, st>>>(A, M, tiles_x, ntiles);
   cudaError_t e2 = cudaGetLastError();
   if (e2 != cudaSuccess) {
      *err = cudaGetErrorString(e2);
      return AMPE_ECUDA;
   }
   return AMPE_OK;
}

template <int ND, int Q, int CONC, bool SYMM, bool WT, class SEL>
static int launch3(const FusedArgs& A, cudaStream_t st, const char** err)
{
   constexpr int TX = 32;
   constexpr int TY = (ND == 2) ? AMPE_T2Y : AMPE_T3Y;
   constexpr int TZ = (ND == 2) ? 1 : AMPE_T3Z;
   constexpr int NT = (ND == 2) ? AMPE_NT2 : AMPE_NT3;
   using TT = Tile3<ND, Q, CONC, SYMM, WT, SEL, TX, TY, TZ, NT>;
   const Params& p = A.p;
   auto kern = rhs_tile_kernel<TT>;
   // the dynamic shared-memory attribute belongs to the device's context: one flag per device
   static bool configured_dev[AMPE_MAX_DEVICES] = {};
   bool& configured = configured_dev[current_device()];
   if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TT::SMEM_BYTES);
      if (e != cudaSuccess) {
         *err = cudaGetErrorString(e);
         return AMPE_ECUDA;
      }
      configured = true;
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
   kern<<<grid, NT, TT::SMEM_BYTES, st>>>(A);
   cudaError_t e2 = cudaGetLastError();
   if (e2 != cudaSuccess) {
      *err = cudaGetErrorString(e2);
      return AMPE_ECUDA;
   }
   return AMPE_OK;
}

// 3D plane-marching kernel (no quaternion symmetry): column 32 x TY, NZ planes per block
template <int Q, int CONC, bool WT, class SEL, int PART = 0>
static int launch_march(const FusedArgs& A, cudaStream_t st, const char** err)
{
   using TT = March3<Q, CONC, WT, SEL, (PART == 0) ? AMPE_MY : ((PART == 1) ? AMPE_SPLIT_MY1 : AMPE_SPLIT_MY2), AMPE_MZ, PART>;
   const Params& p = A.p;
   auto kern = rhs_march_kernel<TT>;
   // the dynamic shared-memory attribute belongs to the device's context: one flag per device
   static bool configured_dev[AMPE_MAX_DEVICES] = {};
   bool& configured = configured_dev[current_device()];
   if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TT::SMEM_BYTES);
      if (e != cudaSuccess) {
         *err = cudaGetErrorString(e);
         return AMPE_ECUDA;
      }
      configured = true;
   }
   const int nslab = A.s_end - A.s_begin;
   if (nslab <= 0) return AMPE_OK;
   dim3 grid((p.n[0] + TT::TX - 1) / TT::TX, (p.n[1] + TT::TY - 1) / TT::TY, (nslab + TT::NZ - 1) / TT::NZ);
   kern<<<grid, TT::NT, TT::SMEM_BYTES, st>>>(A);
   cudaError_t e2 = cudaGetLastError();
   if (e2 != cudaSuccess) {
      *err = cudaGetErrorString(e2);
      return AMPE_ECUDA;
   }
   return AMPE_OK;
}

// launch of one (NDIM, qlen, CONC, SYMM, WT, SEL): 3D without symmetry marches, the rest tiles
template <int ND, int Q, int CONC, bool SYMM, bool WT, class SEL>
static int launch_any(const FusedArgs& A, cudaStream_t st, const char** err)
{
   if constexpr (!SEL::fixed) {
      // energy diagnostics: tile geometry for every model, runtime selectors only
      if (A.energy_partials) {
         using TT = Tile3<ND, Q, CONC, SYMM, WT, SEL, 32, (ND == 2) ? AMPE_T2Y : AMPE_T3Y, (ND == 2) ? 1 : AMPE_T3Z,
                          (ND == 2) ? AMPE_NT2 : AMPE_NT3>;
         return launch_energy<TT>(A, st, err);
      }
   }
   if constexpr (ND == 3 && !SYMM) {
      return launch_march<Q, CONC, WT, SEL>(A, st, err);
   }
   else {
      if constexpr (ND == 2 && !SYMM) {
         const int rc = launch_tma<Q, CONC, WT, SEL>(A, st, err);
         if (rc >= 0) return rc;
      }
      return launch3<ND, Q, CONC, SYMM, WT, SEL>(A, st, err);
   }
}

// runtime-selector instantiations of one (NDIM, qlen): every composition form, symmetry for qlen 4
template <int ND, int Q>
int dispatch3_runtime(const FusedArgs& A, cudaStream_t st, const char** err)
{
   const Params& p = A.p;
   const bool symm = p.symm;
   if (symm && Q != 4) {
      *err = "symmetry needs qlen=4 in this build";
      return AMPE_EINVAL;
   }
   constexpr bool S4 = (Q == 4);
   switch (p.conc_form) {
      case 0:
      case AMPE_CONC_CAHN_HILLIARD:
         if (p.with_T) {
            if (symm) return launch_any<ND, Q, 0, S4, true, SelRuntime>(A, st, err);
            return launch_any<ND, Q, 0, false, true, SelRuntime>(A, st, err);
         }
         if (symm) return launch_any<ND, Q, 0, S4, false, SelRuntime>(A, st, err);
         return launch_any<ND, Q, 0, false, false, SelRuntime>(A, st, err);
      case AMPE_CONC_KKS:
         if (symm) return launch_any<ND, Q, AMPE_CONC_KKS, S4, false, SelRuntime>(A, st, err);
         return launch_any<ND, Q, AMPE_CONC_KKS, false, false, SelRuntime>(A, st, err);
      case AMPE_CONC_EBS:
         if (symm) return launch_any<ND, Q, AMPE_CONC_EBS, S4, false, SelRuntime>(A, st, err);
         return launch_any<ND, Q, AMPE_CONC_EBS, false, false, SelRuntime>(A, st, err);
   }
   *err = "unknown conc_rhs_form";
   return AMPE_EINVAL;
}

}  // namespace ampe
