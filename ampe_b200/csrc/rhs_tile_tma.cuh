// Fused evaluateRHSFunction kernel, 2D persistent form with TMA staging.
//
// Same arithmetic as rhs_tile_kernel (tile_compute, rhs_math.cuh); what changes is how a tile
// reaches shared memory.  The grid is persistent (resident blocks x SMs); every block walks the
// tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... with a two-stage ring of staged fields:
// while tile t is computed, ONE thread has already issued the TMA box copies
// (cp.async.bulk.tensor, one (TX+4) x (TY+2) box per field -- two halo columns, because a box
// must start on a 16-byte boundary -- completion counted in bytes on an mbarrier) of tile t + gridDim.x into the other stage.  That removes the per-element staging
// instructions (14 % of the issue slots of rhs_tile_kernel, profiles/r01b) and the per-tile
// prologue, and hides the HBM latency of a tile behind the arithmetic of the previous one
// instead of behind other resident blocks.
//
// TMA cannot wrap periodically, so only tiles whose halo lies inside the slab are loaded by
// TMA; tiles touching the periodic boundary / the slab ghost planes / an overhanging edge are
// staged by stage_tile (cp.async per element) when the block reaches them (~6 % of the tiles
// at 2048^2).
#pragma once
#include <cuda.h>

#include "rhs_tile.cuh"

namespace ampe {

struct alignas(64) TmaMaps {
   CUtensorMap phi, T, q, conc, cl, ca;  // 3-D maps (x, row, component)
};

AMPE_DEV void mbar_init(uint64_t* bar, unsigned count)
{
   const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
AMPE_DEV void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
   const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
AMPE_DEV void mbar_wait(uint64_t* bar, unsigned parity)
{
   const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
   asm volatile(
       "{\n"
       ".reg .pred p;\n"
       "WAIT_%=:\n"
       "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
       "@p bra DONE_%=;\n"
       "bra WAIT_%=;\n"
       "DONE_%=:\n"
       "}\n" ::"r"(a),
       "r"(parity)
       : "memory");
}
// one (box0 x box1 x 1) box of a 3-D tensor map into shared memory, completion on `bar`
AMPE_DEV void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
{
   const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
   const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
   asm volatile(
       "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(d),
       "l"(reinterpret_cast<uint64_t>(map)), "r"(b), "r"(c0), "r"(c1), "r"(c2)
       : "memory");
}

template <class TT>
struct TmaTile {
   static constexpr int STAGE = TT::O_FC;  // doubles of staged fields per ring slot (multiple of 16)
   static constexpr int NF = 1 + (TT::WT ? 1 : 0) + TT::Q + (TT::CONC == AMPE_CONC_KKS ? 1 : 0) + (TT::CONC != 0 ? 2 : 0);
   static constexpr unsigned BOX_BYTES = (unsigned)(TT::SX * TT::SY * sizeof(double));
   static constexpr size_t SMEM_BYTES = (size_t)(2 * STAGE + TT::F_END) * sizeof(double) + 2 * sizeof(uint64_t) + 128;
};

template <class TT>
__global__ void __launch_bounds__(TT::NT, (TT::NT >= 512) ? 2 : 3) rhs_tile_tma_kernel(const __grid_constant__ FusedArgs A,
                                                             const __grid_constant__ TmaMaps M, int tiles_x,
                                                             int ntiles)
{
   static_assert(TT::ND == 2 && !TT::SYMM && TT::XH == 2, "TMA form: 2D without quaternion symmetry, even box start");
   using TM = TmaTile<TT>;
   constexpr int Q = TT::Q, CONC = TT::CONC, S = TT::S, TX = TT::TX, TY = TT::TY;
   constexpr bool WT = TT::WT;
   const Params& p = A.p;
   extern __shared__ double smem_raw[];
   // 128-byte aligned carve-up: [stage 0][stage 1][face arrays][2 mbarriers]
   double* smem = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
   double* sf = smem + 2 * TM::STAGE;
   uint64_t* bar = reinterpret_cast<uint64_t*>(sf + TT::F_END);
   const int n0 = p.n[0], ns = p.n[1];

   if (threadIdx.x == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
   }
   __syncthreads();

   auto origin = [&](int t, int& ox, int& oy) {
      ox = (t % tiles_x) * TX;
      oy = (t / tiles_x) * TY + A.s_begin;
   };
   // the tile and its 1-cell halo lie inside the slab: no wrap, no ghost plane, no overhang
   auto inside = [&](int ox, int oy) { return ox >= 1 && ox + TX + 1 <= n0 && oy >= 1 && oy + TY + 1 <= ns; };
   auto issue = [&](int ox, int oy, int stage) {
      double* d = smem + stage * TM::STAGE;
      uint64_t* b = &bar[stage];
      mbar_expect_tx(b, TM::NF * TM::BOX_BYTES);
      tma_load_3d(d + TT::O_PHI, &M.phi, ox - 2, oy - 1, 0, b);
      if (WT) tma_load_3d(d + TT::O_T, &M.T, ox - 2, oy - 1, 0, b);
#pragma unroll
      for (int m = 0; m < Q; m++) tma_load_3d(d + TT::O_Q + m * S, &M.q, ox - 2, oy - 1, m, b);
      if (CONC == AMPE_CONC_KKS) tma_load_3d(d + TT::O_C, &M.conc, ox - 2, oy - 1, 0, b);
      if (CONC != 0) {
         // ctx-owned slab-ghosted arrays: row -1 of the slab is row 0 of the map
         tma_load_3d(d + TT::O_CL, &M.cl, ox - 2, oy, 0, b);
         tma_load_3d(d + TT::O_CA, &M.ca, ox - 2, oy, 0, b);
      }
   };

   int t = blockIdx.x;
   unsigned ph0 = 0u, ph1 = 0u;  // mbarrier phase parity per ring slot
   if (t < ntiles && threadIdx.x == 0) {
      int ox, oy;
      origin(t, ox, oy);
      if (inside(ox, oy)) issue(ox, oy, 0);
   }
#pragma unroll 1
   for (int it = 0; t < ntiles; t += gridDim.x, it++) {
      const int stage = it & 1;
      double* s = smem + stage * TM::STAGE;
      int ox, oy;
      origin(t, ox, oy);
      // prefetch the next tile of this block into the other slot (free since the barrier that
      // closed the previous iteration)
      const int tn = t + gridDim.x;
      if (tn < ntiles && threadIdx.x == 0) {
         int nx, ny;
         origin(tn, nx, ny);
         if (inside(nx, ny)) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            issue(nx, ny, stage ^ 1);
         }
      }
      if (inside(ox, oy)) {
         mbar_wait(&bar[stage], stage ? ph1 : ph0);
         if (stage)
            ph1 ^= 1u;
         else
            ph0 ^= 1u;
      } else {
         stage_tile<TT>(A, s, nullptr, ox, oy, 0);
         __syncthreads();
      }
      tile_compute<TT>(A, s, sf, nullptr, nullptr, nullptr, ox, oy, 0);
      __syncthreads();  // all threads are done with this slot and with the face arrays
   }
}

}  // namespace ampe
