// Fused evaluateRHSFunction kernel, 3D plane-marching form.
//
// A block owns a 32 x TY column of cells and marches NZ planes along z (the slab axis).
// Shared memory holds a ring of four staged planes (each 34 x (TY+2) per field, i.e. with
// the in-plane halo incl. corners): planes k-1, k, k+1 are what the stencils of plane k
// read, plane k+2 is in flight (cp.async issued one step ahead, so HBM latency is hidden
// behind the arithmetic of a whole plane and the halo is re-read only in x,y: 1.33x for
// TY=8 against 2.4x for the 32x4x4 tile kernel).  Per step every thread computes the lower
// x and y faces of its cell (exchanged through shared memory, the two tile-edge rows by one
// warp each) and the UPPER z face, which it keeps in registers: next step it is the lower z
// face of the same thread's next cell, so z fluxes never touch shared memory.
//
// Arithmetic: rhs_math.cuh (same functions as the tile kernel; z neighbours are addressed
// through the ring-slot offsets ZOff instead of a constant stride).
// fillScratch (QuatIntegrator.cc:2873-2955) = periodic wrap in x,y while staging + ghost
// planes along z (opposite interior planes on one rank, halo buffers on several).
#pragma once
#include "rhs_math.cuh"
#include "tile_shape.h"

namespace ampe {

// PART selects the outputs of one launch (experiment AMPE_B200_SPLIT3D, EBS models only):
//   0  everything (one launch per evaluation, the default)
//   1  phase, quaternion (and temperature) RHS: no composition flux, c_l / c_a are not staged
//   2  composition RHS only: instantiated with Q_ = 0, the quaternions are not staged
// Every output is computed by the same expressions in either form (bit-identical); two light
// kernels trade re-staged phi planes (HBM has 4x headroom) for registers and resident warps.
template <int Q_, int CONC_, bool WT_, class SEL_, int TY_, int NZ_, int PART_ = 0>
struct March3 {
   // AMPE_MARCH_EDGE_WARP: one extra warp per block computes the tile's upper edge faces (x = TX: TY faces, y = TY:
   // 32 faces) and nothing else, so that no warp of the block carries a fourth face per plane (the block barrier
   // after the faces waits for the slowest warp)
   // (measured, profiles/r02j_ab.log: AuNi_3D / EBS 13.80 -> 13.38 ms, GG3D / KKS 4.82 -> 5.06 ms: on for the fused
   //  EBS instantiation, whose faces carry the CALPHAD mobilities; -DAMPE_MARCH_EDGE_WARP=0 / 1 forces it)
#ifdef AMPE_MARCH_EDGE_WARP
   static constexpr int EDGE_WARP = AMPE_MARCH_EDGE_WARP;
#else
   static constexpr int EDGE_WARP = (CONC_ == AMPE_CONC_EBS && PART_ == 0) ? 1 : 0;
#endif
   static constexpr int ND = 3, Q = Q_, CONC = CONC_, TX = 32, TY = TY_, NZ = NZ_, NT = 32 * (TY_ + EDGE_WARP);
   // phase flux through the faces (3D anisotropic interface energy): the runtime-selector instantiations carry it,
   // the compile-time selector sets of the shipped decks use the simple stencil
   static constexpr bool SYMM = false, WT = WT_, HAS_PF = !SEL_::fixed && (Q_ == 4);
   static constexpr int PART = PART_;
   // resident blocks per SM the register allocation is capped for (tile_shape.h)
   static constexpr int MINB = (PART_ == 0) ? AMPE_MARCH_MINB : ((PART_ == 1) ? AMPE_SPLIT_MINB1 : AMPE_SPLIT_MINB2);
   static constexpr int CF = (PART_ == 1) ? 0 : CONC_;  // composition-flux form this launch evaluates
   static_assert(PART_ == 0 || CONC_ == AMPE_CONC_EBS, "split launches: EBS composition model only");
   static_assert(PART_ != 2 || Q_ == 0, "the composition part does not stage the quaternions");
   using SEL = SEL_;
   static constexpr int SX = TX + 2, SYP = TY + 2;
   static constexpr int SP = SX * SYP;  // staged cells of one plane
   static constexpr int S = SP;         // field stride inside a ring slot
   static constexpr int O_PHI = 0;
   static constexpr int O_T = SP;
   static constexpr int O_Q = O_T + (WT ? SP : 0);
   static constexpr int O_C = O_Q + Q * SP;
   static constexpr int O_CL = O_C + (CF == AMPE_CONC_KKS ? SP : 0);
   static constexpr int O_CA = O_CL + (CF != 0 ? SP : 0);
   static constexpr int SLOT = O_CA + (CF != 0 ? SP : 0);  // doubles per ring slot
   static constexpr int NSLOT = 4;
   // in-plane face values, indexed like the staged plane (lower face of staged cell c)
   static constexpr int O_FXQ = NSLOT * SLOT, O_FYQ = O_FXQ + SP, O_FXC = O_FYQ + SP, O_FYC = O_FXC + SP;
   static constexpr int O_FXP = O_FYC + SP, O_FYP = O_FXP + (HAS_PF ? SP : 0);  // phase flux (HAS_PF)
   static constexpr int O_END = O_FYP + (HAS_PF ? SP : 0);
   static constexpr size_t SMEM_BYTES = (size_t)O_END * sizeof(double);
   static constexpr int NE = (SP + NT - 1) / NT;  // staged elements per thread and field
};

template <class TT>
__global__ void __launch_bounds__(TT::NT, TT::MINB) rhs_march_kernel(const __grid_constant__ FusedArgs A)
{
   using R = Rhs3<TT>;
   using SEL = typename TT::SEL;
   constexpr int Q = TT::Q, CONC = TT::CF, NT = TT::NT, TX = TT::TX, TY = TT::TY, SX = TT::SX;
   constexpr int SP = TT::SP, SLOT = TT::SLOT, NE = TT::NE;
   constexpr bool WT = TT::WT;
   const Params& p = A.p;
   extern __shared__ double smem[];
   const int* s_iq = nullptr;
   const double(*s_qr)[4] = nullptr;
   const int* s_conj = nullptr;

   const int n0 = p.n[0], n1 = p.n[1], n2 = p.n[2];
   const int ns = n2;
   const long long plane = (long long)n0 * n1;
   const long long ncell = plane * n2;
   const int lane = threadIdx.x % 32, row = threadIdx.x / 32;
   const int ox = blockIdx.x * TX, oy = blockIdx.y * TY;
   const int z0 = A.s_begin + blockIdx.z * TT::NZ;
   const int zend = min(z0 + TT::NZ, A.s_end);
   const bool own = (row < TY);                // false: the edge warp (AMPE_MARCH_EDGE_WARP)
   const int c = (lane + 1) + SX * (row + 1);  // this thread's cell inside a staged plane

   // ---- staging descriptors: the same in-plane elements every plane ------------------------
   int e_d[NE], e_ip[NE];
#pragma unroll
   for (int n = 0; n < NE; n++) {
      int e = threadIdx.x + n * NT;
      e = (e < SP) ? e : -1;
      const int ee = (e < 0) ? 0 : e;
      int gx = ox - 1 + ee % SX, gy = oy - 1 + ee / SX;
      if (AMPE_CLAMP(0)) {  // zero-slope boundary: the ghost cell is the adjacent interior cell
         gx = (gx < 0) ? 0 : ((gx >= n0) ? n0 - 1 : gx);
      } else {
         gx %= n0;
         gx = (gx < 0) ? gx + n0 : gx;
      }
      if (AMPE_CLAMP(1)) {
         gy = (gy < 0) ? 0 : ((gy >= n1) ? n1 - 1 : gy);
      } else {
         gy %= n1;
         gy = (gy < 0) ? gy + n1 : gy;
      }
      e_d[n] = e;
      e_ip[n] = gx + n0 * gy;
   }
   const int src_lo = AMPE_CLAMP(2) ? 0 : ns - 1, src_hi = AMPE_CLAMP(2) ? ns - 1 : 0;
   // cp.async of slab plane sl (-1 .. ns) into ring slot `slot`
   auto load_plane = [&](int sl, int slot) {
      double* dst = smem + slot * SLOT;
      sl = (sl > ns) ? ns : sl;
      const long long og = (long long)(sl + 1) * plane;  // slab-ghosted ctx arrays
      const double *b_phi, *b_T, *b_q, *b_c;
      long long qcomp;
      if (A.wrap_slab) {
         // one rank: the ghost planes are the opposite interior planes of the same array
         // ghost width 1: plane -1 / ns is the opposite interior plane (periodic) or the adjacent one (zero slope)
         const int slw = (sl < 0) ? src_lo : ((sl >= ns) ? src_hi : sl);
         const long long o = (long long)slw * plane;
         b_phi = A.phi.base + o;
         b_T = WT ? A.T.base + o : nullptr;
         b_q = (Q > 0) ? A.q.base + o : nullptr;
         b_c = (CONC == AMPE_CONC_KKS) ? A.conc.base + o : nullptr;
         qcomp = A.q.comp;
      } else {
         // slab neighbours' planes live in separate halo buffers (ampe_rhs_set_halo)
         const int region = (sl < 0) ? 1 : ((sl >= ns) ? 2 : 0);
         const long long o = (long long)((region == 0) ? sl : ((region == 1) ? sl + 1 : sl - ns)) * plane;
         auto sel = [&](const Field& f) { return ((region == 0) ? f.base : ((region == 1) ? f.lo : f.hi)) + o; };
         b_phi = sel(A.phi);
         b_T = WT ? sel(A.T) : nullptr;
         b_q = (Q > 0) ? sel(A.q) : nullptr;
         b_c = (CONC == AMPE_CONC_KKS) ? sel(A.conc) : nullptr;
         qcomp = (region == 0) ? A.q.comp : A.q.hcomp;
      }
#pragma unroll
      for (int n = 0; n < NE; n++) {
         const int d = e_d[n], ip = e_ip[n];
         if (d >= 0) {
            cp_async8(dst + TT::O_PHI + d, b_phi + ip);
            if (WT) cp_async8(dst + TT::O_T + d, b_T + ip);
#pragma unroll
            for (int m = 0; m < Q; m++) cp_async8(dst + TT::O_Q + m * SP + d, b_q + m * qcomp + ip);
            if (CONC == AMPE_CONC_KKS) cp_async8(dst + TT::O_C + d, b_c + ip);
            if (CONC != 0) {
               cp_async8(dst + TT::O_CL + d, A.cl + og + ip);
               cp_async8(dst + TT::O_CA + d, A.ca + og + ip);
            }
         }
      }
   };

   // ---- global bookkeeping of this thread's column ---------------------------------------------
   const int gi = ox + lane, gj = oy + row;
   const bool in_i = gi < n0, in_j = own && gj < n1;
   const bool inr_x = (gi <= n0) && in_j;  // lower x face bounds a cell of the domain
   const bool inr_y = in_i && (gj <= n1);
   const bool inr_c = in_i && in_j;
   const long long col = gi + (long long)n0 * gj;
   const long long wrap_x = (gi == n0) ? n0 : 0;       // faces of overhanging cells wrap periodically
   const long long wrap_y = (gj == n1) ? plane : 0;
   // tile-edge faces: x = TX handled by warp 0 (lanes < TY), y = TY by the last warp
   const bool edge_x = (row == (TT::EDGE_WARP ? TY : 0)) && (lane < TY);
   const bool edge_y = (row == (TT::EDGE_WARP ? TY : TY - 1));
   const int cex = (TX + 1) + SX * (lane + 1);  // staged cell whose lower x face is the tile's x edge
   const int cey = (lane + 1) + SX * (TY + 1);
   bool inr_ex = false, inr_ey = false;
   long long col_ex = 0, col_ey = 0;
   {
      int g = ox + TX;
      const int gje = oy + lane;
      inr_ex = (g - 1 < n0) && (gje < n1);
      g = (g >= n0) ? g % n0 : g;
      col_ex = g + (long long)n0 * gje;
      int gjy = oy + TY;
      inr_ey = in_i && (gjy - 1 < n1);
      gjy = (gjy >= n1) ? gjy % n1 : gjy;
      col_ey = gi + (long long)n0 * gjy;
   }

   double* fxq = smem + TT::O_FXQ;
   double* fyq = smem + TT::O_FYQ;
   double* fxc = smem + TT::O_FXC;
   double* fyc = smem + TT::O_FYC;
   double* fxp = smem + TT::O_FXP;
   double* fyp = smem + TT::O_FYP;
   (void)fxp;
   (void)fyp;
   const bool evolve_quat = (Q > 0) && AMPE_SEL(evolve_quat);
   (void)evolve_quat;

   // slab ranks: the column's first / last block stages the neighbours' ghost planes
   wait_ghost_planes(A, z0 == 0, zend >= ns);
   // ---- prologue: planes z0-1 and z0 (z0+1 is prefetched by the first step) -------------------
   load_plane(z0 - 1, 1);
   load_plane(z0, 2);

   double fzq_lo = 0.0, fzc_lo = 0.0, fzp_lo = 0.0;  // lower z face of the current cell (registers)
   int j = 0;                          // ring slot of plane k-1
   // step k = z0-1 only produces the z face between planes z0-1 and z0
#pragma unroll 1
   for (int k = z0 - 1; k < zend; ++k, j = (j + 1) & 3) {
      cp_async_wait_all();
      __syncthreads();  // plane k+1 has landed; every thread is done with step k-1
      if (k + 2 <= zend) load_plane(k + 2, (j + 3) & 3);
      const double* sk = smem + ((j + 1) & 3) * SLOT;  // plane k
      ZOff z;
      z.m = (j - ((j + 1) & 3)) * SLOT;
      z.p = (((j + 2) & 3) - ((j + 1) & 3)) * SLOT;
      const bool active = (k >= z0);
      const long long gcell = col + (long long)k * plane;
      if (active && own) {
         // lower x / y faces of the own cell
         {
            const FaceVal v = R::template face<0>(A, sk, s_iq, s_qr, s_conj, c, c - 1, z, gcell - wrap_x, inr_x,
                                                  A.write_lag && inr_x);
            if (Q > 0) fxq[c] = v.fc;
            if (CONC != 0) fxc[c] = v.cf;
            if constexpr (TT::HAS_PF) fxp[c] = v.pf;
         }
         {
            const FaceVal v = R::template face<1>(A, sk, s_iq, s_qr, s_conj, c, c - SX, z, gcell - wrap_y, inr_y,
                                                  A.write_lag && inr_y);
            if (Q > 0) fyq[c] = v.fc;
            if (CONC != 0) fyc[c] = v.cf;
            if constexpr (TT::HAS_PF) fyp[c] = v.pf;
         }
      }
      if (active) {
         // upper edge faces of the tile (they belong to the neighbouring column)
         if (edge_x) {
            const FaceVal v = R::template face<0>(A, sk, s_iq, s_qr, s_conj, cex, cex - 1, z,
                                                  col_ex + (long long)k * plane, inr_ex, false);
            if (Q > 0) fxq[cex] = v.fc;
            if (CONC != 0) fxc[cex] = v.cf;
            if constexpr (TT::HAS_PF) fxp[cex] = v.pf;
         }
         if (edge_y) {
            const FaceVal v = R::template face<1>(A, sk, s_iq, s_qr, s_conj, cey, cey - SX, z,
                                                  col_ey + (long long)k * plane, inr_ey, false);
            if (Q > 0) fyq[cey] = v.fc;
            if (CONC != 0) fyc[cey] = v.cf;
            if constexpr (TT::HAS_PF) fyp[cey] = v.pf;
         }
      }
      // upper z face: between plane k (lower) and k+1; index of the lower face of cell k+1.
      // The block above recomputes the same value in its first step (identical bits).
      FaceVal vz;
      vz.fc = 0.0, vz.pf = 0.0, vz.cf = 0.0;
      if (own)
         vz = R::template face<2>(A, sk, s_iq, s_qr, s_conj, c + z.p, c, z, gcell + plane, inr_c, A.write_lag && inr_c);
      __syncthreads();  // in-plane faces visible
      if (active && inr_c) {
         CellFaces<3> F;
         F.fcl[0] = (Q > 0) ? fxq[c] : 0.0;
         F.fcu[0] = (Q > 0) ? fxq[c + 1] : 0.0;
         F.fcl[1] = (Q > 0) ? fyq[c] : 0.0;
         F.fcu[1] = (Q > 0) ? fyq[c + SX] : 0.0;
         F.fcl[2] = fzq_lo;
         F.fcu[2] = vz.fc;
         F.cfl[0] = (CONC != 0) ? fxc[c] : 0.0;
         F.cfu[0] = (CONC != 0) ? fxc[c + 1] : 0.0;
         F.cfl[1] = (CONC != 0) ? fyc[c] : 0.0;
         F.cfu[1] = (CONC != 0) ? fyc[c + SX] : 0.0;
         F.cfl[2] = fzc_lo;
         F.cfu[2] = vz.cf;
         F.pfl[0] = TT::HAS_PF ? fxp[c] : 0.0;
         F.pfu[0] = TT::HAS_PF ? fxp[c + 1] : 0.0;
         F.pfl[1] = TT::HAS_PF ? fyp[c] : 0.0;
         F.pfu[1] = TT::HAS_PF ? fyp[c + SX] : 0.0;
         F.pfl[2] = fzp_lo;
         F.pfu[2] = vz.pf;
         R::cell(A, sk, s_iq, s_qr, s_conj, c, z, F, gcell, ncell);
      }
      fzq_lo = vz.fc;
      fzc_lo = vz.cf;
      fzp_lo = vz.pf;
   }
}

}  // namespace ampe
