// Fused evaluateRHSFunction kernel, tile form (all 2D models; 3D with quaternion symmetry or
// when the marching kernel does not apply).
//
// One launch computes the phase, orientation, composition and temperature right-hand sides
// of a tile of cells from a shared-memory stage of the state fields (1-cell halo incl.
// edges/corners): (A) stage tile+halo in smem with cp.async; (B) every FACE of the tile
// exactly once -- each thread computes the lower faces of the cells it owns, the tile's upper
// boundary faces are one extra pass of two warps -- face coefficient, phase flux, composition
// flux -> smem; (C) every CELL: divergences + pointwise terms -> global.  The arithmetic is in
// rhs_math.cuh.  fillScratch (QuatIntegrator.cc:2873-2955) is the periodic wrap / halo-plane
// selection done while staging.
#pragma once
#include <type_traits>

#include "rhs_math.cuh"

// the per-thread row loops of the face and cell phases: rolled (small code) by default, -DAMPE_TILE_UNROLL_ROWS
// unrolls them (A/B builds: more independent work per thread, more registers)
// Measured (profiles/r02k_ab.log): unrolled, the symmetry-aware AuNi_2D kernel gains 3 % (1.905 -> 1.847 ms), the
// Dendrite2D kernel loses 3 % (0.133 -> 0.137 ms): unrolled for the symmetry-aware instantiations only.
#ifdef AMPE_TILE_UNROLL_ROWS
#define AMPE_TILE_ROW_UNROLL _Pragma("unroll")
#else
#define AMPE_TILE_ROW_UNROLL _Pragma("unroll(TT::SYMM ? TT::CPT : 1)")
#endif

namespace ampe {

// ---- tile geometry + shared-memory carve-up -----------------------------------------------
// XH: staged halo width in x.  1 is what the stencils need; the TMA kernel stages 2 so that every
// box starts on an even column (cp.async.bulk.tensor faults on a box start that is not 16-byte
// aligned: measured, tools/probe/tma_probe3.cu).
template <int ND_, int Q_, int CONC_, bool SYMM_, bool WT_, class SEL_, int TX_, int TY_, int TZ_, int NT_, int XH_ = 1>
struct Tile3 {
   static constexpr int ND = ND_, Q = Q_, CONC = CONC_, TX = TX_, TY = TY_, TZ = TZ_, NT = NT_;
   static constexpr bool SYMM = SYMM_, WT = WT_;
   static constexpr int PART = 0, CF = CONC_;  // the tile kernels always evaluate every component
   using SEL = SEL_;
   static constexpr int HZ = (ND == 3) ? 1 : 0;
   static constexpr int XH = XH_;
   static constexpr int SX = TX + 2 * XH, SY = TY + 2, SZ = TZ + 2 * HZ;
   // doubles per staged field, padded to 128 B: a TMA box may land on every field
   static constexpr int S = (SX * SY * SZ + 15) / 16 * 16;
   static constexpr int FX = TX + 1, FY = TY + 1, FZ = TZ + HZ;
   static constexpr int NFB = FX * FY * FZ;
   static constexpr int NW = NT / 32;          // warps
   static constexpr int ROWS = TY * TZ;        // tile rows of 32 cells
   static constexpr int CPT = ROWS / NW;       // rows (cells) per thread
   static_assert(TX == 32 && ROWS % NW == 0, "one warp per tile row");
   static_assert(TY % NW == 0 || NW % TY == 0, "row stride must stay inside a plane");
   // staged field offsets (doubles)
   static constexpr int O_PHI = 0;
   static constexpr int O_T = S;
   static constexpr int O_Q = O_T + (WT ? S : 0);
   static constexpr int O_C = O_Q + Q * S;
   static constexpr int O_CL = O_C + (CONC == AMPE_CONC_KKS ? S : 0);
   static constexpr int O_CA = O_CL + (CONC != 0 ? S : 0);
   static constexpr int O_FC = O_CA + (CONC != 0 ? S : 0);  // quaternion face coefficient
   static constexpr bool HAS_PF = (ND == 2) && (!SEL::fixed || SEL::flux_type != AMPE_FLUX_SIMPLE);
   static constexpr int O_PF = O_FC + (Q > 0 ? ND * NFB : 0);
   static constexpr int O_CF = O_PF + (HAS_PF ? ND * NFB : 0);
   static constexpr int O_END = O_CF + (CONC != 0 ? ND * NFB : 0);
   static constexpr size_t SMEM_BYTES = (size_t)O_END * sizeof(double) + (SYMM ? (size_t)ND * S * sizeof(int) : 0);
   // the same face arrays relative to O_FC (kernels that keep them outside the staged fields)
   static constexpr int F_FC = 0, F_PF = O_PF - O_FC, F_CF = O_CF - O_FC, F_END = O_END - O_FC;
   // staged strides / face-box strides per direction
   __host__ __device__ static constexpr int str(int a) { return a == 0 ? 1 : (a == 1 ? SX : SX * SY); }
   __host__ __device__ static constexpr int ftr(int a) { return a == 0 ? 1 : (a == 1 ? FX : FX * FY); }
   AMPE_DEV static int sidx(int i, int j, int k) { return (i + XH) + SX * ((j + 1) + SY * (k + HZ)); }
   AMPE_DEV static int fidx(int i, int j, int k) { return i + FX * (j + FY * k); }
};

// ---- (A) stage a tile (+1-cell halo incl. edges/corners) of every state field ---------------
// cp.async straight into shared memory (no register staging).  fillScratch
// (QuatIntegrator.cc:2873-2955) = the periodic wrap / halo-plane selection done here.
template <class TT>
AMPE_DEV void stage_tile(const FusedArgs& A, double* s, int* s_iq, int ox, int oy, int oz)
{
   constexpr int ND = TT::ND, Q = TT::Q, CONC = TT::CONC, S = TT::S, NT = TT::NT, NW = TT::NW;
   constexpr bool SYMM = TT::SYMM, WT = TT::WT;
   const Params& p = A.p;
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;
   const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
   (void)n1;
   // A staged row has SX = 32 + 2 XH elements: one warp copies x = 0..31 of a row per instruction,
   // the tail elements of all rows are gathered into one extra pass of the block.
   // Index arithmetic is split by what it depends on: the wrapped global column of a staged x (one modulo per
   // thread, outside the row loop) and the per-row plane offsets (no division: a row index only needs a
   // compare-and-shift wrap); copy_elem adds the two.  (r01c: staging and index arithmetic were 21.6 % of the
   // Dendrite2D kernel's instructions, two modulos per staged element.)
   {
      constexpr int NROWS = TT::SY * TT::SZ;
      auto wrap_col = [&](int xs) {
         int gx = ox - TT::XH + xs;
         if (AMPE_CLAMP(0)) return (gx < 0) ? 0 : ((gx >= n0) ? n0 - 1 : gx);  // zero-slope boundary
         gx %= n0;
         return (gx < 0) ? gx + n0 : gx;
      };
      // element (row r, staged x) of every staged field; gx = wrap_col(xs)
      auto copy_elem = [&](int r, int xs, int gx) {
         const int lj = r % TT::SY - 1;
         const int lk = (ND == 3) ? (r / TT::SY - 1) : 0;
         int sl;
         int inplane = gx;  // offset inside a slab plane
         if (ND == 3) {
            int gj = oy + lj;  // in [-1, n1 + TY]: tiles overhang the domain by less than one tile
            if (AMPE_CLAMP(1)) {
               gj = (gj < 0) ? 0 : ((gj >= n1) ? n1 - 1 : gj);
            } else {
               gj = (gj < 0) ? gj + n1 : gj;
               gj = (gj >= n1) ? gj % n1 : gj;
            }
            sl = oz + lk;
            inplane += n0 * gj;
         } else {
            sl = oy + lj;
         }
         // tiles may overhang the domain: rows beyond the upper ghost plane are only read by
         // out-of-range cells, clamp them onto the ghost plane
         sl = (sl > ns) ? ns : sl;
         const int d = r * TT::SX + xs;
         const long long og = (long long)(sl + 1) * plane + inplane;  // slab-ghosted ctx arrays
         if (A.wrap_slab) {
            // one rank: the ghost planes are the opposite interior planes of the same array
            // (or, with a zero-slope boundary along the slab axis, the adjacent interior plane)
            const int slw = AMPE_CLAMP(ND - 1) ? ((sl < 0) ? 0 : ((sl >= ns) ? ns - 1 : sl))
                                            : ((sl < 0) ? sl + ns : ((sl >= ns) ? sl - ns : sl));
            const long long o = (long long)slw * plane + inplane;
            cp_async8(s + TT::O_PHI + d, A.phi.base + o);
            if (WT) cp_async8(s + TT::O_T + d, A.T.base + o);
#pragma unroll
            for (int m = 0; m < Q; m++) cp_async8(s + TT::O_Q + m * S + d, A.q.base + m * A.q.comp + o);
            if (CONC == AMPE_CONC_KKS) cp_async8(s + TT::O_C + d, A.conc.base + o);
         } else {
            // slab neighbours' planes live in separate halo buffers (ampe_rhs_set_halo)
            const int region = (sl < 0) ? 1 : ((sl >= ns) ? 2 : 0);
            const long long o = (long long)((region == 0) ? sl : ((region == 1) ? sl + 1 : sl - ns)) * plane + inplane;
            auto src = [&](const Field& f, int m) {
               const double* b = (region == 0) ? f.base : ((region == 1) ? f.lo : f.hi);
               return b + m * ((region == 0) ? f.comp : f.hcomp) + o;
            };
            cp_async8(s + TT::O_PHI + d, src(A.phi, 0));
            if (WT) cp_async8(s + TT::O_T + d, src(A.T, 0));
#pragma unroll
            for (int m = 0; m < Q; m++) cp_async8(s + TT::O_Q + m * S + d, src(A.q, m));
            if (CONC == AMPE_CONC_KKS) cp_async8(s + TT::O_C + d, src(A.conc, 0));
         }
         if (CONC != 0) {
            cp_async8(s + TT::O_CL + d, A.cl + og);
            cp_async8(s + TT::O_CA + d, A.ca + og);
         }
         if (SYMM) {
#pragma unroll
            for (int a = 0; a < ND; a++) cp_async4(s_iq + a * S + d, A.iq[a] + og);
         }
      };
      const int gx_lane = wrap_col(lane);
#pragma unroll 1
      for (int r = warp; r < NROWS; r += NW) copy_elem(r, lane, gx_lane);
      // tail: (SX - 32) elements per row; consecutive threads take the tail elements of one row
      constexpr int NTAIL = TT::SX - 32;
#pragma unroll 1
      for (int t = threadIdx.x; t < NTAIL * NROWS; t += NT) {
         const int xs = 32 + t % NTAIL;
         copy_elem(t / NTAIL, xs, wrap_col(xs));
      }
      cp_async_wait_all();
   }
}

// ---- (B) faces + (C) cells of one staged tile ------------------------------------------------
// s: staged fields (Tile3 offsets), sf: face arrays (F_FC / F_PF / F_CF), origin (ox, oy, oz).
// Ends with the cells' global stores; the caller synchronises before s / sf are reused.
template <class TT>
AMPE_DEV void tile_compute(const FusedArgs& A, const double* s, double* sf, const int* s_iq,
                           const double (*s_qr)[4], const int* s_conj, int ox, int oy, int oz)
{
   using R = Rhs3<TT>;
   using SEL = typename TT::SEL;
   constexpr int ND = TT::ND, Q = TT::Q, CONC = TT::CONC, NT = TT::NT, NW = TT::NW;
   constexpr int TX = TT::TX, TY = TT::TY, TZ = TT::TZ, CPT = TT::CPT;
   const Params& p = A.p;
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;  // planes along the slab axis
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;  // slab plane size
   const long long ncell = (long long)n0 * n1 * n2;
   const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;

   // rows owned by this warp: r = warp + u*NW; inside a plane the row stride is constant
   constexpr int RSTEP_J = (NW < TY) ? NW : 0;            // rows advance in y ...
   constexpr int RSTEP_K = (NW < TY) ? 0 : NW / TY;       // ... or in z
   const int lj0 = warp % TY, lk0 = warp / TY;
   constexpr int CSTEP = RSTEP_J * TT::SX + RSTEP_K * TT::SX * TT::SY;
   constexpr int FSTEP = RSTEP_J * TT::FX + RSTEP_K * TT::FX * TT::FY;
   static_assert(NW >= TY || TZ == 1 || TY % NW == 0, "row ownership");
   // NW < TY in 3D would wrap rows across planes; only allowed when CPT rows stay in one plane
   static_assert(!(ND == 3 && NW < TY && CPT * NW > TY), "3D: a thread's rows must stay in one plane");

   // per-direction bookkeeping of a face: does it bound a cell of the domain, and where does it
   // live in the lagged arrays (x and in-plane y wrap periodically; the slab axis has ns+1 planes)
   auto face_meta = [&](int a, int gi, int gj, int gk, bool& inr, long long& gface) {
      inr = (gi - (a == 0) < n0) && (gj - (a == 1) < n1) && (gk - (a == 2) < n2);
      if (gi >= n0) gi %= n0;
      if (ND == 3 && gj >= n1) gj %= n1;
      gface = gi + (long long)n0 * (gj + (long long)n1 * gk);
   };

   // run one face and store what it produced in the face box of the tile
   constexpr ZOff ZT = {-TT::SX * TT::SY, TT::SX * TT::SY};
   auto do_face = [&](auto dir, int c, int f, long long gface, bool inr, bool wr) {
      constexpr int a = decltype(dir)::value;
      const FaceVal v = R::template face<a>(A, s, s_iq, s_qr, s_conj, c, c - TT::str(a), ZT, gface, inr, wr);
      if (Q > 0) sf[TT::F_FC + a * TT::NFB + f] = v.fc;
      if constexpr (TT::HAS_PF && a < 2) sf[TT::F_PF + a * TT::NFB + f] = v.pf;
      if (CONC != 0) sf[TT::F_CF + a * TT::NFB + f] = v.cf;
   };
   using D0 = std::integral_constant<int, 0>;
   using D1 = std::integral_constant<int, 1>;
   using D2 = std::integral_constant<int, (ND == 3 ? 2 : 0)>;

   const bool need_faces = (Q > 0 && AMPE_SEL(evolve_quat)) ||
                           (TT::HAS_PF && AMPE_SEL(flux_type) != AMPE_FLUX_SIMPLE) || CONC != 0;
   if (need_faces) {
      // lower faces of the owned cells.  In the lagged arrays the lower face of cell (i,j,k) has
      // the cell's own index; only the faces of overhanging cells on the periodic upper edge wrap.
      {
         int c = TT::sidx(lane, lj0, lk0);
         int f = TT::fidx(lane, lj0, lk0);
         const int gi = ox + lane;
         int gj = oy + lj0, gk = oz + lk0;
         long long gcell = gi + (long long)n0 * (gj + (long long)n1 * gk);
         const long long gstep = (long long)RSTEP_J * n0 + (long long)RSTEP_K * plane;
AMPE_TILE_ROW_UNROLL
         for (int u = 0; u < CPT; u++) {
            const bool in_i = gi < n0, in_j = gj < n1, in_k = gk < n2;
            {
               const bool inr = (gi <= n0) && in_j && in_k;
               do_face(D0(), c, f, gcell - ((gi == n0) ? n0 : 0), inr, A.write_lag && inr);
            }
            {
               const bool inr = in_i && (gj <= n1) && in_k;
               do_face(D1(), c, f, gcell - ((ND == 3 && gj == n1) ? plane : 0), inr, A.write_lag && inr);
            }
            if constexpr (ND == 3) {
               const bool inr = in_i && in_j && (gk <= n2);
               do_face(D2(), c, f, gcell, inr, A.write_lag && inr);
            }
            c += CSTEP;
            f += FSTEP;
            gj += RSTEP_J;
            gk += RSTEP_K;
            gcell += gstep;
         }
      }
      // upper boundary faces of the tile; they belong to the neighbouring tile except on the
      // extra plane ns of the slab axis, which this tile refreshes in the lagged arrays.
      // x = TX: TY*TZ faces (first warps)
#pragma unroll 1
      for (int e = threadIdx.x; e < TY * TZ; e += NT) {
         const int lj = e % TY, lk = e / TY;
         bool inr;
         long long gface;
         face_meta(0, ox + TX, oy + lj, oz + lk, inr, gface);
         do_face(D0(), TT::sidx(TX, lj, lk), TT::fidx(TX, lj, lk), gface, inr, false);
      }
      // y = TY: TX*TZ faces, on the last warps so that they run beside the x pass
#pragma unroll 1
      for (int e = NT - 1 - threadIdx.x; e < TX * TZ; e += NT) {
         const int li = e % TX, lk = e / TX;
         bool inr;
         long long gface;
         face_meta(1, ox + li, oy + TY, oz + lk, inr, gface);
         const bool top = (ND == 2) && (oy + TY == ns);
         do_face(D1(), TT::sidx(li, TY, lk), TT::fidx(li, TY, lk), gface, inr, A.write_lag && top && inr);
      }
      if constexpr (ND == 3) {
#pragma unroll 1
         for (int e = threadIdx.x; e < TX * TY; e += NT) {
            const int li = e % TX, lj = e / TX;
            bool inr;
            long long gface;
            face_meta(2, ox + li, oy + lj, oz + TZ, inr, gface);
            const bool top = (oz + TZ == ns);
            do_face(D2(), TT::sidx(li, lj, TZ), TT::fidx(li, lj, TZ), gface, inr, A.write_lag && top && inr);
         }
      }
   }
   __syncthreads();

   // ---- (C) cells -----------------------------------------------------------------
   {
      int c = TT::sidx(lane, lj0, lk0);
      int fb = TT::fidx(lane, lj0, lk0);
      const int gi = ox + lane;
      int gj = oy + lj0, gk = oz + lk0;
AMPE_TILE_ROW_UNROLL
      for (int u = 0; u < CPT; u++) {
         bool ok = (gi < n0) && (gj < n1) && (gk < n2);
         if (ND == 2) ok = ok && (gj < A.s_end);
         if (ND == 3) ok = ok && (gk < A.s_end);
         if (ok) {
            const long long gcell = gi + (long long)n0 * (gj + (long long)n1 * gk);
            CellFaces<ND> F;
#pragma unroll
            for (int a = 0; a < ND; a++) {
               const int fl = a * TT::NFB + fb, fu = fl + TT::ftr(a);
               F.fcl[a] = (Q > 0) ? sf[TT::F_FC + fl] : 0.0;
               F.fcu[a] = (Q > 0) ? sf[TT::F_FC + fu] : 0.0;
               F.cfl[a] = (CONC != 0) ? sf[TT::F_CF + fl] : 0.0;
               F.cfu[a] = (CONC != 0) ? sf[TT::F_CF + fu] : 0.0;
               F.pfl[a] = (TT::HAS_PF && a < 2) ? sf[TT::F_PF + fl] : 0.0;
               F.pfu[a] = (TT::HAS_PF && a < 2) ? sf[TT::F_PF + fu] : 0.0;
            }
            R::cell(A, s, s_iq, s_qr, s_conj, c, ZT, F, gcell, ncell);
         }
         c += CSTEP;
         fb += FSTEP;
         gj += RSTEP_J;
         gk += RSTEP_K;
      }
   }
}

template <class TT>
__global__ void __launch_bounds__(TT::NT, (TT::NT <= 256) ? (TT::SYMM ? 2 : (TT::SEL::fixed ? 4 : 3)) : 1) rhs_tile_kernel(const __grid_constant__ FusedArgs A)
{
   using R = Rhs3<TT>;
   using SEL = typename TT::SEL;
   constexpr int ND = TT::ND, Q = TT::Q, CONC = TT::CONC, S = TT::S, NT = TT::NT, NW = TT::NW;
   constexpr int TX = TT::TX, TY = TT::TY, TZ = TT::TZ, CPT = TT::CPT;
   constexpr bool SYMM = TT::SYMM, WT = TT::WT;
   const Params& p = A.p;
   extern __shared__ double smem[];
   double* s = smem;
   int* s_iq = reinterpret_cast<int*>(smem + TT::O_END);  // ND*S ints (SYMM)
   __shared__ double s_qr[SYMM ? 48 : 1][4];
   __shared__ int s_conj[SYMM ? 48 : 1];
   if (SYMM && Q == 4) {
      for (int t = threadIdx.x; t < 48 * 4; t += NT) s_qr[t / 4][t % 4] = A.qr[t];
      for (int t = threadIdx.x; t < 48; t += NT) s_conj[t] = A.conj[t];
   }

   // ---- tile origin -------------------------------------------------------------
   const int ox = blockIdx.x * TX;
   const int oy = blockIdx.y * TY + ((ND == 2) ? A.s_begin : 0);
   const int oz = (ND == 3) ? (blockIdx.z * TZ + A.s_begin) : 0;
   {
      // ghost planes along the slab axis this tile stages: rows / planes -1 and ns
      const int ns = (ND == 3) ? p.n[2] : p.n[1];
      const int o_s = (ND == 3) ? oz : oy, t_s = (ND == 3) ? TZ : TY;
      wait_ghost_planes(A, o_s == 0, o_s + t_s >= ns);
   }
   stage_tile<TT>(A, s, s_iq, ox, oy, oz);
   __syncthreads();
   tile_compute<TT>(A, s, s + TT::O_FC, s_iq, s_qr, s_conj, ox, oy, oz);
}

}  // namespace ampe
