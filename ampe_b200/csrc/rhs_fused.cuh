// Fused evaluateRHSFunction kernel: one launch computes the phase, orientation,
// composition and temperature right-hand sides of a tile of cells from a
// shared-memory stage of the state fields (1-cell halo incl. edges/corners).
//
// Replaces, per evaluation, the reference's unfused sweeps:
//   fillScratch                      QuatIntegrator.cc:2873-2955  (periodic wrap / halo planes on load)
//   quatdiffs, quatgrad_cell/side,   {2d,3d}/quatdiffs.m4, quatgrad.m4
//   quatgrad_modulus
//   quatmobility                     {2d,3d}/mobility.m4
//   gradient_flux / isotropic /      {2d,3d}/quatrhs.m4
//   anisotropic_gradient_flux, computerhspbg, computerhsbiaswell, computerhstemp
//   compute_face_coef, compute_flux_from_gradq, compute_lambda_flux,
//   add_quat_proj_op                 {2d,3d}/quatfacops.m4
//   correctrhsquatforsymmetry        {2d,3d}/correctrhsquatforsymmetry.m4
//   CALPHAD/Quadratic free energies + driving force, EBS face diffusivities,
//   add_flux / concentrationflux / concentration_pfmdiffusion, computerhsconcentration
//
// Structure: (A) stage tile+halo in smem with cp.async; (B) every FACE of the tile
// exactly once -- each thread computes the lower faces of the cells it owns, the
// tile's upper boundary faces are spread over the block -- anisotropic phase flux,
// quaternion face coefficient, composition flux -> smem; (C) every CELL:
// divergences + pointwise terms -> global.  Operation order inside each
// expression follows the Fortran so that results agree with the CPU restatement
// to rounding of the transcendental functions only.
#pragma once
#include "calphad.cuh"
#include "params.h"
#include "pointwise.cuh"

namespace ampe {

// caller-owned ghost-0 field + ghost planes along the slab axis (last axis)
struct Field {
   const double* base;
   const double* lo;  // ng planes below plane 0
   const double* hi;  // ng planes above plane ns-1
   long long comp;    // component stride in base
   long long hcomp;   // component stride in lo/hi
};

struct FusedArgs {
   Params p;
   Field phi, T, q, conc;
   // ctx-owned, slab-ghosted (ns+2 planes, pointer at plane -1): c_l, c_a
   const double* cl;
   const double* ca;
   const int* iq[3];  // slab-ghosted symmetry rotation indices (lower faces)
   const double* qr;  // 48x4 rotation table (setqr, quat.f:165-286), device global
   const int* conj;   // conjugate index table
   double* out_phi;
   double* out_q;
   double* out_c;
   double* out_T;
   // lagged face data (QuatIntegrator.cc:2804-2809, 3268-3269): 1/|grad q|_floor per
   // face and the composition face diffusivities, (ns+1) planes along the slab axis
   double* lagN[3];
   double* lagD0[3];
   double* lagD1[3];
   int use_lag;    // fd_flag != 0 && lag_quat_sidegrad
   int write_lag;  // refresh the lagged data
   int s_begin, s_end;  // slab-axis range of cells to compute [begin, end)
   int force_generic;   // host side only: use the runtime-selector instantiation
   int wrap_slab;       // 1: no halo buffers, ghost planes = opposite interior planes (one rank)
   const double* df;    // CALPHAD driving force (f_l-f_a)-mu(c_l-c_a) per cell from the KKS kernel
};

template <int Q>
AMPE_DEV void symm_rotate(const double* q, int iq, double* qp, const double (*s_qr)[4],
                          const int* s_conj)
{
   if (Q == 4) {
      if (iq < 0) iq = s_conj[-iq - 1];
      if (iq == 1) {
#pragma unroll
         for (int m = 0; m < 4; m++) qp[m] = q[m];
      } else {
         quatmult4(q, s_qr[iq - 1], qp);
      }
   } else if (Q == 2) {
      // quatsymmrotate2 (quat.f:449-520): rotations (1,0),(0,1),(-1,0),(0,-1); conj 1,4,3,2
      if (iq < 0) iq = (iq == -2) ? 4 : ((iq == -4) ? 2 : -iq);
      const double r0 = (iq == 1) ? 1.0 : ((iq == 3) ? -1.0 : 0.0);
      const double r1 = (iq == 2) ? 1.0 : ((iq == 4) ? -1.0 : 0.0);
      if (iq == 1) {
         qp[0] = q[0];
         qp[1] = q[1];
      } else {
         qp[0] = q[0] * r0 - q[1] * r1;
         qp[1] = q[0] * r1 + q[1] * r0;
      }
   }
}

// libm evaluation of the 2D anisotropy angle functions exactly as written in
// anisotropic_gradient_flux (2d/quatrhs.m4:192-214); selected with AMPE_B200_LIBM_TRIG=1
static __device__ __noinline__ void aniso_trig_libm(double dphidx, double dphidy, double qa,
                                                    int knumber, int qlen, double* sn, double* cs)
{
   double theta;
   if (fabs(dphidx) > (double)1.e-12f)
      theta = atan(dphidy / dphidx);
   else
      theta = 0.5 * 3.141592653589793;  // 4.d0*atan(1.d0)
   const double ang = (qlen == 4) ? 2.0 * acos(qa) : acos(qa);
   sincos(knumber * (theta - ang), sn, cs);
}

AMPE_DEV void cp_async8(void* smem_dst, const void* gmem_src)
{
   const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
   asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
}
AMPE_DEV void cp_async4(void* smem_dst, const void* gmem_src)
{
   const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
   asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}
AMPE_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int ND, int TX, int TY, int TZ>
struct TileGeom {
   static constexpr int HZ = (ND == 3) ? 1 : 0;
   static constexpr int SX = TX + 2, SY = TY + 2, SZ = TZ + 2 * HZ;
   static constexpr int S = SX * SY * SZ;    // staged cells
   static constexpr int NC = TX * TY * TZ;   // computed cells
   // face box: lower faces of the cells (0..TX, 0..TY, 0..TZ), one array per direction
   static constexpr int FX = TX + 1, FY = TY + 1, FZ = TZ + HZ;
   static constexpr int NFB = FX * FY * FZ;
   // faces on the upper boundary of the tile, not owned by a cell of the tile
   static constexpr int E0 = TY * TZ, E1 = TX * TZ, E2 = (ND == 3) ? TX * TY : 0;
   static constexpr int NE = E0 + E1 + E2;
   AMPE_DEV static int sidx(int i, int j, int k) { return (i + 1) + SX * ((j + 1) + SY * (k + HZ)); }
   AMPE_DEV static int fidx(int i, int j, int k) { return i + FX * (j + FY * k); }
};

// CONC: 0 none, 2 KKS(quadratic), 3 EBS(CALPHAD)  (AMPE_CONC_*)
template <int ND, int Q, int CONC, bool SYMM, int TX, int TY, int TZ, int NT>
__global__ void __launch_bounds__(NT)
    rhs_fused_kernel(const __grid_constant__ FusedArgs A)
{
   using G = TileGeom<ND, TX, TY, TZ>;
   constexpr int S = G::S;
   constexpr int NFB = G::NFB;
   constexpr int QN = (Q > 0) ? Q : 1;
   static_assert(TX == 32, "one warp per tile row");
   static_assert((TX * TY * TZ) % NT == 0, "cells per thread must be integral");
   constexpr int CPT = TX * TY * TZ / NT;  // cells owned by a thread
   constexpr int RP = NT / TX;             // tile rows per pass (= warps)
   const Params& p = A.p;
   extern __shared__ double smem[];

   // ---- shared memory carve-up ------------------------------------------------
   double* s_phi = smem;
   double* s_T = s_phi + S;                       // only if with_T
   double* s_q = s_T + (p.with_T ? S : 0);        // Q*S
   double* s_c = s_q + Q * S;                     // conc (KKS form)
   double* s_cl = s_c + (CONC == AMPE_CONC_KKS ? S : 0);
   double* s_ca = s_cl + (CONC != 0 ? S : 0);
   double* s_fc = s_ca + (CONC != 0 ? S : 0);     // quaternion face coefficient, ND*NFB
   double* s_pf = s_fc + ((Q > 0 && p.evolve_quat) ? ND * NFB : 0);  // phase flux (non-simple)
   double* s_cf = s_pf + (p.flux_type != AMPE_FLUX_SIMPLE ? ND * NFB : 0);  // composition flux
   double* s_end = s_cf + (CONC != 0 ? ND * NFB : 0);
   int* s_iq = reinterpret_cast<int*>(s_end);     // ND*S ints (SYMM)
   __shared__ double s_qr[SYMM ? 48 : 1][4];
   __shared__ int s_conj[SYMM ? 48 : 1];
   if (SYMM && Q == 4) {
      for (int t = threadIdx.x; t < 48 * 4; t += NT) s_qr[t / 4][t % 4] = A.qr[t];
      for (int t = threadIdx.x; t < 48; t += NT) s_conj[t] = A.conj[t];
   }

   // ---- tile origin -------------------------------------------------------------
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;  // planes along the slab axis
   const int ox = blockIdx.x * TX;
   const int oy = blockIdx.y * TY + ((ND == 2) ? A.s_begin : 0);
   const int oz = (ND == 3) ? (blockIdx.z * TZ + A.s_begin) : 0;
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;  // slab plane size
   const long long ncell = (long long)n0 * n1 * n2;
   const double Tuni = p.T_uniform;
   const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;

   // ---- (A) stage: one warp per staged row, cp.async (no register staging) ------------
   {
      // x indices of the two elements a lane copies: staged x = lane and lane + 32 (< SX)
      int gx0 = (ox + lane - 1) % n0;
      gx0 = (gx0 < 0) ? gx0 + n0 : gx0;
      const int gx1 = (ox + lane + 31) % n0;
      const bool second = lane < G::SX - 32;
      for (int r = warp; r < G::SY * G::SZ; r += RP) {
         const int lj = r % G::SY - 1;
         const int lk = (ND == 3) ? (r / G::SY - 1) : 0;
         int gj = oy + lj, sl;
         long long rowoff;
         if (ND == 3) {
            gj = gj % n1;
            gj = (gj < 0) ? gj + n1 : gj;
            sl = oz + lk;
            rowoff = (long long)n0 * gj;
         } else {
            sl = gj;
            rowoff = 0;
         }
         // tiles may overhang the domain: rows beyond the upper ghost plane are only read by
         // out-of-range cells, clamp them onto the ghost plane
         if (sl > ns) sl = ns;
         const bool below = sl < 0, above = sl >= ns;
         const long long og = (long long)(sl + 1) * plane + rowoff;  // slab-ghosted arrays
         const int srow = r * G::SX;
         auto row_copy = [&](const Field& f, int m, double* dst) {
            const double* src;
            if (below)
               src = f.lo + (long long)(sl + 1) * plane + rowoff + m * f.hcomp;
            else if (above)
               src = f.hi + (long long)(sl - ns) * plane + rowoff + m * f.hcomp;
            else
               src = f.base + (long long)sl * plane + rowoff + m * f.comp;
            cp_async8(dst + srow + lane, src + gx0);
            if (second) cp_async8(dst + srow + lane + 32, src + gx1);
         };
         row_copy(A.phi, 0, s_phi);
         if (p.with_T) row_copy(A.T, 0, s_T);
#pragma unroll
         for (int m = 0; m < Q; m++) row_copy(A.q, m, s_q + m * S);
         if (CONC == AMPE_CONC_KKS) row_copy(A.conc, 0, s_c);
         if (CONC != 0) {
            cp_async8(s_cl + srow + lane, A.cl + og + gx0);
            cp_async8(s_ca + srow + lane, A.ca + og + gx0);
            if (second) {
               cp_async8(s_cl + srow + lane + 32, A.cl + og + gx1);
               cp_async8(s_ca + srow + lane + 32, A.ca + og + gx1);
            }
         }
         if (SYMM) {
#pragma unroll
            for (int a = 0; a < ND; a++) {
               cp_async4(s_iq + a * S + srow + lane, A.iq[a] + og + gx0);
               if (second) cp_async4(s_iq + a * S + srow + lane + 32, A.iq[a] + og + gx1);
            }
         }
      }
      cp_async_wait_all();
   }
   __syncthreads();

   const int soff[3] = {1, G::SX, G::SX * G::SY};  // staged strides
   const int foff[3] = {1, G::FX, G::FX * G::FY};  // face-box strides

   // symmetric / plain difference of direction a at staged cell c (lower face of c)
   auto qdiff = [&](int a, int c, double* d) {
      const int cm = c - soff[a];
      if constexpr (Q == 0) {
      } else if constexpr (SYMM) {
         double q2[QN], q2p[QN];
#pragma unroll
         for (int m = 0; m < Q; m++) q2[m] = s_q[m * S + cm];
         symm_rotate<Q>(q2, s_iq[a * S + c], q2p, s_qr, s_conj);
#pragma unroll
         for (int m = 0; m < Q; m++) d[m] = s_q[m * S + c] - q2p[m];
      } else {
#pragma unroll
         for (int m = 0; m < Q; m++) d[m] = s_q[m * S + c] - s_q[m * S + cm];
      }
   };

   // ---- (B) one face: lower face of local cell (li,lj,lk) in direction a --------------
   // lag_owner: this work item is the one that refreshes the lagged arrays for the face
   auto face = [&](int a, int li, int lj, int lk, bool lag_owner) {
      const int c = G::sidx(li, lj, lk);
      const int cm = c - soff[a];
      const int f = a * NFB + G::fidx(li, lj, lk);
      const double phi_c = s_phi[c], phi_m = s_phi[cm];

      // global (wrapped) face index for the lagged arrays
      int gi = ox + li, gj = oy + lj;
      const int gk = oz + lk;
      const bool inrange = (gi - (a == 0) < n0) && (gj - (a == 1) < n1) && (gk - (a == 2) < n2);
      gi = (gi >= n0) ? gi % n0 : gi;
      if (ND == 3) gj = (gj >= n1) ? gj % n1 : gj;
      const long long gface = (ND == 3) ? (gi + (long long)n0 * (gj + (long long)n1 * gk))
                                        : (gi + (long long)n0 * gj);
      const bool wr = A.write_lag && lag_owner && inrange;

      // ---- quaternion face coefficient (compute_face_coef) ----
      if constexpr (Q > 0) if (p.evolve_quat) {
         double normi;
         if (A.use_lag) {
            normi = inrange ? A.lagN[a][gface] : 0.0;
         } else {
            double g2 = 0.0;
#pragma unroll
            for (int n = 0; n < ND; n++) {
               if (n == a) {
                  double d[QN];
                  qdiff(a, c, d);
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = p.dinv[a] * d[m];
                     g2 = g2 + g * g;
                  }
               } else {
                  const int ct = c + soff[n], cmt = cm + soff[n];
                  double g[QN];
                  if (SYMM && Q > 1) {
                     double d1[QN], d1p[QN], d2[QN], d2p[QN], d3[QN], d4[QN], d4p[QN], d0[QN];
                     qdiff(n, ct, d1);
                     qdiff(n, cmt, d2);
                     qdiff(n, cm, d3);
                     qdiff(n, c, d0);
                     symm_rotate<Q>(d1, -s_iq[n * S + ct], d1p, s_qr, s_conj);
                     symm_rotate<Q>(d2, -s_iq[n * S + cmt], d2p, s_qr, s_conj);
#pragma unroll
                     for (int m = 0; m < Q; m++) d4[m] = d2p[m] + d3[m];
                     symm_rotate<Q>(d4, s_iq[a * S + c], d4p, s_qr, s_conj);
#pragma unroll
                     for (int m = 0; m < Q; m++) g[m] = p.p25inv[n] * (d4p[m] + d1p[m] + d0[m]);
                  } else {
#pragma unroll
                     for (int m = 0; m < Q; m++) {
                        const double* qm = s_q + m * S;
                        // diff_t(cm+e_t) + diff_t(cm) + diff_t(c+e_t) + diff_t(c)
                        g[m] = p.p25inv[n] *
                               ((qm[cmt] - qm[cmt - soff[n]]) + (qm[cm] - qm[cm - soff[n]]) +
                                (qm[ct] - qm[ct - soff[n]]) + (qm[c] - qm[c - soff[n]]));
                     }
                  }
#pragma unroll
                  for (int m = 0; m < Q; m++) g2 = g2 + g[m] * g[m];
               }
            }
            normi = eval_grad_normi(g2, p.grad_floor_type, p.floor2, p.max_normi);
            if (wr) A.lagN[a][gface] = normi;
         }
         const double phia = average_func(phi_m, phi_c, p.avg_func);
         const double tempa = p.with_T ? 0.5 * (s_T[cm] + s_T[c]) : 0.5 * (Tuni + Tuni);
         const double diff = p.misorientation_factor * tempa * interp_func(phia, p.orient_interp1);
         const double hphi2 = interp_func(phia, p.orient_interp2);
         s_fc[f] = -normi * diff - p.epsq2 * hphi2;
      }

      // ---- phase flux, non-simple stencils ----
      if constexpr (Q > 0 && ND == 2) if (p.flux_type == AMPE_FLUX_ANISOTROPIC) {
         // anisotropic_gradient_flux, 2d/quatrhs.m4:154-256
         const int t = 1 - a;
         const double dn = (phi_c - phi_m) * p.dinv[a];
         const double dt = 0.25 *
                           (s_phi[cm + soff[t]] - s_phi[cm - soff[t]] + s_phi[c + soff[t]] -
                            s_phi[c - soff[t]]) *
                           p.dinv[t];
         const double dphidx = (a == 0) ? dn : dt;
         const double dphidy = (a == 0) ? dt : dn;
         double qa = 0.5 * (s_q[cm] + s_q[c]);
         if (qa > 1.0) qa = 1.0;
         if (qa < -1.0) qa = -1.0;
         double sn, cs;
         if (p.knumber == 4 && !p.libm_trig) {
            // cos/sin of 4(theta - psi) without atan/acos/sincos: theta = atan(y/x) enters only
            // through cos 4theta = 1 - 8 x^2 y^2 / r^4, sin 4theta = 4 x y (x^2 - y^2) / r^4, and
            // psi = acos(q) through the Chebyshev polynomials cos 4psi = T4(q),
            // sin 4psi = sqrt((1-q)(1+q)) U3(q).  Absolute error ~1e-16, the same as the libm
            // evaluation of the reference's expression (DESIGN.md "Transcendentals").
            double c4t = 1.0, s4t = 0.0;  // theta = pi/2 branch
            if (fabs(dphidx) > (double)1.e-12f) {
               const double x2 = dphidx * dphidx, y2 = dphidy * dphidy;
               const double r2 = x2 + y2;
               const double inv = 1.0 / (r2 * r2);
               c4t = 1.0 - 8.0 * x2 * y2 * inv;
               s4t = 4.0 * dphidx * dphidy * (x2 - y2) * inv;
            }
            const double q2 = qa * qa;
            double c4p = 8.0 * q2 * (q2 - 1.0) + 1.0;
            double s4p = sqrt((1.0 - qa) * (1.0 + qa)) * (4.0 * qa * (2.0 * q2 - 1.0));
            if (Q == 4) {  // psi = 2 acos(q): one more angle doubling
               const double c8 = 2.0 * c4p * c4p - 1.0;
               s4p = 2.0 * s4p * c4p;
               c4p = c8;
            }
            cs = c4t * c4p + s4t * s4p;
            sn = s4t * c4p - c4t * s4p;
         } else {
            aniso_trig_libm(dphidx, dphidy, qa, p.knumber, Q, &sn, &cs);
         }
         const double epstheta = p.epsilon_phase * (1.0 + p.nu * cs);
         const double depsdtheta = -p.knumber * p.epsilon_phase * p.nu * sn;
         s_pf[f] = (a == 0) ? (epstheta * epstheta * dphidx - epstheta * depsdtheta * dphidy)
                            : (epstheta * epstheta * dphidy + epstheta * depsdtheta * dphidx);
      }
      if (ND == 2 && p.flux_type == AMPE_FLUX_ISOTROPIC) {
         // compute_flux_isotropic, 2d/quatrhs.m4:106-151
         const int t = 1 - a;
         s_pf[f] = p.iso_dinv[a] * ((s_phi[c - soff[t]] - s_phi[cm - soff[t]]) +
                                    (phi_c - phi_m) * 10.0 +
                                    (s_phi[c + soff[t]] - s_phi[cm + soff[t]]));
      }

      // ---- composition flux ----
      if constexpr (CONC == AMPE_CONC_EBS) {
         double Dl, Da;
         if (A.use_lag) {
            Dl = inrange ? A.lagD0[a][gface] : 0.0;
            Da = inrange ? A.lagD1[a][gface] : 0.0;
         } else {
            // MobilityCompositionDiffusionStrategy.cc:296-327 + setPFMDiffOnPatch
            const double c_l = 0.5 * (s_cl[c] + s_cl[cm]);
            const double c_a = 0.5 * (s_ca[c] + s_ca[cm]);
            const double dl = diffusion_mobility(p.ct, 0, c_l) * calphad_d2f(p.ct, c_l, 0);
            const double da = diffusion_mobility(p.ct, 1, c_a) * calphad_d2f(p.ct, c_a, 1);
            const double phia = average_func(phi_c, phi_m, p.conc_avg_func);
            const double hphi = interp_func(phia, p.diffusion_interp);
            Dl = (1. - hphi) * dl;
            Da = hphi * da;
            if (wr) {
               A.lagD0[a][gface] = Dl;
               A.lagD1[a][gface] = Da;
            }
         }
         // add_flux (3d/flux.m4:53-66), liquid then solid (EBSCompositionRHSStrategy.cc:258-292)
         double fl = p.dinv[a] * (Dl * (s_cl[c] - s_cl[cm]));
         fl = fl + p.dinv[a] * (Da * (s_ca[c] - s_ca[cm]));
         s_cf[f] = fl;
      } else if constexpr (CONC == AMPE_CONC_KKS) {
         double D0, Dp;
         if (A.use_lag) {
            D0 = inrange ? A.lagD0[a][gface] : 0.0;
            Dp = inrange ? A.lagD1[a][gface] : 0.0;
         } else {
            // concentration_pfmdiffusion (3d/concentrationdiffusion.m4:55-75), uniform T:
            // D_liquid / D_solid already hold d*exp(-q0/R * 2/(T+T))
            const double vphi = average_func(phi_m, phi_c, p.conc_avg_func);
            const double hphi = interp_func(vphi, p.energy_interp);
            D0 = (1.0 - hphi) * p.D_liquid + hphi * p.D_solid;
            // setDiffCoeffForPhaseOnPatch (KKSCompositionRHSStrategy.cc:298-308)
            const double c_l = 0.5 * (s_cl[c] + s_cl[cm]);
            const double c_a = 0.5 * (s_ca[c] + s_ca[cm]);
            const double hp = deriv_interp_func(average_func(phi_c, phi_m, p.conc_avg_func),
                                                p.energy_interp);
            Dp = D0 * hp * (c_l - c_a);
            if (wr) {
               A.lagD0[a][gface] = D0;
               A.lagD1[a][gface] = Dp;
            }
         }
         // concentrationflux (2d/concentrationrhs.m4:52-76)
         s_cf[f] = p.dinv[a] * (D0 * (s_c[c] - s_c[cm]) + Dp * (phi_c - phi_m));
      }
   };

   const bool need_faces = (Q > 0 && p.evolve_quat) || p.flux_type != AMPE_FLUX_SIMPLE || CONC != 0;
   if (need_faces) {
      // lower faces of the owned cells
#pragma unroll 1
      for (int u = 0; u < CPT; u++) {
         const int r = warp + u * RP;
         const int lj = r % TY, lk = r / TY;
#pragma unroll 1
         for (int a = 0; a < ND; a++) face(a, lane, lj, lk, true);
      }
      // upper boundary faces of the tile; they belong to the neighbouring tile except on the
      // extra plane ns of the slab axis, which this tile refreshes in the lagged arrays
#pragma unroll 1
      for (int e = threadIdx.x; e < G::NE; e += NT) {
         int a, li, lj, lk;
         if (e < G::E0) {
            a = 0, li = TX, lj = e % TY, lk = e / TY;
         } else if (e < G::E0 + G::E1) {
            const int g = e - G::E0;
            a = 1, li = g % TX, lj = TY, lk = g / TX;
         } else {
            const int g = e - G::E0 - G::E1;
            a = 2, li = g % TX, lj = g / TX, lk = TZ;
         }
         const bool top = (a == ND - 1) && (((ND == 2) ? oy + lj : oz + lk) == ns);
         face(a, li, lj, lk, top);
      }
   }
   __syncthreads();

   // ---- (C) cells -----------------------------------------------------------------
#pragma unroll 1
   for (int u = 0; u < CPT; u++) {
      const int r = warp + u * RP;
      const int li = lane, lj = r % TY, lk = r / TY;
      const int gi = ox + li, gj = oy + lj, gk = oz + lk;
      if (gi >= n0 || gj >= n1 || gk >= n2) continue;
      if (ND == 2 && gj >= A.s_end) continue;
      if (ND == 3 && gk >= A.s_end) continue;
      const long long gcell = gi + (long long)n0 * (gj + (long long)n1 * gk);
      const int c = G::sidx(li, lj, lk);
      const int fb = G::fidx(li, lj, lk);
      // lower / upper face per direction in the face box
      const int flo[3] = {fb, NFB + fb, 2 * NFB + fb};
      const int fup[3] = {fb + foff[0], NFB + fb + foff[1], 2 * NFB + fb + foff[2]};
      const double phi = s_phi[c];
      const double temp = p.with_T ? s_T[c] : Tuni;

      double phase_rhs = 0.0;
      if (p.with_phase) {
         // computerhspbg (2d/quatrhs.m4:328-402, 3d:430-512)
         double diff_term;
         if (p.flux_type == AMPE_FLUX_SIMPLE) {
            // gradient_flux inlined: flux = (phi(c) - phi(c-e))*(eps2/h)
            diff_term = ((s_phi[c + 1] - phi) * p.eps2_dinv[0] - (phi - s_phi[c - 1]) * p.eps2_dinv[0]) *
                        p.dinv[0];
            diff_term = diff_term + ((s_phi[c + soff[1]] - phi) * p.eps2_dinv[1] -
                                     (phi - s_phi[c - soff[1]]) * p.eps2_dinv[1]) *
                                        p.dinv[1];
            if (ND == 3)
               diff_term = diff_term + ((s_phi[c + soff[2]] - phi) * p.eps2_dinv[2] -
                                        (phi - s_phi[c - soff[2]]) * p.eps2_dinv[2]) *
                                           p.dinv[2];
         } else {
            diff_term = (s_pf[fup[0]] - s_pf[flo[0]]) * p.dinv[0];
            diff_term = diff_term + (s_pf[fup[1]] - s_pf[flo[1]]) * p.dinv[1];
            if (ND == 3) diff_term = diff_term + (s_pf[fup[2]] - s_pf[flo[2]]) * p.dinv[2];
         }
         double rhs = diff_term;
         rhs = rhs - p.phi_well_scale * deriv_well_func(phi, 'd');
         if constexpr (Q > 0) if (p.evolve_quat) {
            // gradient modulus (quatgrad_cell[_symm] + quatgrad_modulus, or from sides compact)
            double s = 0.0;
            if (p.modulus_from_cells) {
               double gc[ND][QN];
#pragma unroll
               for (int a = 0; a < ND; a++) {
                  double dl[QN], du[QN];
                  qdiff(a, c, dl);
                  qdiff(a, c + soff[a], du);
                  if (SYMM && Q > 1) {
                     double dup[QN];
                     symm_rotate<Q>(du, -s_iq[a * S + c + soff[a]], dup, s_qr, s_conj);
#pragma unroll
                     for (int m = 0; m < Q; m++) gc[a][m] = (dup[m] + dl[m]) * p.p5inv[a];
                  } else {
#pragma unroll
                     for (int m = 0; m < Q; m++) gc[a][m] = (du[m] + dl[m]) * p.p5inv[a];
                  }
               }
#pragma unroll
               for (int m = 0; m < Q; m++) {
                  s = s + gc[0][m] * gc[0][m] + gc[1][m] * gc[1][m];
                  if (ND == 3) s = s + gc[ND - 1][m] * gc[ND - 1][m];
               }
               s = sqrt(s);
            } else {
#pragma unroll
               for (int a = 0; a < ND; a++) {
                  double dl[QN], du[QN];
                  qdiff(a, c, dl);
                  qdiff(a, c + soff[a], du);
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = p.dinv[a] * dl[m];
                     s = s + g * g;
                  }
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = p.dinv[a] * du[m];
                     s = s + g * g;
                  }
               }
               s = sqrt(0.5 * s);
            }
            const double p1p = deriv_interp_func(phi, p.orient_interp1);
            const double p2p = deriv_interp_func(phi, p.orient_interp2);
            rhs = rhs - p.misorientation_factor * temp * p1p * s - p2p * p.epsilonq2_half * s * s;
         }
         // addDrivingForce
         if (p.free_energy == AMPE_FE_BIASWELL) {
            // computerhsbiaswell (2d/quatrhs.m4:834-843)
            const double m = p.bias_coeff * atan(p.bias_gamma * (p.meltingT - temp));
            rhs = rhs + m * phi * (1.0 - phi);
         } else if (CONC == AMPE_CONC_EBS && p.free_energy == AMPE_FE_CALPHAD) {
            // CALPHADFreeEnergyStrategyBinary.cc:321-323, 638-663
            const double c_l = s_cl[c], c_a = s_ca[c];
            double f_l = calphad_f(p.ct, c_l, 0);
            f_l *= p.inv_vm_l;
            double f_a = calphad_f(p.ct, c_a, 1);
            f_a *= p.inv_vm_a;
            double mu = calphad_mu(p.ct, c_a, 1);
            mu *= p.inv_vm_a;
            const double hp = deriv_interp_func(phi, p.energy_interp);
            rhs += hp * ((f_l - f_a) - mu * (c_l - c_a));
         } else if (CONC == AMPE_CONC_KKS && p.free_energy == AMPE_FE_QUADRATIC) {
            // QuadraticFreeEnergyStrategy.cc:242-243, 512-530
            const double c_l = s_cl[c], c_a = s_ca[c];
            double f_l = p.quad_A[0] * (c_l - p.quad_ceq[0]) * (c_l - p.quad_ceq[0]);
            f_l *= p.inv_vm_l;
            double f_a = p.quad_A[1] * (c_a - p.quad_ceq[1]) * (c_a - p.quad_ceq[1]);
            f_a *= p.inv_vm_a;
            const double mu = (2. * p.quad_A[0] * (c_l - p.quad_ceq[0])) * p.inv_vm_l;
            const double hp = deriv_interp_func(phi, p.energy_interp);
            rhs += hp * ((f_l - f_a) - mu * (c_l - c_a));
         }
         phase_rhs = rhs * p.phi_mobility;  // PhaseRHSStrategyWithQ.cc:297
         A.out_phi[gcell] = phase_rhs;
      }

      if constexpr (Q > 0) if (p.evolve_quat) {
         // compute_flux_from_gradq + compute_lambda_flux + add_quat_proj_op
         double divm[QN], qc[QN];
         double dlo[ND][QN], dup_[ND][QN];
#pragma unroll
         for (int a = 0; a < ND; a++) {
            qdiff(a, c, dlo[a]);
            qdiff(a, c + soff[a], dup_[a]);
         }
         double lam = 0.0, sumq2 = 0.0;
#pragma unroll
         for (int m = 0; m < Q; m++) {
            qc[m] = s_q[m * S + c];
            double dv = 0.0, lv = 0.0;
#pragma unroll
            for (int a = 0; a < ND; a++) {
               const double fu = s_fc[fup[a]] * (p.dinv[a] * dup_[a][m]);
               const double fl = s_fc[flo[a]] * (p.dinv[a] * dlo[a][m]);
               dv = (a == 0) ? (fu - fl) * p.dinv[a] : dv + (fu - fl) * p.dinv[a];
               lv = (a == 0) ? (fu - fl) * p.p5inv[a] : lv + (fu - fl) * p.p5inv[a];
            }
            divm[m] = dv;
            lam = lam - qc[m] * lv;
            sumq2 = sumq2 + qc[m] * qc[m];
         }
         lam = lam / sumq2;
         const double mob = quat_mobility(phi, p.quat_mobility_func, p.quat_mobility,
                                          p.min_quat_mobility, p.quat_mobility_alt);
         double rq[QN];
#pragma unroll
         for (int m = 0; m < Q; m++) {
            if (Q != 1)
               rq[m] = 0.0 - mob * (divm[m] + 2.0 * qc[m] * lam);
            else
               rq[m] = 0.0 - mob * divm[m];
         }
         if (SYMM) {
            // correctrhsquatforsymmetry (2d/...m4:73-140): dlo/dup_ are the symmetric diffs
            double tmp[QN];
            double dpr[ND][QN];
#pragma unroll
            for (int a = 0; a < ND; a++) {
               if (Q > 1)
                  symm_rotate<Q>(dup_[a], -s_iq[a * S + c + soff[a]], dpr[a], s_qr, s_conj);
               else
                  dpr[a][0] = dup_[a][0];
            }
#pragma unroll
            for (int m = 0; m < Q; m++) {
               double tt = 0.0;
#pragma unroll
               for (int a = 0; a < ND; a++) {
                  const double nsd_u = s_q[m * S + c + soff[a]] - qc[m];
                  const double nsd_l = qc[m] - s_q[m * S + c - soff[a]];
                  const double term = p.dinv2[a] * (s_fc[fup[a]] * (nsd_u - dpr[a][m]) -
                                                    s_fc[flo[a]] * (nsd_l - dlo[a][m]));
                  tt = (a == 0) ? term : tt + term;
               }
               tmp[m] = tt;
            }
            if (Q > 1) {
               double beta = 0.0, lambda = 0.0;
#pragma unroll
               for (int m = 0; m < Q; m++) {
                  beta = beta + qc[m] * qc[m];
                  lambda = lambda + qc[m] * tmp[m];
               }
               lambda = lambda / beta;
#pragma unroll
               for (int m = 0; m < Q; m++) rq[m] = rq[m] + mob * (tmp[m] - lambda * qc[m]);
            } else {
               rq[0] = rq[0] + mob * tmp[0];
            }
         }
#pragma unroll
         for (int m = 0; m < Q; m++) A.out_q[gcell + m * ncell] = rq[m];
      }

      if (CONC != 0) {
         // computerhsconcentration (3d/concentrationrhs.m4:412-458)
         double s = p.dinv[0] * (s_cf[fup[0]] - s_cf[flo[0]]) + p.dinv[1] * (s_cf[fup[1]] - s_cf[flo[1]]);
         if (ND == 3) s = s + p.dinv[2] * (s_cf[fup[2]] - s_cf[flo[2]]);
         A.out_c[gcell] = p.conc_mobility * s;
      }

      if (p.with_T) {
         // computerhstemp + laplacian (2d/quatrhs.m4:787-803, 2d/laplacian.m4:37-52)
         const double dtx = (s_T[c - 1] - 2.0 * temp + s_T[c + 1]);
         const double dty = (s_T[c - soff[1]] - 2.0 * temp + s_T[c + soff[1]]);
         double dterm = dtx * p.dinv2[0] + dty * p.dinv2[1];
         if (ND == 3) {
            const double dtz = (s_T[c - soff[2]] - 2.0 * temp + s_T[c + soff[2]]);
            dterm = dterm + dtz * p.dinv2[2];
         }
         double r = p.thermal_diffusivity * dterm;
         if (p.with_phase) {
            const double gamma = p.latent_heat / p.cp;
            r = r + gamma * phase_rhs;
         }
         A.out_T[gcell] = r;
      }
   }
}

// ---- KKS pre-pass: (c_l, c_a) per cell on the slab + its ghost planes -----------------
// CALPHADequilibriumPhaseConcentrationsStrategy.cc:162-454 (Newton, warm start from *_ref)
// QuadraticEquilibriumPhaseConcentrationsStrategy.cc:42-144 (closed form, appendix.tex:462-490)
struct KksArgs {
   Params p;
   Field phi, conc;
   const double* cl_ref;  // slab-ghosted
   const double* ca_ref;
   double* cl;            // slab-ghosted outputs
   double* ca;
   double* df;            // ghost-0: CALPHAD driving force per interior cell (may be null)
   int* nfail;
   int s_begin, s_end;    // slab index range incl. ghosts: [-1, ns+1)
};

template <int ND>
__global__ void __launch_bounds__(256) kks_kernel(const __grid_constant__ KksArgs A)
{
   const Params& p = A.p;
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;
   const long long total = plane * (A.s_end - A.s_begin);
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      const int sl = (int)(idx / plane) + A.s_begin;
      const long long inplane = idx % plane;
      double phi, conc;
      if (sl < 0) {
         phi = A.phi.lo[(long long)(sl + 1) * plane + inplane];
         conc = A.conc.lo[(long long)(sl + 1) * plane + inplane];
      } else if (sl >= ns) {
         phi = A.phi.hi[(long long)(sl - ns) * plane + inplane];
         conc = A.conc.hi[(long long)(sl - ns) * plane + inplane];
      } else {
         phi = A.phi.base[(long long)sl * plane + inplane];
         conc = A.conc.base[(long long)sl * plane + inplane];
      }
      const long long og = (long long)(sl + 1) * plane + inplane;
      const double hphi = interp_func(phi, p.conc_interp);
      double x0, x1;
      if (p.free_energy == AMPE_FE_CALPHAD) {
         x0 = A.cl_ref[og];
         x1 = A.ca_ref[og];
         double lg[4];
         const int st = kks_newton(p.ct, conc, hphi, x0, x1, p.newton_tol, p.newton_max_its,
                                   p.newton_alpha, lg);
         if (st < 0) atomicAdd(A.nfail, 1);
         if (A.df && sl >= 0 && sl < ns)
            A.df[(long long)sl * plane + inplane] =
                calphad_driving_force(p.ct, x0, x1, lg, p.inv_vm_l, p.inv_vm_a);
      } else {
         const double h = clamp01(hphi);
         x0 = (conc - h * (p.quad_ceq[1] - p.quad_rla * p.quad_ceq[0])) /
              ((1.0 - h) + h * p.quad_rla);
         x1 = (conc - (1.0 - h) * (p.quad_ceq[0] - p.quad_ral * p.quad_ceq[1])) /
              ((1.0 - h) * p.quad_ral + h);
      }
      A.cl[og] = x0;
      A.ca[og] = x1;
   }
}

// ---- Cahn-Hilliard composition RHS (config C1), ghost width 2 ---------------------------
// add_cahnhilliarddoublewell_flux (2d/concentrationrhs.m4:85-137) + computerhsconcentration
struct ChArgs {
   Params p;
   Field conc;  // lo/hi hold 2 planes each
   double* out_c;
   int s_begin, s_end;
};

template <int ND>
__global__ void __launch_bounds__(256) ch_kernel(const __grid_constant__ ChArgs A)
{
   const Params& p = A.p;
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;
   const long long total = plane * (A.s_end - A.s_begin);
   auto at = [&](int i, int j, int k) -> double {
      i = (i < 0) ? i + n0 : ((i >= n0) ? i - n0 : i);
      int sl;
      long long inplane;
      if (ND == 3) {
         j = (j < 0) ? j + n1 : ((j >= n1) ? j - n1 : j);
         sl = k;
         inplane = i + (long long)n0 * j;
      } else {
         sl = j;
         inplane = i;
      }
      if (sl < 0) return A.conc.lo[(long long)(sl + 2) * plane + inplane];
      if (sl >= ns) return A.conc.hi[(long long)(sl - ns) * plane + inplane];
      return A.conc.base[(long long)sl * plane + inplane];
   };
   auto mu = [&](int i, int j, int k) -> double {
      const double c = at(i, j, k);
      double lap = p.ch_dinv2[0] * (-2.0 * c + at(i - 1, j, k) + at(i + 1, j, k)) +
                   p.ch_dinv2[1] * (-2.0 * c + at(i, j - 1, k) + at(i, j + 1, k));
      if (ND == 3) lap = lap + p.ch_dinv2[2] * (-2.0 * c + at(i, j, k - 1) + at(i, j, k + 1));
      return 2.0 * p.ch_well_scale * (c - p.ch_ca) * (p.ch_cb - c) * (p.ch_cb + p.ch_ca - 2.0 * c) -
             p.ch_kappa * lap;
   };
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
        idx += (long long)gridDim.x * blockDim.x) {
      const long long cell = idx + (long long)A.s_begin * plane;
      const int i = (int)(cell % n0);
      const int j = (int)((cell / n0) % n1);
      const int k = (int)(cell / ((long long)n0 * n1));
      const double mc = mu(i, j, k);
      // scatter order of the reference: flux(i) = (0 - M/h mu(i-1)) + M/h mu(i)
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < ND; a++) {
         const double mm = mu(i - (a == 0), j - (a == 1), k - (a == 2));
         const double mp = mu(i + (a == 0), j + (a == 1), k + (a == 2));
         const double flo = (0.0 - p.ch_mdinv[a] * mm) + p.ch_mdinv[a] * mc;
         const double fup = (0.0 - p.ch_mdinv[a] * mc) + p.ch_mdinv[a] * mp;
         s = (a == 0) ? p.dinv[a] * (fup - flo) : s + p.dinv[a] * (fup - flo);
      }
      A.out_c[cell] = p.conc_mobility * s;
   }
}

}  // namespace ampe
