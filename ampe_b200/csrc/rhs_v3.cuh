// Fused evaluateRHSFunction kernel, generation 3.
//
// One launch computes the phase, orientation, composition and temperature right-hand
// sides of a tile of cells from a shared-memory stage of the state fields (1-cell halo
// incl. edges/corners).  It replaces, per evaluation, the reference's unfused sweeps
// (SURVEY.md 3.2):
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
// exactly once -- each thread computes the lower faces of the cells it owns, the tile's
// upper boundary faces are one extra pass of two warps -- anisotropic phase flux,
// quaternion face coefficient, composition flux -> smem; (C) every CELL: divergences +
// pointwise terms -> global.
//
// What generation 3 changes against generation 2 (profiles/README.md): the kernel was
// issue-bound (1350-4650 thread-instructions per cell, only 30 % of them fp64).
//  * the face direction is a template parameter: every shared-memory offset is an
//    immediate (gen 2 looped over a runtime direction: 10 % IMAD + 5 % LEA);
//  * the model selectors (interpolation / averaging / floor / mobility characters, flux
//    type, free energy) are compile-time constants for the parameter sets of the shipped
//    decks (SelFixed), with a runtime-selector instantiation (SelRuntime) for every
//    other combination -- same code, the selector chains fold away (7 % BRA);
//  * transverse side gradients use one central difference per cell,
//    q(x+e)-q(x-e), instead of the four one-sided differences of quatgrad_side
//    (identical in exact arithmetic; 4 loads + 3 adds per component instead of 8 + 7);
//  * 1/sqrt, 1/x through straight-line Newton (fastmath.cuh); one reciprocal in the
//    harmonic average; the CALPHAD face diffusivity M(c) c(1-c) d2f/dc2 without division;
//  * explicit fma() on the polynomial / dot-product chains (library is built with
//    --fmad=false so that rounding is fixed by the source, not by the optimiser).
// All of it stays within the parity bar (1e-12 with the floor metric, tests/parity.py).
#pragma once
#include "calphad.cuh"
#include "fastmath.cuh"
#include "params.h"
#include "pointwise.cuh"
#include "rhs_fused.cuh"  // Field, FusedArgs, symm_rotate, cp_async helpers

namespace ampe {

// ---- selector policies -----------------------------------------------------------------
struct SelRuntime {
   static constexpr bool fixed = false;
   static constexpr int with_phase = 1, evolve_quat = 1, flux_type = 0, free_energy = 0;
   static constexpr int modulus_from_cells = 1, knumber = 4, libm_trig = 0;
   static constexpr char energy_interp = 0, diffusion_interp = 0, orient_interp1 = 0,
                         orient_interp2 = 0, avg_func = 0, conc_avg_func = 0,
                         grad_floor_type = 0, quat_mobility_func = 0;
};
// QuatModelParameters defaults (orient 'q'/'c', floor 'm', mobility 'p', diffusion 'l',
// modulus from cells) + the deck's own choices
template <int FLUX, int FE, char EI, char AVG, char CAVG>
struct SelFixed {
   static constexpr bool fixed = true;
   static constexpr int with_phase = 1, evolve_quat = 1, flux_type = FLUX, free_energy = FE;
   static constexpr int modulus_from_cells = 1, knumber = 4, libm_trig = 0;
   static constexpr char energy_interp = EI, diffusion_interp = 'l', orient_interp1 = 'q',
                         orient_interp2 = 'c', avg_func = AVG, conc_avg_func = CAVG,
                         grad_floor_type = 'm', quat_mobility_func = 'p';
};
using SelDendrite = SelFixed<AMPE_FLUX_ANISOTROPIC, AMPE_FE_BIASWELL, 'p', 'h', 'h'>;  // examples/Dendrite2D
using SelAuNi = SelFixed<AMPE_FLUX_SIMPLE, AMPE_FE_CALPHAD, 'p', 'a', 'a'>;            // examples/AuNi_{2D,3D}
using SelHBSM = SelFixed<AMPE_FLUX_SIMPLE, AMPE_FE_QUADRATIC, 'h', 'a', 'a'>;          // tests/TwoGrainsQuadratic

#define AMPE_SEL(name) (SEL::fixed ? SEL::name : p.name)

// ---- pointwise functions, generation-3 forms ---------------------------------------------
// average_func (functions.f:333-367): harmonic 2/(1/a+1/b) evaluated as 2ab/(a+b)
AMPE_DEV double average3(double a, double b, char type)
{
   if (type == 'a') return 0.5 * (a + b);
   const double r = (2.0 * a * b) * rcp_fast(a + b);
   return (a < 1.0e-16 || b < 1.0e-16) ? 0.0 : r;
}
// eval_grad_normi (quat.f:1497-1537), 'm': 1/sqrt through rsqrt
AMPE_DEV double grad_normi3(double g2, char floor_type, double floor2, double max_normi)
{
   if (floor_type == 'm') {
      const double r = rsqrt_fast(fmax(g2, floor2));
      return (g2 > floor2) ? r : max_normi;
   }
   return eval_grad_normi_rare(g2, floor_type, floor2, max_normi);
}
// interp_func 'p' / deriv with explicit fma
AMPE_DEV double interp3(double phi, char type)
{
   if (type == 'p') {
      const double t = clamp01(phi);
      return t * t * t * fma(t, fma(6.0, t, -15.0), 10.0);
   }
   return interp_func(phi, type);
}

// quatmobility 'p' (3d/mobility.m4:42-92) with explicit fma
AMPE_DEV double quat_mobility3(double phi, char func, double scale, double minm, double alt)
{
   if (func == 'p' || func == 'P') {
      const double t = clamp01(phi);
      const double qfunc = 1.0 - t * t * t * fma(t, fma(6.0, t, -15.0), 10.0);
      return fma(scale - minm, qfunc, minm);
   }
   return quat_mobility_rare(phi, func, scale, minm, alt);
}

// CALPHAD face diffusivity of one phase at face-averaged concentration c0:
//   D = [c0 c1 (c0 M1 + c1 M0) 1e12] * d2f/dc2,  d2f = fmix'' + RT (1/c0 + 1/c1)
// (computeDiffusionMobilityBinaryPhase, CALPHADMobility.cc:200-219, times
//  computeSecondDerivativeFreeEnergy; MobilityCompositionDiffusionStrategy.cc:296-327).
// With c0, c1 > 1e-8 (always, away from the xlogx extension) c0 c1 d2f = c0 c1 fmix'' + RT,
// which needs no division.
AMPE_DEV double ebs_phase_diffusivity(const CalphadT& t, int ph, double c0)
{
   const double c1 = 1. - c0;
   const double dc = c0 - c1;
   const double cc = c0 * c1;
   double m[2];
#pragma unroll
   for (int sp = 0; sp < 2; sp++) {
      const double* qq = t.qAB[sp][ph];
      const double poly = fma(dc, fma(dc, fma(dc, qq[3], qq[2]), qq[1]), qq[0]);
      const double dG = fma(cc, poly, fma(c0, t.qA[sp][ph], c1 * t.qB[sp][ph]));
      m[sp] = exp(dG * t.RTinv);
   }
   const double mm = fma(c0, m[1], c1 * m[0]) * t.RTinv;
   if (c0 > AMPE_SMALLX && c1 > AMPE_SMALLX) {
      const double* L = t.L[ph];
      const double tt = 2.0 * c0 - 1.0;
      const double f0 = fma(tt, fma(tt, fma(tt, L[3], L[2]), L[1]), L[0]);
      const double f1 = fma(tt, fma(tt, 6.0 * L[3], 4.0 * L[2]), 2.0 * L[1]);
      const double f2 = fma(24.0 * L[3], tt, 8.0 * L[2]);
      const double fm2 = fma(cc, f2, fma(2.0 * (1.0 - 2.0 * c0), f1, -2.0 * f0));
      return (mm * 1.e12) * fma(cc, fm2, t.RT);
   }
   return (cc * mm * 1.e12) * calphad_d2f(t, c0, ph);
}

// ---- tile geometry + shared-memory carve-up -----------------------------------------------
template <int ND_, int Q_, int CONC_, bool SYMM_, bool WT_, class SEL_, int TX_, int TY_, int TZ_, int NT_>
struct Tile3 {
   static constexpr int ND = ND_, Q = Q_, CONC = CONC_, TX = TX_, TY = TY_, TZ = TZ_, NT = NT_;
   static constexpr bool SYMM = SYMM_, WT = WT_;
   using SEL = SEL_;
   static constexpr int HZ = (ND == 3) ? 1 : 0;
   static constexpr int SX = TX + 2, SY = TY + 2, SZ = TZ + 2 * HZ;
   static constexpr int S = SX * SY * SZ;
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
   // staged strides / face-box strides per direction
   __host__ __device__ static constexpr int str(int a) { return a == 0 ? 1 : (a == 1 ? SX : SX * SY); }
   __host__ __device__ static constexpr int ftr(int a) { return a == 0 ? 1 : (a == 1 ? FX : FX * FY); }
   AMPE_DEV static int sidx(int i, int j, int k) { return (i + 1) + SX * ((j + 1) + SY * (k + HZ)); }
   AMPE_DEV static int fidx(int i, int j, int k) { return i + FX * (j + FY * k); }
};

template <class TT>
struct Rhs3 {
   using SEL = typename TT::SEL;
   static constexpr int ND = TT::ND, Q = TT::Q, CONC = TT::CONC, S = TT::S, NFB = TT::NFB;
   static constexpr bool SYMM = TT::SYMM, WT = TT::WT;
   static constexpr int QN = (Q > 0) ? Q : 1;

   // symmetric / plain difference of direction a at staged cell c (lower face of c)
   template <int a>
   AMPE_DEV static void qdiff(const double* s, const int* s_iq, const double (*s_qr)[4],
                              const int* s_conj, int c, double* d)
   {
      constexpr int sa = TT::str(a);
      const double* sq = s + TT::O_Q;
      if constexpr (Q == 0) {
      } else if constexpr (SYMM) {
         double q2[QN], q2p[QN];
#pragma unroll
         for (int m = 0; m < Q; m++) q2[m] = sq[m * S + c - sa];
         symm_rotate<Q>(q2, s_iq[a * S + c], q2p, s_qr, s_conj);
#pragma unroll
         for (int m = 0; m < Q; m++) d[m] = sq[m * S + c] - q2p[m];
      } else {
#pragma unroll
         for (int m = 0; m < Q; m++) d[m] = sq[m * S + c] - sq[m * S + c - sa];
      }
   }

   // |grad q|^2 on the lower face of c in direction a (quatgrad_side[_symm] + quatgrad_modulus
   // of the side gradient, 3d/quatgrad.m4:376-720)
   template <int a>
   AMPE_DEV static double face_grad2(const double* s, const int* s_iq, const double (*s_qr)[4],
                                     const int* s_conj, const Params& p, int c)
   {
      constexpr int sa = TT::str(a);
      const int cm = c - sa;
      const double* sq = s + TT::O_Q;
      double g2 = 0.0;
#pragma unroll
      for (int n = 0; n < ND; n++) {
         const int sn = TT::str(n);
         if (n == a) {
            double d[QN];
            qdiff<a>(s, s_iq, s_qr, s_conj, c, d);
#pragma unroll
            for (int m = 0; m < Q; m++) {
               const double g = p.dinv[a] * d[m];
               g2 = fma(g, g, g2);
            }
         } else if (SYMM && Q > 1) {
            // quatgrad_side_symm: the four one-sided differences are rotated into the
            // frame of the face before they are averaged
            double g[QN];
            const int ct = c + sn, cmt = cm + sn;
            double d1[QN], d1p[QN], d2[QN], d2p[QN], d3[QN], d4[QN], d4p[QN], d0[QN];
            if (n == 0) {
               qdiff<0>(s, s_iq, s_qr, s_conj, ct, d1);
               qdiff<0>(s, s_iq, s_qr, s_conj, cmt, d2);
               qdiff<0>(s, s_iq, s_qr, s_conj, cm, d3);
               qdiff<0>(s, s_iq, s_qr, s_conj, c, d0);
            } else if (n == 1) {
               qdiff<1>(s, s_iq, s_qr, s_conj, ct, d1);
               qdiff<1>(s, s_iq, s_qr, s_conj, cmt, d2);
               qdiff<1>(s, s_iq, s_qr, s_conj, cm, d3);
               qdiff<1>(s, s_iq, s_qr, s_conj, c, d0);
            } else {
               qdiff<(ND == 3 ? 2 : 0)>(s, s_iq, s_qr, s_conj, ct, d1);
               qdiff<(ND == 3 ? 2 : 0)>(s, s_iq, s_qr, s_conj, cmt, d2);
               qdiff<(ND == 3 ? 2 : 0)>(s, s_iq, s_qr, s_conj, cm, d3);
               qdiff<(ND == 3 ? 2 : 0)>(s, s_iq, s_qr, s_conj, c, d0);
            }
            symm_rotate<Q>(d1, -s_iq[n * S + ct], d1p, s_qr, s_conj);
            symm_rotate<Q>(d2, -s_iq[n * S + cmt], d2p, s_qr, s_conj);
#pragma unroll
            for (int m = 0; m < Q; m++) d4[m] = d2p[m] + d3[m];
            symm_rotate<Q>(d4, s_iq[a * S + c], d4p, s_qr, s_conj);
#pragma unroll
            for (int m = 0; m < Q; m++) g[m] = p.p25inv[n] * (d4p[m] + d1p[m] + d0[m]);
#pragma unroll
            for (int m = 0; m < Q; m++) g2 = g2 + g[m] * g[m];
         } else {
#pragma unroll
            for (int m = 0; m < Q; m++) {
               const double* qm = sq + m * S;
               // central differences of the two cells sharing the face
               const double g = p.p25inv[n] * ((qm[cm + sn] - qm[cm - sn]) + (qm[c + sn] - qm[c - sn]));
               g2 = fma(g, g, g2);
            }
         }
      }
      return g2;
   }

   // ---- (B) one face: lower face of the staged cell c in direction a ------------------------
   //  f      index in the face box;  gface  index in the lagged arrays;  inrange  the face
   //  belongs to a cell of the domain;  wr  this work item refreshes the lagged arrays
   template <int a>
   AMPE_DEV static void face(const FusedArgs& A, double* s, const int* s_iq, const double (*s_qr)[4],
                             const int* s_conj, int c, int f, long long gface, bool inrange, bool wr)
   {
      const Params& p = A.p;
      constexpr int sa = TT::str(a);
      const int cm = c - sa;
      const double phi_c = s[TT::O_PHI + c], phi_m = s[TT::O_PHI + cm];
      const bool evolve_quat = (Q > 0) && AMPE_SEL(evolve_quat);
      const int flux_type = AMPE_SEL(flux_type);

      // ---- quaternion face coefficient (compute_face_coef, quatfacops.m4) ----
      if constexpr (Q > 0) if (evolve_quat) {
         double normi;
         if (A.use_lag) {
            normi = inrange ? A.lagN[a][gface] : 0.0;
         } else {
            const double g2 = face_grad2<a>(s, s_iq, s_qr, s_conj, p, c);
            normi = grad_normi3(g2, AMPE_SEL(grad_floor_type), p.floor2, p.max_normi);
            if (wr) A.lagN[a][gface] = normi;
         }
         const double phia = average3(phi_m, phi_c, AMPE_SEL(avg_func));
         const double tempa = WT ? 0.5 * (s[TT::O_T + cm] + s[TT::O_T + c]) : 0.5 * (p.T_uniform + p.T_uniform);
         const double diff = p.misorientation_factor * tempa * interp3(phia, AMPE_SEL(orient_interp1));
         const double hphi2 = interp3(phia, AMPE_SEL(orient_interp2));
         s[TT::O_FC + a * NFB + f] = -normi * diff - p.epsq2 * hphi2;
      }

      // ---- phase flux, non-simple stencils (2D) ----
      if constexpr (TT::HAS_PF && Q > 0) if (flux_type == AMPE_FLUX_ANISOTROPIC) {
         // anisotropic_gradient_flux, 2d/quatrhs.m4:154-256
         constexpr int st = TT::str(1 - a);
         const double* sp = s + TT::O_PHI;
         const double dn = (phi_c - phi_m) * p.dinv[a];
         const double dt = 0.25 * (sp[cm + st] - sp[cm - st] + sp[c + st] - sp[c - st]) * p.dinv[1 - a];
         const double dphidx = (a == 0) ? dn : dt;
         const double dphidy = (a == 0) ? dt : dn;
         double qa = 0.5 * (s[TT::O_Q + cm] + s[TT::O_Q + c]);
         qa = fmin(1.0, fmax(-1.0, qa));
         double sn, cs;
         if (AMPE_SEL(knumber) == 4 && !AMPE_SEL(libm_trig)) {
            // cos/sin of 4(theta - psi) without atan/acos/sincos (DESIGN.md "Transcendentals")
            double c4t = 1.0, s4t = 0.0;  // theta = pi/2 branch
            const double x2 = dphidx * dphidx, y2 = dphidy * dphidy;
            const double r2 = x2 + y2;
            if (fabs(dphidx) > (double)1.e-12f) {
               const double inv = rcp_fast(r2 * r2);
               c4t = fma(-8.0 * x2 * y2, inv, 1.0);
               s4t = 4.0 * dphidx * dphidy * (x2 - y2) * inv;
            }
            const double q2 = qa * qa;
            double c4p = fma(8.0 * q2, q2 - 1.0, 1.0);
            double s4p = sqrt_fast((1.0 - qa) * (1.0 + qa)) * (4.0 * qa * fma(2.0, q2, -1.0));
            if (Q == 4) {  // psi = 2 acos(q): one more angle doubling
               const double c8 = fma(2.0 * c4p, c4p, -1.0);
               s4p = 2.0 * s4p * c4p;
               c4p = c8;
            }
            cs = fma(c4t, c4p, s4t * s4p);
            sn = fma(s4t, c4p, -(c4t * s4p));
         } else {
            aniso_trig_libm(dphidx, dphidy, qa, p.knumber, Q, &sn, &cs);
         }
         const double epstheta = p.epsilon_phase * fma(p.nu, cs, 1.0);
         const double depsdtheta = -p.knumber * p.epsilon_phase * p.nu * sn;
         const double e2 = epstheta * epstheta, ed = epstheta * depsdtheta;
         s[TT::O_PF + a * NFB + f] = (a == 0) ? fma(e2, dphidx, -(ed * dphidy)) : fma(e2, dphidy, ed * dphidx);
      }
      if constexpr (TT::HAS_PF) if (flux_type == AMPE_FLUX_ISOTROPIC) {
         // compute_flux_isotropic, 2d/quatrhs.m4:106-151
         constexpr int st = TT::str(1 - a);
         const double* sp = s + TT::O_PHI;
         s[TT::O_PF + a * NFB + f] = p.iso_dinv[a] * ((sp[c - st] - sp[cm - st]) + (phi_c - phi_m) * 10.0 +
                                                  (sp[c + st] - sp[cm + st]));
      }

      // ---- composition flux ----
      if constexpr (CONC == AMPE_CONC_EBS) {
         const double* scl = s + TT::O_CL;
         const double* sca = s + TT::O_CA;
         double Dl, Da;
         if (A.use_lag) {
            Dl = inrange ? A.lagD0[a][gface] : 0.0;
            Da = inrange ? A.lagD1[a][gface] : 0.0;
         } else {
            // MobilityCompositionDiffusionStrategy.cc:296-327 + setPFMDiffOnPatch
            const double c_l = 0.5 * (scl[c] + scl[cm]);
            const double c_a = 0.5 * (sca[c] + sca[cm]);
            const double dl = ebs_phase_diffusivity(p.ct, 0, c_l);
            const double da = ebs_phase_diffusivity(p.ct, 1, c_a);
            const double phia = average3(phi_c, phi_m, AMPE_SEL(conc_avg_func));
            const double hphi = interp3(phia, AMPE_SEL(diffusion_interp));
            Dl = (1. - hphi) * dl;
            Da = hphi * da;
            if (wr) {
               A.lagD0[a][gface] = Dl;
               A.lagD1[a][gface] = Da;
            }
         }
         // add_flux (3d/flux.m4:53-66), liquid then solid (EBSCompositionRHSStrategy.cc:258-292)
         double fl = p.dinv[a] * (Dl * (scl[c] - scl[cm]));
         fl = fl + p.dinv[a] * (Da * (sca[c] - sca[cm]));
         s[TT::O_CF + a * NFB + f] = fl;
      } else if constexpr (CONC == AMPE_CONC_KKS) {
         const double* scl = s + TT::O_CL;
         const double* sca = s + TT::O_CA;
         double D0, Dp;
         if (A.use_lag) {
            D0 = inrange ? A.lagD0[a][gface] : 0.0;
            Dp = inrange ? A.lagD1[a][gface] : 0.0;
         } else {
            // concentration_pfmdiffusion (3d/concentrationdiffusion.m4:55-75), uniform T
            const double vphi = average3(phi_m, phi_c, AMPE_SEL(conc_avg_func));
            const double hphi = interp3(vphi, AMPE_SEL(energy_interp));
            D0 = (1.0 - hphi) * p.D_liquid + hphi * p.D_solid;
            // setDiffCoeffForPhaseOnPatch (KKSCompositionRHSStrategy.cc:298-308)
            const double c_l = 0.5 * (scl[c] + scl[cm]);
            const double c_a = 0.5 * (sca[c] + sca[cm]);
            const double hp = deriv_interp_func(average3(phi_c, phi_m, AMPE_SEL(conc_avg_func)),
                                                AMPE_SEL(energy_interp));
            Dp = D0 * hp * (c_l - c_a);
            if (wr) {
               A.lagD0[a][gface] = D0;
               A.lagD1[a][gface] = Dp;
            }
         }
         // concentrationflux (2d/concentrationrhs.m4:52-76)
         s[TT::O_CF + a * NFB + f] = p.dinv[a] * (D0 * (s[TT::O_C + c] - s[TT::O_C + cm]) + Dp * (phi_c - phi_m));
      }
   }

   // ---- (C) one cell ---------------------------------------------------------------------------
   AMPE_DEV static void cell(const FusedArgs& A, const double* s, const int* s_iq, const double (*s_qr)[4],
                             const int* s_conj, int c, int fb, long long gcell, long long ncell)
   {
      const Params& p = A.p;
      const bool evolve_quat = (Q > 0) && AMPE_SEL(evolve_quat);
      const int flux_type = AMPE_SEL(flux_type);
      const int free_energy = AMPE_SEL(free_energy);
      const double phi = s[TT::O_PHI + c];
      const double temp = WT ? s[TT::O_T + c] : p.T_uniform;
      const double* sp = s + TT::O_PHI;
      const double* sq = s + TT::O_Q;

      // quaternion differences on the lower / upper faces (symmetric ones in SYMM mode)
      double dlo[ND][QN], dup[ND][QN];
      if constexpr (Q > 0) if (evolve_quat) {
         qdiff<0>(s, s_iq, s_qr, s_conj, c, dlo[0]);
         qdiff<0>(s, s_iq, s_qr, s_conj, c + TT::str(0), dup[0]);
         qdiff<1>(s, s_iq, s_qr, s_conj, c, dlo[1]);
         qdiff<1>(s, s_iq, s_qr, s_conj, c + TT::str(1), dup[1]);
         if constexpr (ND == 3) {
            qdiff<2>(s, s_iq, s_qr, s_conj, c, dlo[ND - 1]);
            qdiff<2>(s, s_iq, s_qr, s_conj, c + TT::str(2), dup[ND - 1]);
         }
      }

      double phase_rhs = 0.0;
      if (AMPE_SEL(with_phase)) {
         // computerhspbg (2d/quatrhs.m4:328-402, 3d:430-512)
         double diff_term;
         if (!TT::HAS_PF || flux_type == AMPE_FLUX_SIMPLE) {
            // gradient_flux inlined: flux = (phi(c) - phi(c-e))*(eps2/h)
            diff_term = ((sp[c + 1] - phi) * p.eps2_dinv[0] - (phi - sp[c - 1]) * p.eps2_dinv[0]) * p.dinv[0];
            diff_term = diff_term + ((sp[c + TT::str(1)] - phi) * p.eps2_dinv[1] -
                                     (phi - sp[c - TT::str(1)]) * p.eps2_dinv[1]) * p.dinv[1];
            if constexpr (ND == 3)
               diff_term = diff_term + ((sp[c + TT::str(2)] - phi) * p.eps2_dinv[2] -
                                        (phi - sp[c - TT::str(2)]) * p.eps2_dinv[2]) * p.dinv[2];
         } else {
            const double* pf = s + TT::O_PF;
            diff_term = (pf[fb + TT::ftr(0)] - pf[fb]) * p.dinv[0];
            diff_term = diff_term + (pf[NFB + fb + TT::ftr(1)] - pf[NFB + fb]) * p.dinv[1];
         }
         double rhs = diff_term;
         rhs = rhs - p.phi_well_scale * deriv_well_func(phi, 'd');
         if constexpr (Q > 0) if (evolve_quat) {
            // gradient modulus (quatgrad_cell[_symm] + quatgrad_modulus, or from sides compact)
            double sm = 0.0;
            if (AMPE_SEL(modulus_from_cells)) {
#pragma unroll
               for (int a = 0; a < ND; a++) {
                  double du[QN];
                  if (SYMM && Q > 1) {
                     symm_rotate<Q>(dup[a], -s_iq[a * S + c + TT::str(a)], du, s_qr, s_conj);
                  } else {
#pragma unroll
                     for (int m = 0; m < Q; m++) du[m] = dup[a][m];
                  }
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = (du[m] + dlo[a][m]) * p.p5inv[a];
                     sm = fma(g, g, sm);
                  }
               }
               sm = sqrt_fast(sm);
            } else {
#pragma unroll
               for (int a = 0; a < ND; a++) {
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = p.dinv[a] * dlo[a][m];
                     sm = fma(g, g, sm);
                  }
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = p.dinv[a] * dup[a][m];
                     sm = fma(g, g, sm);
                  }
               }
               sm = sqrt_fast(0.5 * sm);
            }
            const double p1p = deriv_interp_func(phi, AMPE_SEL(orient_interp1));
            rhs = rhs - p.misorientation_factor * temp * p1p * sm;
            if (AMPE_SEL(orient_interp2) != 'c') {  // p2 constant: its derivative term vanishes
               const double p2p = deriv_interp_func(phi, AMPE_SEL(orient_interp2));
               rhs = rhs - p2p * p.epsilonq2_half * sm * sm;
            }
         }
         // addDrivingForce
         if (free_energy == AMPE_FE_BIASWELL) {
            // computerhsbiaswell (2d/quatrhs.m4:834-843)
            const double m = p.bias_coeff * atan(p.bias_gamma * (p.meltingT - temp));
            rhs = rhs + m * phi * (1.0 - phi);
         } else if (CONC == AMPE_CONC_EBS && free_energy == AMPE_FE_CALPHAD) {
            // CALPHADFreeEnergyStrategyBinary.cc:321-323, 638-663: (f_l-f_a) - mu (c_l-c_a) comes
            // from the KKS kernel, which has the logarithms of the converged c_l, c_a at hand
            const double hp = deriv_interp_func(phi, AMPE_SEL(energy_interp));
            rhs += hp * A.df[gcell];
         } else if (CONC == AMPE_CONC_KKS && free_energy == AMPE_FE_QUADRATIC) {
            // QuadraticFreeEnergyStrategy.cc:242-243, 512-530
            const double c_l = s[TT::O_CL + c], c_a = s[TT::O_CA + c];
            double f_l = p.quad_A[0] * (c_l - p.quad_ceq[0]) * (c_l - p.quad_ceq[0]);
            f_l *= p.inv_vm_l;
            double f_a = p.quad_A[1] * (c_a - p.quad_ceq[1]) * (c_a - p.quad_ceq[1]);
            f_a *= p.inv_vm_a;
            const double mu = (2. * p.quad_A[0] * (c_l - p.quad_ceq[0])) * p.inv_vm_l;
            const double hp = deriv_interp_func(phi, AMPE_SEL(energy_interp));
            rhs += hp * ((f_l - f_a) - mu * (c_l - c_a));
         }
         phase_rhs = rhs * p.phi_mobility;  // PhaseRHSStrategyWithQ.cc:297
         A.out_phi[gcell] = phase_rhs;
      }

      if constexpr (Q > 0) if (evolve_quat) {
         // compute_flux_from_gradq + compute_lambda_flux + add_quat_proj_op
         const double* fc = s + TT::O_FC;
         double fcl[ND], fcu[ND];
#pragma unroll
         for (int a = 0; a < ND; a++) {
            fcl[a] = fc[a * NFB + fb];
            fcu[a] = fc[a * NFB + fb + TT::ftr(a)];
         }
         // div(fc grad q) per component in the reference's operation order: the projection
         // below (and the symmetry correction) cancel most of it, so parity needs its exact
         // rounding.  compute_lambda_flux sums the same face differences scaled by 0.5/h, which
         // is exactly half of 1/h: lambda = -(q.div)/(2|q|^2) bit for bit, and
         // 2 q lambda = -q (q.div)/|q|^2 needs no second accumulation.
         double divm[QN], qc[QN];
         double qdiv = 0.0, sumq2 = 0.0;
#pragma unroll
         for (int m = 0; m < Q; m++) {
            qc[m] = sq[m * S + c];
            double dv = 0.0;
#pragma unroll
            for (int a = 0; a < ND; a++) {
               const double fu = fcu[a] * (p.dinv[a] * dup[a][m]);
               const double fl = fcl[a] * (p.dinv[a] * dlo[a][m]);
               dv = (a == 0) ? (fu - fl) * p.dinv[a] : dv + (fu - fl) * p.dinv[a];
            }
            divm[m] = dv;
            qdiv = qdiv + qc[m] * dv;
            sumq2 = sumq2 + qc[m] * qc[m];
         }
         const double lamq = qdiv / sumq2;
         const double mob = quat_mobility(phi, AMPE_SEL(quat_mobility_func), p.quat_mobility,
                                          p.min_quat_mobility, p.quat_mobility_alt);
         double rq[QN];
#pragma unroll
         for (int m = 0; m < Q; m++) {
            if (Q != 1)
               rq[m] = 0.0 - mob * (divm[m] - qc[m] * lamq);
            else
               rq[m] = 0.0 - mob * divm[m];
         }
         if constexpr (SYMM) {
            // correctrhsquatforsymmetry (2d/...m4:73-140): dlo/dup are the symmetric diffs
            double tmp[QN];
            double dpr[ND][QN];
#pragma unroll
            for (int a = 0; a < ND; a++) {
               if (Q > 1)
                  symm_rotate<Q>(dup[a], -s_iq[a * S + c + TT::str(a)], dpr[a], s_qr, s_conj);
               else
                  dpr[a][0] = dup[a][0];
            }
#pragma unroll
            for (int m = 0; m < Q; m++) {
               double tt = 0.0;
#pragma unroll
               for (int a = 0; a < ND; a++) {
                  const double nsd_u = sq[m * S + c + TT::str(a)] - qc[m];
                  const double nsd_l = qc[m] - sq[m * S + c - TT::str(a)];
                  const double term = p.dinv2[a] * (fcu[a] * (nsd_u - dpr[a][m]) - fcl[a] * (nsd_l - dlo[a][m]));
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

      if constexpr (CONC != 0) {
         // computerhsconcentration (3d/concentrationrhs.m4:412-458)
         const double* cf = s + TT::O_CF;
         double sm = p.dinv[0] * (cf[fb + TT::ftr(0)] - cf[fb]) +
                     p.dinv[1] * (cf[NFB + fb + TT::ftr(1)] - cf[NFB + fb]);
         if constexpr (ND == 3) sm = sm + p.dinv[2] * (cf[2 * NFB + fb + TT::ftr(2)] - cf[2 * NFB + fb]);
         A.out_c[gcell] = p.conc_mobility * sm;
      }

      if constexpr (WT) {
         // computerhstemp + laplacian (2d/quatrhs.m4:787-803, 2d/laplacian.m4:37-52)
         const double* sT = s + TT::O_T;
         const double dtx = (sT[c - 1] - 2.0 * temp + sT[c + 1]);
         const double dty = (sT[c - TT::str(1)] - 2.0 * temp + sT[c + TT::str(1)]);
         double dterm = dtx * p.dinv2[0] + dty * p.dinv2[1];
         if constexpr (ND == 3) {
            const double dtz = (sT[c - TT::str(2)] - 2.0 * temp + sT[c + TT::str(2)]);
            dterm = dterm + dtz * p.dinv2[2];
         }
         double r = p.thermal_diffusivity * dterm;
         if (AMPE_SEL(with_phase)) {
            r = fma(p.latent_over_cp, phase_rhs, r);
         }
         A.out_T[gcell] = r;
      }
   }
};

template <class TT>
__global__ void __launch_bounds__(TT::NT) rhs_fused3_kernel(const __grid_constant__ FusedArgs A)
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
   const int n0 = p.n[0], n1 = p.n[1], n2 = (ND == 3) ? p.n[2] : 1;
   const int ns = (ND == 3) ? n2 : n1;  // planes along the slab axis
   const int ox = blockIdx.x * TX;
   const int oy = blockIdx.y * TY + ((ND == 2) ? A.s_begin : 0);
   const int oz = (ND == 3) ? (blockIdx.z * TZ + A.s_begin) : 0;
   const long long plane = (ND == 3) ? (long long)n0 * n1 : (long long)n0;  // slab plane size
   const long long ncell = (long long)n0 * n1 * n2;
   const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;

   // ---- (A) stage: cp.async straight into shared memory (no register staging) ------------
   // A staged row has SX = 34 elements: one warp copies x = 0..31 of a row per instruction,
   // the two tail elements of all rows are gathered into one extra pass of the block.
   {
      constexpr int NROWS = TT::SY * TT::SZ;
      // element (row r, staged x) of every staged field
      auto copy_elem = [&](int r, int xs) {
         int gx = (ox - 1 + xs) % n0;
         gx = (gx < 0) ? gx + n0 : gx;
         const int lj = r % TT::SY - 1;
         const int lk = (ND == 3) ? (r / TT::SY - 1) : 0;
         int sl;
         int inplane = gx;  // offset inside a slab plane
         if (ND == 3) {
            int gj = (oy + lj) % n1;
            gj = (gj < 0) ? gj + n1 : gj;
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
            const int slw = (sl < 0) ? sl + ns : ((sl >= ns) ? sl - ns : sl);
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
#pragma unroll 1
      for (int r = warp; r < NROWS; r += NW) copy_elem(r, lane);
#pragma unroll 1
      for (int t = threadIdx.x; t < 2 * NROWS; t += NT) copy_elem(t >> 1, 32 + (t & 1));
      cp_async_wait_all();
   }
   __syncthreads();

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
#pragma unroll 1
         for (int u = 0; u < CPT; u++) {
            const bool in_i = gi < n0, in_j = gj < n1, in_k = gk < n2;
            {
               const bool inr = (gi <= n0) && in_j && in_k;
               R::template face<0>(A, s, s_iq, s_qr, s_conj, c, f, gcell - ((gi == n0) ? n0 : 0), inr,
                                   A.write_lag && inr);
            }
            {
               const bool inr = in_i && (gj <= n1) && in_k;
               R::template face<1>(A, s, s_iq, s_qr, s_conj, c, f,
                                   gcell - ((ND == 3 && gj == n1) ? plane : 0), inr, A.write_lag && inr);
            }
            if constexpr (ND == 3) {
               const bool inr = in_i && in_j && (gk <= n2);
               R::template face<(ND == 3 ? 2 : 0)>(A, s, s_iq, s_qr, s_conj, c, f, gcell, inr, A.write_lag && inr);
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
         R::template face<0>(A, s, s_iq, s_qr, s_conj, TT::sidx(TX, lj, lk), TT::fidx(TX, lj, lk), gface, inr, false);
      }
      // y = TY: TX*TZ faces, on the last warps so that they run beside the x pass
#pragma unroll 1
      for (int e = NT - 1 - threadIdx.x; e < TX * TZ; e += NT) {
         const int li = e % TX, lk = e / TX;
         bool inr;
         long long gface;
         face_meta(1, ox + li, oy + TY, oz + lk, inr, gface);
         const bool top = (ND == 2) && (oy + TY == ns);
         R::template face<1>(A, s, s_iq, s_qr, s_conj, TT::sidx(li, TY, lk), TT::fidx(li, TY, lk), gface, inr,
                             A.write_lag && top && inr);
      }
      if constexpr (ND == 3) {
#pragma unroll 1
         for (int e = threadIdx.x; e < TX * TY; e += NT) {
            const int li = e % TX, lj = e / TX;
            bool inr;
            long long gface;
            face_meta(2, ox + li, oy + lj, oz + TZ, inr, gface);
            const bool top = (oz + TZ == ns);
            R::template face<(ND == 3 ? 2 : 0)>(A, s, s_iq, s_qr, s_conj, TT::sidx(li, lj, TZ), TT::fidx(li, lj, TZ),
                                                gface, inr, A.write_lag && top && inr);
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
#pragma unroll 1
      for (int u = 0; u < CPT; u++) {
         bool ok = (gi < n0) && (gj < n1) && (gk < n2);
         if (ND == 2) ok = ok && (gj < A.s_end);
         if (ND == 3) ok = ok && (gk < A.s_end);
         if (ok) {
            const long long gcell = gi + (long long)n0 * (gj + (long long)n1 * gk);
            R::cell(A, s, s_iq, s_qr, s_conj, c, fb, gcell, ncell);
         }
         c += CSTEP;
         fb += FSTEP;
         gj += RSTEP_J;
         gk += RSTEP_K;
      }
   }
}

}  // namespace ampe
