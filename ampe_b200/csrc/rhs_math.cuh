// Per-face and per-cell arithmetic of the fused evaluateRHSFunction kernels, independent
// of how the state is staged in shared memory (tile kernel: rhs_tile.cuh; plane-marching
// 3D kernel: rhs_march.cuh).
//
// Reference routines restated here:
//   quatdiffs, quatgrad_cell/side, quatgrad_modulus        {2d,3d}/quatdiffs.m4, quatgrad.m4
//   quatmobility                                           {2d,3d}/mobility.m4
//   gradient_flux / compute_flux_isotropic / anisotropic_gradient_flux, computerhspbg,
//   computerhsbiaswell, computerhstemp                     {2d,3d}/quatrhs.m4
//   compute_face_coef, compute_flux_from_gradq, compute_lambda_flux, add_quat_proj_op
//                                                          {2d,3d}/quatfacops.m4
//   correctrhsquatforsymmetry                              {2d,3d}/correctrhsquatforsymmetry.m4
//   Quadratic free energy + driving force, EBS face diffusivities (CALPHADMobility),
//   add_flux / concentrationflux / concentration_pfmdiffusion, computerhsconcentration
//
// Design rules (profiles/README.md has the measurements behind them):
//  * the face direction is a template parameter: every shared-memory offset is an immediate;
//  * model selectors (interpolation / averaging / floor / mobility characters, flux type,
//    free energy) are compile-time constants for the parameter sets of the shipped decks
//    (SelFixed) and runtime values otherwise (SelRuntime): same code, the selector chains fold;
//  * transverse side gradients use one central difference per cell, q(x+e)-q(x-e), instead
//    of the four one-sided differences of quatgrad_side (identical in exact arithmetic);
//  * 1/sqrt, 1/x through straight-line Newton (fastmath.cuh); the CALPHAD face diffusivity
//    M(c) c(1-c) d2f/dc2 without division; explicit fma() on polynomial chains (the library
//    is built with --fmad=false: rounding is fixed by the source, not by the optimiser);
//  * where the result is a small difference of large terms (projected quaternion
//    divergence, symmetry correction, composition flux differences) the reference's
//    operation order is kept term by term, because only matching rounding gives 1e-12 there.
#pragma once
#include "calphad.cuh"
#include "fastmath.cuh"
#include "params.h"
#include "pointwise.cuh"
#include "rhs_common.cuh"

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
// -DAMPE_NO_STREAM_DIFFS keeps all 2*ND*Q quaternion differences of a cell live (A/B builds)
#ifdef AMPE_NO_STREAM_DIFFS
#define AMPE_STREAM_DIFFS &&false
#else
#define AMPE_STREAM_DIFFS
#endif

// -DAMPE_ATAN_FAST: straight-line atan_fast (fastmath.cuh, <= 1.9 ulp) in the bias-well term instead of
// CUDA's atan (A/B builds; the default stays CUDA's until measured and parity-checked on the GPU)
#ifdef AMPE_ATAN_FAST
#define AMPE_ATAN(x) atan_fast(x)
#else
#define AMPE_ATAN(x) atan(x)
#endif

// does the parameter record select exactly the compile-time model SEL?
template <class SEL>
static bool sel_matches(const Params& p)
{
   return p.with_phase == SEL::with_phase && p.evolve_quat == SEL::evolve_quat &&
          p.flux_type == SEL::flux_type && p.free_energy == SEL::free_energy &&
          p.modulus_from_cells == SEL::modulus_from_cells && p.knumber == SEL::knumber &&
          p.libm_trig == SEL::libm_trig && p.energy_interp == SEL::energy_interp &&
          p.diffusion_interp == SEL::diffusion_interp && p.orient_interp1 == SEL::orient_interp1 &&
          p.orient_interp2 == SEL::orient_interp2 && p.avg_func == SEL::avg_func &&
          p.conc_avg_func == SEL::conc_avg_func && p.grad_floor_type == SEL::grad_floor_type &&
          p.quat_mobility_func == SEL::quat_mobility_func;
}

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
      const double r = rsqrt_fast((g2 > floor2) ? g2 : floor2);
      return (g2 > floor2) ? r : max_normi;
   }
   return eval_grad_normi_rare(g2, floor_type, floor2, max_normi);
}
// interp_func 'p' with explicit fma
AMPE_DEV double interp3(double phi, char type)
{
   if (type == 'p') {
      const double t = clamp01(phi);
      return t * t * t * fma(t, fma(6.0, t, -15.0), 10.0);
   }
   return interp_func(phi, type);
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
      m[sp] = exp_fast<false>(dG * t.RTinv);  // range checked on the host: mobility_exponent_bound (ctx.cu)
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

// offsets (in doubles) from a staged cell to the same cell one plane down / up along z.
// Tile kernels pass the constant +-SX*SY; the marching kernel passes the ring-slot offsets.
struct ZOff {
   int m, p;
};

// what one face contributes to the divergences of the two cells it separates
struct FaceVal {
   double fc;  // quaternion face coefficient        compute_face_coef
   double pf;  // phase flux, non-simple stencils    anisotropic_gradient_flux / compute_flux_isotropic
   double cf;  // composition flux                   add_flux / concentrationflux
};
// the 2*ND faces of one cell
template <int ND>
struct CellFaces {
   double fcl[ND], fcu[ND], cfl[ND], cfu[ND];
   double pfl[ND], pfu[ND];
};

// TT supplies: ND, Q, CONC, SYMM, WT, SEL, HAS_PF, S (doubles per staged field), SX (row pitch),
// field offsets O_PHI, O_T, O_Q, O_C, O_CL, O_CA.
template <class TT>
struct Rhs3 {
   using SEL = typename TT::SEL;
   static constexpr int ND = TT::ND, Q = TT::Q, CONC = TT::CONC, S = TT::S;
   static constexpr int CF = TT::CF;      // composition-flux form evaluated by this launch (0: none)
   static constexpr int PART = TT::PART;  // 2: composition outputs only (rhs_march.cuh)
   static constexpr bool SYMM = TT::SYMM, WT = TT::WT;
   static constexpr int QN = (Q > 0) ? Q : 1;

   // staged offset of one step up / down in direction n
   AMPE_DEV static int up(int n, ZOff z) { return n == 0 ? 1 : (n == 1 ? TT::SX : z.p); }
   AMPE_DEV static int dn(int n, ZOff z) { return n == 0 ? -1 : (n == 1 ? -TT::SX : z.m); }

   // difference of direction a across the face between staged cells cm (lower) and c (upper);
   // symmetric mode rotates the lower neighbour by the face's rotation index (quatdiffs_symm)
   template <int a>
   AMPE_DEV static void qdiff(const double* s, const int* s_iq, const double (*s_qr)[4],
                              const int* s_conj, int c, int cm, double* d)
   {
      const double* sq = s + TT::O_Q;
      if constexpr (Q == 0) {
      } else if constexpr (SYMM) {
         double q2[QN], q2p[QN];
#pragma unroll
         for (int m = 0; m < Q; m++) q2[m] = sq[m * S + cm];
         symm_rotate<Q>(q2, s_iq[a * S + c], q2p, s_qr, s_conj);
#pragma unroll
         for (int m = 0; m < Q; m++) d[m] = sq[m * S + c] - q2p[m];
      } else {
#pragma unroll
         for (int m = 0; m < Q; m++) d[m] = sq[m * S + c] - sq[m * S + cm];
      }
   }

   // |grad q|^2 on the face between cm and c (normal direction a): quatgrad_side[_symm] +
   // modulus of the side gradient (3d/quatgrad.m4:376-720)
   template <int a>
   AMPE_DEV static double face_grad2(const double* s, const int* s_iq, const double (*s_qr)[4],
                                     const int* s_conj, const Params& p, int c, int cm, ZOff z)
   {
      const double* sq = s + TT::O_Q;
      double g2 = 0.0;
#pragma unroll
      for (int n = 0; n < ND; n++) {
         const int un = up(n, z), dnn = dn(n, z);
         if (n == a) {
            double d[QN];
            qdiff<a>(s, s_iq, s_qr, s_conj, c, cm, d);
#pragma unroll
            for (int m = 0; m < Q; m++) {
               const double g = p.dinv[a] * d[m];
               g2 = fma(g, g, g2);
            }
         } else if (SYMM && Q > 1) {
            // quatgrad_side_symm: the four one-sided differences are rotated into the frame
            // of the face before they are averaged (reference operation order)
            double g[QN];
            const int ct = c + un, cmt = cm + un;
            double d1[QN], d1p[QN], d2[QN], d2p[QN], d3[QN], d4[QN], d4p[QN], d0[QN];
            if (n == 0) {
               qdiff<0>(s, s_iq, s_qr, s_conj, ct, ct + dnn, d1);
               qdiff<0>(s, s_iq, s_qr, s_conj, cmt, cmt + dnn, d2);
               qdiff<0>(s, s_iq, s_qr, s_conj, cm, cm + dnn, d3);
               qdiff<0>(s, s_iq, s_qr, s_conj, c, c + dnn, d0);
            } else if (n == 1) {
               qdiff<1>(s, s_iq, s_qr, s_conj, ct, ct + dnn, d1);
               qdiff<1>(s, s_iq, s_qr, s_conj, cmt, cmt + dnn, d2);
               qdiff<1>(s, s_iq, s_qr, s_conj, cm, cm + dnn, d3);
               qdiff<1>(s, s_iq, s_qr, s_conj, c, c + dnn, d0);
            } else {
               qdiff<2>(s, s_iq, s_qr, s_conj, ct, ct + dnn, d1);
               qdiff<2>(s, s_iq, s_qr, s_conj, cmt, cmt + dnn, d2);
               qdiff<2>(s, s_iq, s_qr, s_conj, cm, cm + dnn, d3);
               qdiff<2>(s, s_iq, s_qr, s_conj, c, c + dnn, d0);
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
               const double g = p.p25inv[n] * ((qm[cm + un] - qm[cm + dnn]) + (qm[c + un] - qm[c + dnn]));
               g2 = fma(g, g, g2);
            }
         }
      }
      return g2;
   }

   // ---- one face: between staged cells cm (lower) and c (upper), normal direction a ----------
   //  gface  index in the lagged arrays;  inrange  the face bounds a cell of the domain;
   //  wr  this work item refreshes the lagged arrays
   // -DAMPE_FACE_NOINLINE: one out-of-line instance per direction instead of one inlined copy per call site (the
   // main face pass and the tile-edge pass instantiate the same direction twice: the symmetry-aware tile kernel is
   // 5500 SASS instructions and stalls on instruction fetch)
#ifdef AMPE_FACE_NOINLINE
#define AMPE_FACE_DEV __device__ __noinline__
#else
#define AMPE_FACE_DEV AMPE_DEV
#endif
   template <int a>
   AMPE_FACE_DEV static FaceVal face(const FusedArgs& A, const double* s, const int* s_iq,
                                const double (*s_qr)[4], const int* s_conj, int c, int cm, ZOff z,
                                long long gface, bool inrange, bool wr)
   {
      const Params& p = A.p;
      FaceVal out;
      out.fc = 0.0, out.pf = 0.0, out.cf = 0.0;
      const double phi_c = s[TT::O_PHI + c], phi_m = s[TT::O_PHI + cm];
      const bool evolve_quat = (Q > 0) && AMPE_SEL(evolve_quat);
      const int flux_type = AMPE_SEL(flux_type);

      // ---- quaternion face coefficient (compute_face_coef, quatfacops.m4) ----
      if constexpr (Q > 0) if (evolve_quat) {
         double normi;
         if (A.use_lag) {
            normi = inrange ? A.lagN[a][gface] : 0.0;
         } else {
            const double g2 = face_grad2<a>(s, s_iq, s_qr, s_conj, p, c, cm, z);
            normi = grad_normi3(g2, AMPE_SEL(grad_floor_type), p.floor2, p.max_normi);
            if (wr) A.lagN[a][gface] = normi;
         }
         const double phia = average3(phi_m, phi_c, AMPE_SEL(avg_func));
         const double tempa = WT ? 0.5 * (s[TT::O_T + cm] + s[TT::O_T + c]) : 0.5 * (p.T_uniform + p.T_uniform);
         const double diff = p.misorientation_factor * tempa * interp3(phia, AMPE_SEL(orient_interp1));
         const double hphi2 = interp3(phia, AMPE_SEL(orient_interp2));
         out.fc = -normi * diff - p.epsq2 * hphi2;
      }

      // ---- phase flux, non-simple stencils (2D) ----
      if constexpr (TT::HAS_PF && ND == 3 && Q == 4) if (flux_type == AMPE_FLUX_ANISOTROPIC) {
         // anisotropic_gradient_flux, 3d/quatrhs.m4:180-349 with compute_dgamma (:149-177): cubic harmonic of the
         // interface normal rotated into the crystal frame by the face-averaged quaternion (nu = eps4).  Evaluated
         // with the reference's own operations (division, sqrt): only the runtime-selector kernels carry it.
         const double* sp = s + TT::O_PHI;
         const double* sq = s + TT::O_Q;
         double g[3];
#pragma unroll
         for (int t = 0; t < 3; t++) {
            if (t == a) {
               g[t] = (phi_c - phi_m) * p.dinv[t];
            } else {
               const int ut = up(t, z), dt_ = dn(t, z);
               g[t] = 0.25 * (sp[cm + ut] - sp[cm + dt_] + sp[c + ut] - sp[c + dt_]) * p.dinv[t];
            }
         }
         const double eps4 = p.nu, epsilon = p.epsilon_phase;
         const double factor = 4. * eps4 / (1. - 3. * eps4);
         const double gphi2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
         double dg_a = 0.0, n4 = 0.0;
         if (fabs(gphi2) > (double)1.e-12f) {
            const double nni = 1. / sqrt(gphi2);
            const double n[4] = {0., g[0] * nni, g[1] * nni, g[2] * nni};
            double qq[4], qp[4], qtmp[4], np[4], dg[4], dgamma[4];
#pragma unroll
            for (int m = 0; m < 4; m++) qq[m] = 0.5 * (sq[m * S + cm] + sq[m * S + c]);
            qp[0] = qq[0], qp[1] = -qq[1], qp[2] = -qq[2], qp[3] = -qq[3];
            quatmult4(n, qp, qtmp);
            quatmult4(qq, qtmp, np);
            const double a2 = np[1] * np[1], a3 = np[2] * np[2], a4 = np[3] * np[3];
            n4 = a2 * a2 + a3 * a3 + a4 * a4;
            dg[0] = 0.;
            dg[1] = np[1] * (np[1] * np[1] - n4);
            dg[2] = np[2] * (np[2] * np[2] - n4);
            dg[3] = np[3] * (np[3] * np[3] - n4);
            quatmult4(dg, qq, qtmp);
            quatmult4(qp, qtmp, dgamma);
            dg_a = dgamma[a + 1];
         }
         // (degenerate gradient: the reference's dgamma component of this direction is 0 and sqrt(gphi2) ~ 0)
         const double gamma = epsilon * (1. - 3. * eps4) * (1. + factor * n4);
         out.pf = gamma * gamma * g[a] + 16. * epsilon * gamma * eps4 * sqrt(gphi2) * dg_a;
      }
      if constexpr (TT::HAS_PF && ND == 2 && Q > 0 && a < 2) if (flux_type == AMPE_FLUX_ANISOTROPIC) {
         // anisotropic_gradient_flux, 2d/quatrhs.m4:154-256
         constexpr int st = (a == 0) ? TT::SX : 1;
         const double* sp = s + TT::O_PHI;
         const double dn_ = (phi_c - phi_m) * p.dinv[a];
         const double dt = 0.25 * (sp[cm + st] - sp[cm - st] + sp[c + st] - sp[c - st]) * p.dinv[1 - a];
         const double dphidx = (a == 0) ? dn_ : dt;
         const double dphidy = (a == 0) ? dt : dn_;
         double qa = 0.5 * (s[TT::O_Q + cm] + s[TT::O_Q + c]);
         qa = (qa > 1.0) ? 1.0 : ((qa < -1.0) ? -1.0 : qa);
         double sn, cs;
         if (AMPE_SEL(knumber) == 4 && !AMPE_SEL(libm_trig)) {
            // cos/sin of 4(theta - psi) without atan/acos/sincos: theta = atan(y/x) enters only
            // through cos 4theta = 1 - 8 x^2 y^2 / r^4, sin 4theta = 4 x y (x^2 - y^2) / r^4, and
            // psi = acos(q) through the Chebyshev polynomials cos 4psi = T4(q),
            // sin 4psi = sqrt((1-q)(1+q)) U3(q)  (DESIGN.md "Transcendentals")
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
         out.pf = (a == 0) ? fma(e2, dphidx, -(ed * dphidy)) : fma(e2, dphidy, ed * dphidx);
      }
      if constexpr (TT::HAS_PF && ND == 2 && a < 2) if (flux_type == AMPE_FLUX_ISOTROPIC) {
         // compute_flux_isotropic, 2d/quatrhs.m4:106-151
         constexpr int st = (a == 0) ? TT::SX : 1;
         const double* sp = s + TT::O_PHI;
         out.pf = p.iso_dinv[a] * ((sp[c - st] - sp[cm - st]) + (phi_c - phi_m) * 10.0 + (sp[c + st] - sp[cm + st]));
      }

      // ---- composition flux ----
      if constexpr (CF == AMPE_CONC_EBS) {
         const double* scl = s + TT::O_CL;
         const double* sca = s + TT::O_CA;
         double Dl, Da;
         if (A.use_lag) {
            Dl = inrange ? A.lagD0[a][gface] : 0.0;
            Da = inrange ? A.lagD1[a][gface] : 0.0;
         } else {
            if (AMPE_SEL(free_energy) == AMPE_FE_CALPHAD) {
               // MobilityCompositionDiffusionStrategy.cc:296-327 + setPFMDiffOnPatch
               const double c_l = 0.5 * (scl[c] + scl[cm]);
               const double c_a = 0.5 * (sca[c] + sca[cm]);
               const double dl = ebs_phase_diffusivity(p.ct, 0, c_l);
               const double da = ebs_phase_diffusivity(p.ct, 1, c_a);
               const double phia = average3(phi_c, phi_m, AMPE_SEL(conc_avg_func));
               const double hphi = interp3(phia, AMPE_SEL(diffusion_interp));
               Dl = (1. - hphi) * dl;
               Da = hphi * da;
            } else {
               // diffusion_type "temperature_dependent" (quadratic free energy): concentration_pfmdiffusion_of_temperature
               // (2d/concentrationdiffusion.m4:341-430) with the uniform-T Arrhenius factors of ampe_derive_params
               const double vphi = average3(phi_m, phi_c, AMPE_SEL(avg_func));
               const double hphi = interp3(vphi, AMPE_SEL(diffusion_interp));
               Dl = (1.0 - hphi) * p.D_liquid;
               Da = hphi * p.D_solid;
            }
            if (wr) {
               A.lagD0[a][gface] = Dl;
               A.lagD1[a][gface] = Da;
            }
         }
         // add_flux (3d/flux.m4:53-66), liquid then solid (EBSCompositionRHSStrategy.cc:258-292)
         double fl = p.dinv[a] * (Dl * (scl[c] - scl[cm]));
         fl = fl + p.dinv[a] * (Da * (sca[c] - sca[cm]));
         out.cf = fl;
      } else if constexpr (CF == AMPE_CONC_KKS) {
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
         out.cf = p.dinv[a] * (D0 * (s[TT::O_C + c] - s[TT::O_C + cm]) + Dp * (phi_c - phi_m));
      }
      return out;
   }

   // ---- one cell: divergences of the face values + pointwise terms -> global ------------------
   AMPE_DEV static void cell(const FusedArgs& A, const double* s, const int* s_iq, const double (*s_qr)[4],
                             const int* s_conj, int c, ZOff z, const CellFaces<ND>& F, long long gcell,
                             long long ncell)
   {
      const Params& p = A.p;
      const bool evolve_quat = (Q > 0) && AMPE_SEL(evolve_quat);
      const int flux_type = AMPE_SEL(flux_type);
      const int free_energy = AMPE_SEL(free_energy);
      const double phi = s[TT::O_PHI + c];
      const double temp = WT ? s[TT::O_T + c] : p.T_uniform;
      const double* sp = s + TT::O_PHI;
      const double* sq = s + TT::O_Q;

      // quaternion differences on the lower / upper faces (symmetric ones in SYMM mode).
      // Without symmetry the cell streams direction by direction: the differences of one
      // direction feed the gradient modulus and the divergence accumulators div_m right away, so
      // that only 2Q differences are live instead of 2*ND*Q (the 3D kernel is register-bound).
      // The accumulation order per component (a = 0, 1, 2) is the reference's either way.
      constexpr bool STREAM = !SYMM && (Q > 0) AMPE_STREAM_DIFFS;
      double dlo[STREAM ? 1 : ND][QN], dup[STREAM ? 1 : ND][QN];
      double divm[QN];
      double sm_acc = 0.0;
      if constexpr (Q > 0) if (evolve_quat) {
         if constexpr (STREAM) {
            auto dir = [&](auto dtag, int cu, int cd) {
               constexpr int a = decltype(dtag)::value;
               double lo[QN], hi[QN];
               qdiff<a>(s, s_iq, s_qr, s_conj, c, cd, lo);
               qdiff<a>(s, s_iq, s_qr, s_conj, cu, c, hi);
               if (AMPE_SEL(modulus_from_cells)) {
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = (hi[m] + lo[m]) * p.p5inv[a];
                     sm_acc = fma(g, g, sm_acc);
                  }
               } else {
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = p.dinv[a] * lo[m];
                     sm_acc = fma(g, g, sm_acc);
                  }
#pragma unroll
                  for (int m = 0; m < Q; m++) {
                     const double g = p.dinv[a] * hi[m];
                     sm_acc = fma(g, g, sm_acc);
                  }
               }
#pragma unroll
               for (int m = 0; m < Q; m++) {
                  const double fu = F.fcu[a] * (p.dinv[a] * hi[m]);
                  const double fl = F.fcl[a] * (p.dinv[a] * lo[m]);
                  divm[m] = (a == 0) ? (fu - fl) * p.dinv[a] : divm[m] + (fu - fl) * p.dinv[a];
               }
            };
            dir(std::integral_constant<int, 0>(), c + 1, c - 1);
            dir(std::integral_constant<int, 1>(), c + TT::SX, c - TT::SX);
            if constexpr (ND == 3) dir(std::integral_constant<int, 2>(), c + z.p, c + z.m);
         } else {
            qdiff<0>(s, s_iq, s_qr, s_conj, c, c - 1, dlo[0]);
            qdiff<0>(s, s_iq, s_qr, s_conj, c + 1, c, dup[0]);
            qdiff<1>(s, s_iq, s_qr, s_conj, c, c - TT::SX, dlo[1]);
            qdiff<1>(s, s_iq, s_qr, s_conj, c + TT::SX, c, dup[1]);
            if constexpr (ND == 3) {
               qdiff<2>(s, s_iq, s_qr, s_conj, c, c + z.m, dlo[ND - 1]);
               qdiff<2>(s, s_iq, s_qr, s_conj, c + z.p, c, dup[ND - 1]);
            }
         }
      }

      double phase_rhs = 0.0;
      if (PART != 2 && AMPE_SEL(with_phase)) {
         // computerhspbg (2d/quatrhs.m4:328-402, 3d:430-512)
         double diff_term;
         if (!TT::HAS_PF || flux_type == AMPE_FLUX_SIMPLE) {
            // gradient_flux inlined: flux = (phi(c) - phi(c-e))*(eps2/h)
            diff_term = ((sp[c + 1] - phi) * p.eps2_dinv[0] - (phi - sp[c - 1]) * p.eps2_dinv[0]) * p.dinv[0];
            diff_term = diff_term + ((sp[c + TT::SX] - phi) * p.eps2_dinv[1] -
                                     (phi - sp[c - TT::SX]) * p.eps2_dinv[1]) * p.dinv[1];
            if constexpr (ND == 3)
               diff_term = diff_term + ((sp[c + z.p] - phi) * p.eps2_dinv[2] -
                                        (phi - sp[c + z.m]) * p.eps2_dinv[2]) * p.dinv[2];
         } else {
            diff_term = (F.pfu[0] - F.pfl[0]) * p.dinv[0];
            diff_term = diff_term + (F.pfu[1] - F.pfl[1]) * p.dinv[1];
            if constexpr (ND == 3) diff_term = diff_term + (F.pfu[ND - 1] - F.pfl[ND - 1]) * p.dinv[ND - 1];
         }
         double rhs = diff_term;
         rhs = rhs - p.phi_well_scale * deriv_well_func(phi, 'd');
         if constexpr (Q > 0) if (evolve_quat) {
            // gradient modulus (quatgrad_cell[_symm] + quatgrad_modulus, or from sides compact)
            double sm = 0.0;
            if constexpr (STREAM) {
               sm = AMPE_SEL(modulus_from_cells) ? sqrt_fast(sm_acc) : sqrt_fast(0.5 * sm_acc);
            } else {
               if (AMPE_SEL(modulus_from_cells)) {
#pragma unroll
                  for (int a = 0; a < ND; a++) {
                     double du[QN];
                     if (SYMM && Q > 1) {
                        symm_rotate<Q>(dup[a], -s_iq[a * S + c + up(a, z)], du, s_qr, s_conj);
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
            const double m = p.bias_coeff * AMPE_ATAN(p.bias_gamma * (p.meltingT - temp));
            rhs = rhs + m * phi * (1.0 - phi);
         } else if (free_energy == AMPE_FE_DELTAT) {
            // computerhsdeltatemperature (2d/quatrhs.m4:893-940, 3d/quatrhs.m4:1040-1087): the temperature is
            // smoothed over the cell and its face neighbours; 3D: woff = 0.25/6. is a REAL*4 expression
            double wtemp;
            if constexpr (WT) {
               const double* sT = s + TT::O_T;
               if constexpr (ND == 2) {
                  wtemp = 0.75 * temp + 0.0625 * (sT[c - 1] + sT[c - TT::SX] + sT[c + 1] + sT[c + TT::SX]);
               } else {
                  const double woff = (double)(0.25f / 6.f);
                  wtemp = 0.75 * temp + woff * (sT[c - 1] + sT[c - TT::SX] + sT[c + 1] + sT[c + TT::SX] +
                                                sT[c + z.m] + sT[c + z.p]);
               }
            } else {
               if constexpr (ND == 2) {
                  wtemp = 0.75 * temp + 0.0625 * (temp + temp + temp + temp);
               } else {
                  const double woff = (double)(0.25f / 6.f);
                  wtemp = 0.75 * temp + woff * (temp + temp + temp + temp + temp + temp);
               }
            }
            const double m = p.deltaT_alpha * (p.meltingT - wtemp);
            rhs = rhs + m * deriv_interp_func(phi, AMPE_SEL(energy_interp));
         } else if ((CONC == AMPE_CONC_EBS || CONC == AMPE_CONC_KKS) && free_energy == AMPE_FE_CALPHAD) {
            // (either composition flux form: rhs_form "ebs" examples/AuNi_*, "kks" tests/KKScomposition)
            // CALPHADFreeEnergyStrategyBinary.cc:321-323, 638-663: (f_l-f_a) - mu (c_l-c_a) comes
            // from the KKS kernel, which has the logarithms of the converged c_l, c_a at hand
            const double hp = deriv_interp_func(phi, AMPE_SEL(energy_interp));
            rhs += hp * A.df[gcell];
         } else if ((CONC == AMPE_CONC_KKS || CONC == AMPE_CONC_EBS) && free_energy == AMPE_FE_QUADRATIC) {
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
         // compute_flux_from_gradq + compute_lambda_flux + add_quat_proj_op.
         // div(fc grad q) per component in the reference's operation order: the projection
         // below (and the symmetry correction) cancel most of it, so parity needs its exact
         // rounding.  compute_lambda_flux sums the same face differences scaled by 0.5/h, which
         // is exactly half of 1/h: lambda = -(q.div)/(2|q|^2) bit for bit, and
         // 2 q lambda = -q (q.div)/|q|^2 needs no second accumulation.
         double qc[QN];
         double qdiv = 0.0, sumq2 = 0.0;
#pragma unroll
         for (int m = 0; m < Q; m++) {
            qc[m] = sq[m * S + c];
            if constexpr (!STREAM) {
               double dv = 0.0;
#pragma unroll
               for (int a = 0; a < ND; a++) {
                  const double fu = F.fcu[a] * (p.dinv[a] * dup[a][m]);
                  const double fl = F.fcl[a] * (p.dinv[a] * dlo[a][m]);
                  dv = (a == 0) ? (fu - fl) * p.dinv[a] : dv + (fu - fl) * p.dinv[a];
               }
               divm[m] = dv;
            }
            qdiv = qdiv + qc[m] * divm[m];
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
                  symm_rotate<Q>(dup[a], -s_iq[a * S + c + up(a, z)], dpr[a], s_qr, s_conj);
               else
                  dpr[a][0] = dup[a][0];
            }
#pragma unroll
            for (int m = 0; m < Q; m++) {
               double tt = 0.0;
#pragma unroll
               for (int a = 0; a < ND; a++) {
                  const double nsd_u = sq[m * S + c + up(a, z)] - qc[m];
                  const double nsd_l = qc[m] - sq[m * S + c + dn(a, z)];
                  const double term =
                      p.dinv2[a] * (F.fcu[a] * (nsd_u - dpr[a][m]) - F.fcl[a] * (nsd_l - dlo[a][m]));
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

      if constexpr (CF != 0) {
         // computerhsconcentration (3d/concentrationrhs.m4:412-458)
         double sm = p.dinv[0] * (F.cfu[0] - F.cfl[0]) + p.dinv[1] * (F.cfu[1] - F.cfl[1]);
         if constexpr (ND == 3) sm = sm + p.dinv[2] * (F.cfu[ND - 1] - F.cfl[ND - 1]);
         A.out_c[gcell] = p.conc_mobility * sm;
      }

      if constexpr (WT && PART != 2) {
         // computerhstemp + laplacian (2d/quatrhs.m4:787-803, 2d/laplacian.m4:37-52)
         const double* sT = s + TT::O_T;
         const double dtx = (sT[c - 1] - 2.0 * temp + sT[c + 1]);
         const double dty = (sT[c - TT::SX] - 2.0 * temp + sT[c + TT::SX]);
         double dterm = dtx * p.dinv2[0] + dty * p.dinv2[1];
         if constexpr (ND == 3) {
            const double dtz = (sT[c + z.m] - 2.0 * temp + sT[c + z.p]);
            dterm = dterm + dtz * p.dinv2[2];
         }
         double r = p.thermal_diffusivity * dterm;
         if (AMPE_SEL(with_phase)) r = r + p.latent_over_cp * phase_rhs;
         A.out_T[gcell] = r;
      }
   }
};

}  // namespace ampe
